mkdir -p gpurun_out
timeout 600 python tools/quick_bench.py jacobi2d:64:16384x16384 jacobi2d:64:16384x16384:threads=64:prefetch=12 jacobi2d:64:16384x16384:threads=64:prefetch=12:min_blocks=8 jacobi2d:64:16384x16384:threads=64:prefetch=12:min_blocks=7 jacobi2d:64:16384x16384:threads=64:min_blocks=7 > gpurun_out/r2i_sweep.log 2>&1; cat gpurun_out/r2i_sweep.log
