mkdir -p gpurun_out
timeout 600 python tools/quick_bench.py jacobi2d:64:16384x16384 jacobi2d:64:16384x16384:depth=6 jacobi2d:64:16384x16384:depth=10 jacobi2d:64:16384x16384:depth=12 > gpurun_out/r3g_sweep.log 2>&1; cat gpurun_out/r3g_sweep.log
