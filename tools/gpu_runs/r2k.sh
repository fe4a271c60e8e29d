mkdir -p gpurun_out
for p in 16 24 32 48 64; do SODA_CUDA_PIECES=$p REPS=7 timeout 300 python tools/quick_bench.py jacobi2d:64:16384x16384:e2e=1; done > gpurun_out/r2k_e2e_pieces.log 2>&1; cat gpurun_out/r2k_e2e_pieces.log
