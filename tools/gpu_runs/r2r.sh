mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2r_pytest_gpu.log 2>&1; tail -4 gpurun_out/r2r_pytest_gpu.log
for p in 16 12 24; do SODA_CUDA_PIECES=$p REPS=7 timeout 300 python tools/quick_bench.py jacobi2d:64:16384x16384:e2e=1; done > gpurun_out/r2r_e2e.log 2>&1; cat gpurun_out/r2r_e2e.log
REPS=7 timeout 300 python tools/quick_bench.py heat3d:32:1024x1024x1024:e2e=1 sobel2d:1:32768x32768:e2e=1 >> gpurun_out/r2r_e2e.log 2>&1; tail -2 gpurun_out/r2r_e2e.log
