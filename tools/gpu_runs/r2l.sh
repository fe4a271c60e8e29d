mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2l_pytest_gpu.log 2>&1; tail -8 gpurun_out/r2l_pytest_gpu.log
timeout 600 python tools/quick_bench.py heat3d:32:1024x1024x1024 jacobi3d:32:1024x1024x1024 denoise3d:1:768x768x768 heat3d:1:1024x1024x1024 denoise3d:1:768x768x768:fast=1 > gpurun_out/r2l_sweep.log 2>&1; cat gpurun_out/r2l_sweep.log
