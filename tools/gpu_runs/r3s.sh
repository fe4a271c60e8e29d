mkdir -p gpurun_out
timeout 600 python tools/quick_bench.py heat3d:32:1024x1024x1024 jacobi3d:32:1024x1024x1024 denoise3d:1:768x768x768 denoise3d:1:768x768x768:fast=1 heat3d:1:1024x1024x1024 heat3d:32:1024x1024x128 > gpurun_out/r3s_sweep.log 2>&1; cat gpurun_out/r3s_sweep.log
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r3s_pytest_gpu.log 2>&1; tail -4 gpurun_out/r3s_pytest_gpu.log
