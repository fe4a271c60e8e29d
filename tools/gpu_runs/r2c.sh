mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/r2c_pytest_gpu.log 2>&1; tail -4 gpurun_out/r2c_pytest_gpu.log
timeout 600 python tools/quick_bench.py heat3d:32:1024x1024x1024 jacobi3d:32:1024x1024x1024 denoise3d:1:768x768x768 denoise2d:1:32768x32768 heat3d:1:1024x1024x1024 jacobi2d:64:16384x16384 sobel2d:1:32768x32768 blur:1:32768x32768 denoise3d:1:768x768x768:fast=1 denoise2d:1:32768x32768:fast=1 heat3d:32:1024x1024x1024:fast=1 jacobi2d:64:16384x16384:fast=1 > gpurun_out/r2c_sweep.log 2>&1; cat gpurun_out/r2c_sweep.log
bash tools/ncu_capture.sh r2c_denoise3d soda_denoise3d denoise3d:1:768x768x768
bash tools/ncu_capture.sh r2c_denoise2d soda_denoise2d denoise2d:1:32768x32768
bash tools/ncu_capture.sh r2c_heat3d soda_heat3d heat3d:32:1024x1024x1024
ls -la gpurun_out | tail -8
