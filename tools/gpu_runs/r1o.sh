mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_fullsize_gpu.py tests/test_fastmath_gpu.py -q > gpurun_out/r1o_pytest_fullsize.log 2>&1
tail -25 gpurun_out/r1o_pytest_fullsize.log
REPS=5 timeout 600 python tools/quick_bench.py denoise3d:1:768x768x768 denoise3d:1:768x768x768:fast=1 denoise2d:1:32768x32768 denoise2d:1:32768x32768:fast=1 heat3d:32:1024x1024x1024 heat3d:32:1024x1024x1024:fast=1 jacobi2d:64:16384x16384 jacobi2d:64:16384x16384:fast=1 > gpurun_out/r1o_fastmath.log 2>&1
cat gpurun_out/r1o_fastmath.log
