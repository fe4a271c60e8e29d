mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2m_pytest_gpu.log 2>&1; tail -4 gpurun_out/r2m_pytest_gpu.log
timeout 300 python tools/launch_overhead.py > gpurun_out/r2m_launch_overhead.log 2>&1; cat gpurun_out/r2m_launch_overhead.log
timeout 600 python tools/quick_bench.py heat3d:32:1024x1024x1024 jacobi3d:32:1024x1024x1024 jacobi2d:64:16384x16384 sobel2d:1:32768x32768 blur:1:32768x32768 seidel2d:2:16384x16384 jacobi2d:1:16384x16384 > gpurun_out/r2m_sweep.log 2>&1; cat gpurun_out/r2m_sweep.log
