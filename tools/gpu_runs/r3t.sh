mkdir -p gpurun_out
python tools/ab_libraries.py > gpurun_out/r3t_ab.log 2>&1; cat gpurun_out/r3t_ab.log
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r3t_pytest_gpu.log 2>&1; tail -4 gpurun_out/r3t_pytest_gpu.log
