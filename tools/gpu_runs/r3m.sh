mkdir -p gpurun_out
timeout 600 python tools/quick_bench.py denoise3d:1:768x768x768 denoise2d:1:32768x32768 > gpurun_out/r3m_sweep.log 2>&1; cat gpurun_out/r3m_sweep.log
timeout 900 python -m pytest tests/test_div_exact_gpu.py tests/test_rsqrt_exact.py tests/test_parity_gpu.py tests/test_fullsize_gpu.py tests/test_production_unaligned_gpu.py tests/test_fastmath_gpu.py -m gpu -q -x > gpurun_out/r3m_pytest.log 2>&1; tail -5 gpurun_out/r3m_pytest.log
