mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_fpga_layout_gpu.py tests/test_zz_fpga_staged_gpu.py -q > gpurun_out/r2j_pytest_fpga.log 2>&1; tail -5 gpurun_out/r2j_pytest_fpga.log
timeout 200 python tools/fpga_layout_bench.py > gpurun_out/r2j_fpga_layout_bench.log 2>&1; tail -5 gpurun_out/r2j_fpga_layout_bench.log
