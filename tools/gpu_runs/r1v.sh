mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_fpga_layout_gpu.py -q > gpurun_out/r1v_pytest_fpga.log 2>&1; tail -12 gpurun_out/r1v_pytest_fpga.log
timeout 120 python tools/fpga_layout_bench.py > gpurun_out/r1v_fpga_layout_bench.log 2>&1; cat gpurun_out/r1v_fpga_layout_bench.log | tail -5
