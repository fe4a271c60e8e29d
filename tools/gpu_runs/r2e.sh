mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2e_pytest_gpu.log 2>&1; tail -6 gpurun_out/r2e_pytest_gpu.log
timeout 200 python tools/fpga_layout_bench.py > gpurun_out/r2e_fpga_layout_bench.log 2>&1; tail -5 gpurun_out/r2e_fpga_layout_bench.log
