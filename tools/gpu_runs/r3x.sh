mkdir -p gpurun_out
(echo "# default build"; python tools/fpga_layout_bench.py 2>&1 | grep staged
 echo "# pack with 8-byte bank words and rotated shared reads"; SODA_FPGA_LIBRARY=$PWD/tools/ab/w8/libsoda_fpga_layout.so python tools/fpga_layout_bench.py 2>&1 | grep "staged"
 echo "# its equivalence tests"; SODA_FPGA_LIBRARY=$PWD/tools/ab/w8/libsoda_fpga_layout.so timeout 200 python -m pytest tests/test_fpga_layout_gpu.py tests/test_zz_fpga_staged_gpu.py -m gpu -q 2>&1 | tail -2) > gpurun_out/r3x_pack_word8.log 2>&1
cat gpurun_out/r3x_pack_word8.log
