mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/r2d_pytest_gpu.log 2>&1; tail -4 gpurun_out/r2d_pytest_gpu.log
D3=denoise3d:1:768x768x768
D2=denoise2d:1:32768x32768
timeout 900 python tools/quick_bench.py $D3 $D3:inline=0 $D3:tile=128x16:threads=512 $D3:tile=128x16:threads=512:prefetch=1 $D3:tile=128x32:threads=512:prefetch=1 $D2 $D2:inline=0 $D2:vec=2 $D2:vec=2:inline=0 sobel2d:1:32768x32768 sobel2d:1:32768x32768:inline=1 heat3d:32:1024x1024x1024 > gpurun_out/r2d_sweep.log 2>&1; cat gpurun_out/r2d_sweep.log
timeout 200 python tools/fpga_layout_bench.py > gpurun_out/r2d_fpga_layout_bench.log 2>&1; tail -5 gpurun_out/r2d_fpga_layout_bench.log
