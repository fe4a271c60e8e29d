mkdir -p gpurun_out
timeout 900 python tools/quick_bench.py denoise2d:1:32768x32768:fast=1 denoise2d:1:32768x32768:fast=1:groups=3 denoise2d:1:32768x32768:groups=3 denoise2d:1:32768x32768:groups=3:prefetch=8 denoise2d:1:32768x32768:groups=3:threads=256 > gpurun_out/r2w_sweep.log 2>&1; cat gpurun_out/r2w_sweep.log
