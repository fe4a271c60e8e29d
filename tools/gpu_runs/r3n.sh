mkdir -p gpurun_out
timeout 600 python tools/quick_bench.py denoise3d:1:768x768x768 denoise3d:1:768x768x768:fast=1 denoise2d:1:32768x32768 >> gpurun_out/r3n_sweep.log 2>&1; tail -3 gpurun_out/r3n_sweep.log
