mkdir -p gpurun_out
timeout 600 python bench.py > gpurun_out/r1j_bench.json 2> gpurun_out/r1j_bench.err; tail -c 400 gpurun_out/r1j_bench.json; tail -3 gpurun_out/r1j_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r1j_bench_reference.json 2>&1; tail -c 300 gpurun_out/r1j_bench_reference.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1j_launches_bench_jacobi2d_d8.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r1j_launches.log 2>&1
timeout 300 bash tools/ncu_capture.sh r1j_j2d_d8 soda_ jacobi2d:64:16384x16384
for c in blur:1:32768x32768 sobel2d:1:32768x32768 seidel2d:2:16384x16384; do n=${c%%:*}; timeout 300 bash tools/ncu_capture.sh r1j_$n soda_ $c; done
timeout 300 python tools/quick_bench.py blur:1:2000x1000 blur:1:32768x32768 sobel2d:1:32768x32768 denoise2d:1:32768x32768 jacobi2d:64:16384x16384 jacobi2d:64:16384x16384:depth=1 seidel2d:2:16384x16384 heat3d:32:1024x1024x1024 jacobi3d:32:1024x1024x1024 denoise3d:1:768x768x768 heat3d:32:1024x1024x1024:depth=4 heat3d:32:1024x1024x1024:depth=3 > gpurun_out/r1j_sweep.log 2>&1; cat gpurun_out/r1j_sweep.log
ls gpurun_out
