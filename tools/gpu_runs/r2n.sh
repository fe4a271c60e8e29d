mkdir -p gpurun_out
python bench.py --steps 10 --warmup 3 > gpurun_out/r2n_bench.json 2> gpurun_out/r2n_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2n_bench_reference.json 2>> gpurun_out/r2n_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2n_launches_bench_jacobi2d_d8.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/r2n_launches.log 2>&1
bash tools/ncu_capture.sh r2n_j2d_d8 soda_jacobi2d jacobi2d:64:16384x16384
bash tools/ncu_capture.sh r2n_heat3d_d2 soda_heat3d heat3d:32:1024x1024x1024
bash tools/ncu_capture.sh r2n_sobel2d soda_sobel2d sobel2d:1:32768x32768
cut -c1-400 gpurun_out/r2n_bench.json; echo; cut -c1-300 gpurun_out/r2n_bench_reference.json; tail -3 gpurun_out/r2n_bench.err
