mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
CASES="heat3d:32:1024x1024x1024 jacobi3d:32:1024x1024x1024 denoise3d:16:768x768x768"
timeout 300 $TR --nproc-per-node 8 --master-port 29541 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r1s_bench_n8.json 2> gpurun_out/r1s_bench_n8.err
timeout 300 $TR --nproc-per-node 8 --master-port 29542 tools/slab_bench.py $CASES > gpurun_out/r1s_slab_n8.log 2>&1
grep '^{' gpurun_out/r1s_slab_n8.log | cut -c1-330; cut -c1-220 gpurun_out/r1s_bench_n8.json; tail -3 gpurun_out/r1s_bench_n8.err
