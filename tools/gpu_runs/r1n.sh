mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_slab_gpu.py -x -q > gpurun_out/r1n_pytest_slab_n4.log 2>&1
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
CASES="heat3d:32:1024x1024x1024 jacobi3d:32:1024x1024x1024 denoise3d:16:768x768x768"
timeout 300 $TR --nproc-per-node 4 --master-port 29542 tools/slab_bench.py $CASES > gpurun_out/r1n_slab_n4.log 2>&1
timeout 300 $TR --nproc-per-node 4 --master-port 29541 bench.py --gpus 4 --steps 10 --warmup 3 > gpurun_out/r1n_bench_n4.json 2> gpurun_out/r1n_bench_n4.err
tail -3 gpurun_out/r1n_pytest_slab_n4.log; cat gpurun_out/r1n_slab_n*.log | grep '^{' | cut -c1-330; cut -c1-200 gpurun_out/r1n_bench_n4.json
