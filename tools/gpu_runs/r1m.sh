set -x
mkdir -p gpurun_out
nvidia-smi -L | head -8 > gpurun_out/r1m_gpus.log
timeout 400 python -m pytest tests/test_slab_gpu.py -x -q > gpurun_out/r1m_pytest_slab_n4.log 2>&1
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
CASES="heat3d:32:1024x1024x1024 jacobi3d:32:1024x1024x1024 denoise3d:16:768x768x768"
timeout 300 $TR --nproc-per-node 4 --master-port 29541 bench.py --gpus 4 --steps 10 --warmup 3 > gpurun_out/r1m_bench_n4.json 2> gpurun_out/r1m_bench_n4.err
timeout 300 $TR --nproc-per-node 4 --master-port 29542 tools/slab_bench.py $CASES > gpurun_out/r1m_slab_n4.log 2>&1
timeout 300 $TR --nproc-per-node 2 --master-port 29543 tools/slab_bench.py $CASES > gpurun_out/r1m_slab_n2.log 2>&1
timeout 300 python tools/slab_bench.py $CASES > gpurun_out/r1m_slab_n1.log 2>&1
tail -3 gpurun_out/r1m_pytest_slab_n4.log; cat gpurun_out/r1m_slab_n*.log | grep '^{' | cut -c1-400; cut -c1-200 gpurun_out/r1m_bench_n4.json
