mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r1k_pytest_n2.log 2>&1; tail -4 gpurun_out/r1k_pytest_n2.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r1k_bench_n2.json 2> gpurun_out/r1k_bench_n2.err; tail -c 1500 gpurun_out/r1k_bench_n2.json; tail -3 gpurun_out/r1k_bench_n2.err
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r1k_bench_n1.json 2>/dev/null; head -c 300 gpurun_out/r1k_bench_n1.json
