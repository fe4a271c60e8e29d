# two GPUs: the contract bench as the driver launches it, the slab and sharded
# tests on real peers
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2f_bench_n2.json 2> gpurun_out/r2f_bench_n2.err; tail -c 1500 gpurun_out/r2f_bench_n2.err
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/r2f_bench_n2.json") if l.startswith("{")][-1])
print('value', d['value'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['per_slab'], d['e2e']['copy_ceiling']['ms'])
print('parity', d.get('multi_gpu_parity'))
for e in d['extra']:
  print(e.get('workload'), e.get('build'), e.get('value'), e.get('ms'), (e.get('roofline') or {}).get('frac'), e.get('error'))
PY
timeout 900 python -m pytest tests/test_slab_gpu.py tests/test_sharded_run_gpu.py -q > gpurun_out/r2f_pytest_multi.log 2>&1; tail -5 gpurun_out/r2f_pytest_multi.log
