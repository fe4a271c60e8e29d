# four GPUs: bench as the driver launches it + smoke + reference arm under torchrun
mkdir -p gpurun_out
python -c "import __graft_entry__ as e; e.smoke()" 2>&1 | tail -2
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus 4 --steps 10 --warmup 3 > gpurun_out/r2q_bench_n4.json 2> gpurun_out/r2q_bench_n4.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29556 bench.py --impl reference --gpus 4 --steps 2 --warmup 1 > gpurun_out/r2q_bench_n4_reference.json 2>> gpurun_out/r2q_bench_n4.err
grep -v "Warning\|^\*\|OMP_NUM\|^$" gpurun_out/r2q_bench_n4.err | tail -5
python - <<'PY'
import json
lines = [l for l in open('gpurun_out/r2q_bench_n4.json') if l.startswith('{')]
d = json.loads(lines[-1])
print('value', d['value'], 'ms', d['ms_per_step'], d['kernel'])
e = d['e2e']
print('e2e', e['value'], e['ms_per_step'], 'ceiling', e['copy_ceiling']['ms'], e['copy_ceiling']['gb_per_s_each_direction'])
print('parity', d.get('multi_gpu_parity'))
for x in d['extra']:
  print(x.get('workload'), '|', x.get('build', '')[:5], x.get('value'), x.get('ms'), (x.get('roofline') or {}).get('frac'), x.get('error'))
lines = [l for l in open('gpurun_out/r2q_bench_n4_reference.json') if l.startswith('{')]
print(len(lines), lines[-1][:200])
PY
