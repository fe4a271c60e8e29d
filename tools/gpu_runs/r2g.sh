# eight GPUs: the contract bench as the driver launches it
mkdir -p gpurun_out
N=${1:-8}
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2g_bench_n$N.json 2> gpurun_out/r2g_bench_n$N.err ) 2>&1 | tail -3
grep -v "Warning\|^\*\|OMP_NUM\|^$" gpurun_out/r2g_bench_n$N.err | tail -5
python - $N <<'PY'
import json, sys
n = sys.argv[1]
lines = [l for l in open('gpurun_out/r2g_bench_n%s.json' % n) if l.startswith('{')]
d = json.loads(lines[-1])
print('value', d['value'], 'ms', d['ms_per_step'], 'kernel', d['kernel'])
e = d['e2e']
print('e2e', e['value'], e['ms_per_step'], 'ceiling', e['copy_ceiling']['ms'], e['copy_ceiling']['gb_per_s_each_direction'], e['copy_ceiling']['e2e_fraction_of_ceiling'])
print('per slab h2d', [round(s['h2d_ms'], 1) for s in e['per_slab']])
print('parity', d.get('multi_gpu_parity'))
for x in d['extra']:
  print(x.get('workload'), '|', x.get('build', '')[:5], x.get('value'), x.get('ms'), (x.get('roofline') or {}).get('frac'), x.get('error'))
PY
nvidia-smi topo -m | head -12; free -g | head -2; nproc
