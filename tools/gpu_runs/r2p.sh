mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2p_pytest_gpu.log 2>&1; tail -4 gpurun_out/r2p_pytest_gpu.log
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2p_bench.json 2> gpurun_out/r2p_bench.err; tail -3 gpurun_out/r2p_bench.err
python - <<'PY'
import json
lines = [l for l in open('gpurun_out/r2p_bench.json') if l.startswith('{')]
d = json.loads(lines[-1])
print('value', d['value'], 'e2e', d['e2e']['value'])
for x in d['extra']:
  print(x.get('workload'), '|', x.get('build', '')[:5], x.get('value'), x.get('ms'), (x.get('roofline') or {}).get('frac'), x.get('error'))
PY
