mkdir -p gpurun_out
S=$(date +%s); python bench.py > gpurun_out/r3r_bench_default.json 2> gpurun_out/r3r_bench_default.err; echo "default bench wall $(( $(date +%s) - S )) s"
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/r3r_bench_default.json") if l.startswith("{")][-1])
print({k:d[k] for k in ("metric","value","unit","n_gpus","steps","warmup","ms_per_step","gpu_launches","vs_baseline","dtype")})
print(d["roofline"]); print(d["cpu_baseline"]); print(d["e2e"]["value"], d["clocks"])
PY
S=$(date +%s); python bench.py --impl reference > gpurun_out/r3r_bench_reference.json 2> gpurun_out/r3r_ref.err; echo "reference arm wall $(( $(date +%s) - S )) s"; tail -c 700 gpurun_out/r3r_bench_reference.json
