mkdir -p gpurun_out
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port $1 tools/slab_bench.py $2 $3 $4 $5 2>/dev/null | grep '^{' ; }
(echo "# default (blocks of at most a quarter slab)"; REPS=3 run 29651 heat3d:32:1024x1024x512 jacobi3d:32:1024x1024x512 denoise3d:16:768x768x384; echo "# SODA_CUDA_CHUNKS=4"; SODA_CUDA_CHUNKS=4 REPS=3 run 29652 heat3d:32:1024x1024x512 jacobi3d:32:1024x1024x512) > gpurun_out/r3e_slab_thin.log 2>&1
timeout 600 python -m pytest tests/test_slab_gpu.py -q 2>&1 | tail -2
python - <<'PY'
import json
for l in open('gpurun_out/r3e_slab_thin.log'):
  if l.startswith('#'): print(l.strip())
  elif l.startswith('{'):
    d = json.loads(l); print('  ', d['case'], d['ms'], d['gcell_per_s'])
PY
