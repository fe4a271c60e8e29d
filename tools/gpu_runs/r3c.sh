mkdir -p gpurun_out
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port $1 tools/slab_bench.py $2 $3 $4 $5 2>/dev/null | grep '^{' ; }
(echo "# default (quarter-block faces)"; run 29621 heat3d:32:1024x1024x512; echo "# SODA_CUDA_SLAB_FACES=minimal"; SODA_CUDA_SLAB_FACES=minimal run 29622 heat3d:32:1024x1024x512; echo "# SODA_CUDA_CHUNKS=4"; SODA_CUDA_CHUNKS=4 run 29623 heat3d:32:1024x1024x512; echo "# default again"; run 29624 heat3d:32:1024x1024x512) > gpurun_out/r3c_slab_thin.log 2>&1
python - <<'PY'
import json
for l in open('gpurun_out/r3c_slab_thin.log'):
  if l.startswith('#'): print(l.strip())
  elif l.startswith('{'):
    d = json.loads(l); print('  ', d['case'], d['ms'], d['ms_all'], d['gcell_per_s'])
PY
