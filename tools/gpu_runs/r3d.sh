mkdir -p gpurun_out
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port $1 tools/slab_bench.py $2 $3 $4 $5 2>/dev/null | grep '^{' ; }
(for c in 2 3 4 5 6 8; do echo "# SODA_CUDA_CHUNKS=$c"; SODA_CUDA_CHUNKS=$c REPS=3 run 2963$c heat3d:32:1024x1024x512 jacobi3d:32:1024x1024x512; done
for c in 4 6; do echo "# SODA_CUDA_SLAB_FACES=minimal SODA_CUDA_CHUNKS=$c"; SODA_CUDA_SLAB_FACES=minimal SODA_CUDA_CHUNKS=$c REPS=3 run 2964$c heat3d:32:1024x1024x512; done) > gpurun_out/r3d_slab_chunks.log 2>&1
python - <<'PY'
import json
for l in open('gpurun_out/r3d_slab_chunks.log'):
  if l.startswith('#'): print(l.strip())
  elif l.startswith('{'):
    d = json.loads(l); print('  ', d['case'], d['ms'], d['gcell_per_s'])
PY
