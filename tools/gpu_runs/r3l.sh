mkdir -p gpurun_out
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port $1 tools/slab_bench.py $2 $3 $4 $5 2>/dev/null | grep '^{' ; }
(echo "# default"; REPS=3 run 29651 denoise3d:16:768x768x384
 echo "# MIN_BLOCKS=3 CHUNKS=3"; SODA_CUDA_SLAB_MIN_BLOCKS=3 SODA_CUDA_CHUNKS=3 REPS=3 run 29652 denoise3d:16:768x768x384
 echo "# MIN_BLOCKS=2 CHUNKS=2"; SODA_CUDA_SLAB_MIN_BLOCKS=2 SODA_CUDA_CHUNKS=2 REPS=3 run 29653 denoise3d:16:768x768x384
 echo "# MIN_BLOCKS=3 (library's choice, at least 3)"; SODA_CUDA_SLAB_MIN_BLOCKS=3 REPS=3 run 29654 denoise3d:16:768x768x384 heat3d:32:1024x1024x512
 echo "# FACES=minimal CHUNKS=3"; SODA_CUDA_SLAB_FACES=minimal SODA_CUDA_CHUNKS=3 REPS=3 run 29655 denoise3d:16:768x768x384
 echo "# CHUNKS=6"; SODA_CUDA_CHUNKS=6 REPS=3 run 29656 denoise3d:16:768x768x384
) > gpurun_out/r3l_slab_denoise3d.log 2>&1
python - <<'PY'
import json
for l in open('gpurun_out/r3l_slab_denoise3d.log'):
  if l.startswith('#'): print(l.strip())
  elif l.startswith('{'):
    d = json.loads(l); print('  ', d['case'], d['ms'], d['gcell_per_s'])
PY
