mkdir -p gpurun_out
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port $1 tools/slab_bench.py $2 $3 $4 $5 2>/dev/null | grep '^{' ; }
(echo "# 4 GPUs, 128 planes per rank (the per-rank geometry of the 1024^3 grid on 8 GPUs)"
echo "# default faces"; run 29601 heat3d:32:1024x1024x512 heat3d:32:1024x1024x512:tile=128x32:threads=512 jacobi3d:32:1024x1024x512 jacobi3d:32:1024x1024x512:tile=128x32:threads=512
echo "# SODA_CUDA_SLAB_FACES=minimal"; SODA_CUDA_SLAB_FACES=minimal run 29602 heat3d:32:1024x1024x512 heat3d:32:1024x1024x512:tile=128x32:threads=512
echo "# SODA_CUDA_SLAB_FACES=chunk"; SODA_CUDA_SLAB_FACES=chunk run 29603 heat3d:32:1024x1024x512 heat3d:32:1024x1024x512:tile=128x32:threads=512
echo "# SODA_CUDA_CHUNKS=4"; SODA_CUDA_CHUNKS=4 run 29604 heat3d:32:1024x1024x512
echo "# SODA_CUDA_CHUNKS=3"; SODA_CUDA_CHUNKS=3 run 29605 heat3d:32:1024x1024x512
) > gpurun_out/r3a_slab_thin.log 2>&1
python - <<'PY'
import json
for l in open('gpurun_out/r3a_slab_thin.log'):
  if l.startswith('#'): print(l.strip())
  elif l.startswith('{'):
    d = json.loads(l); print('  ', d['case'], d['ms'], d['gcell_per_s'], d['launches_per_run_rank0'])
PY
