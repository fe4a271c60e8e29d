mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
for mode in minimal chunk minimal chunk; do
  SODA_CUDA_SLAB_FACES=$mode timeout 300 $TR --nproc-per-node 2 --master-port 29541 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r1p_bench_n2_$mode.json 2> gpurun_out/r1p_bench_n2.err
  echo $mode; python -c "
import json,sys
d=json.loads(open('gpurun_out/r1p_bench_n2_$mode.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'])"
done
for mode in minimal chunk; do
SODA_CUDA_SLAB_FACES=$mode timeout 300 $TR --nproc-per-node 2 --master-port 29542 tools/slab_bench.py heat3d:32:1024x1024x512 jacobi2d:64:16384x16384:weak=1 2>&1 | grep '^{' | cut -c1-300 | tee gpurun_out/r1p_slab_n2_$mode.log
done
python bench.py --steps 10 --warmup 3 --no-cpu-baseline | cut -c1-200
