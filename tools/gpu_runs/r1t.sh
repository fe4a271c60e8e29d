mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 300 $TR --nproc-per-node 2 --master-port 29541 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r1t_bench_n2.json 2> gpurun_out/r1t_bench_n2.err
python -c "
import json
d=json.loads(open('gpurun_out/r1t_bench_n2.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e'])"
tail -2 gpurun_out/r1t_bench_n2.err
for i in 0 1; do cat /sys/bus/pci/devices/$(nvidia-smi --query-gpu=pci.bus_id --format=csv,noheader -i $i | cut -c5- | tr A-Z a-z)/numa_node; done; nproc; lscpu | grep -i numa
