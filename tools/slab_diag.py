import os, sys, time
ROOT='/root/repo'
sys.path[:0]=[ROOT+'/soda-compiler_b200']
import torch, torch.distributed as dist
from soda import core, cuda as soda_cuda, cuda_slab
rank=int(os.environ['RANK']); world=int(os.environ['WORLD_SIZE'])
torch.cuda.set_device(rank)
dist.init_process_group('nccl', device_id=torch.device('cuda', rank))
st = core.Stencil.from_file(ROOT+'/benchmarks/jacobi2d.soda', iterate=64)
lib = soda_cuda.compile_stencil(st)
dims=(16384,16384)
x = torch.rand((dims[1],dims[0]), device='cuda')
def bench(tag, runner, n=10):
  for _ in range(2): runner.run(64)
  torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
  t0=time.perf_counter()
  e0=torch.cuda.Event(True); e1=torch.cuda.Event(True); e0.record()
  for _ in range(n): runner.run(64)
  e1.record(); t1=time.perf_counter()
  torch.cuda.synchronize(); t2=time.perf_counter()
  print('rank %d '%rank + '%-28s host issue %.2f ms/run, gpu %.2f ms/run, wall %.2f' % (tag,(t1-t0)/n*1e3, e0.elapsed_time(e1)/n, (t2-t0)/n*1e3), flush=True)
  dist.barrier()
r = cuda_slab.SlabRunner(lib, (dims[0], dims[1]*world), rank, world, exchange='collective'); r.load_local([x]); bench('collective', r)
r = cuda_slab.SlabRunner(lib, (dims[0], dims[1]*world), rank, world, exchange='p2p'); r.load_local([x])
if rank==0: print('flag_by_copy', r.flag_by_copy, flush=True)
bench('p2p', r)
orig_push = r._push_faces; orig_await = r._await_ghosts
def push_nothing(target, bset, first):
  r.tick += 1
r._push_faces = push_nothing; r._await_ghosts = lambda: None
bench('no exchange at all', r)
dist.destroy_process_group()
