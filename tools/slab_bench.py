#!/usr/bin/env python3
"""Multi-GPU timing of slab-partitioned SODA programs (BASELINE configs 4, 5).

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 \
      --master-addr 127.0.0.1 --master-port 29511 tools/slab_bench.py \
      heat3d:32:1024x1024x1024:depth=2 denoise3d:16:768x768x768

A case is ``program:iterate:global dims[:option=value...]``; the GLOBAL grid is
fixed and cut into one slab per rank along the streamed dimension (strong
scaling, as BASELINE.json words configs 4 and 5).  ``weak=1`` multiplies the
streamed extent by the world size instead.  Programs whose ``iterate`` must be 1
in the reference (denoise2d/3d: reference src/soda/core.py:228-233) are applied
``iterate`` times with ``u <- output`` (``feed=input-output``, default 1-0 when
the counts differ).  Times are CUDA events on the launching stream between
barriers, maximum over ranks; rank 0 prints one JSON line per case.  Not the
contract bench (see bench.py).
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, 'soda-compiler_b200')]

import numpy as np                     # noqa: E402
import torch                           # noqa: E402
import torch.distributed as dist       # noqa: E402

from soda import core, cuda as soda_cuda, cuda_slab   # noqa: E402
from soda.codegen import cuda as codegen              # noqa: E402


def parse_case(text):
  parts = text.split(':')
  name, iterate = parts[0], int(parts[1])
  dims = tuple(int(x) for x in parts[2].split('x'))
  options = {}
  for item in parts[3:]:
    key, value = item.split('=')
    options[key] = ([int(v) for v in value.split('x')] if key == 'tile'
                    else value if key in ('style', 'feed', 'exchange')
                    else int(value))
  return name, iterate, dims, options


def stencil_of(name, iterate):
  """The program compiled for ``iterate`` iterations, or for one when the
  reference refuses to iterate it (inputs != outputs)."""
  with open(os.path.join(ROOT, 'benchmarks', name + '.soda')) as handle:
    text = handle.read()
  try:
    return core.Stencil.from_text(text, iterate=iterate), False
  except Exception:   # pylint: disable=broad-except
    return core.Stencil.from_text(text, iterate=1), True


def main():
  rank = int(os.environ.get('RANK', '0'))
  world = int(os.environ.get('WORLD_SIZE', '1'))
  local = int(os.environ.get('LOCAL_RANK', '0'))
  reps = int(os.environ.get('REPS', '5'))
  torch.cuda.set_device(local)
  if world > 1:
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
  for text in sys.argv[1:]:
    name, iterate, dims, options = parse_case(text)
    weak = options.pop('weak', 0)
    feed = options.pop('feed', None)
    exchange = options.pop('exchange', None)
    if weak:
      dims = dims[:-1] + (dims[-1] * world,)
    stencil, applied = stencil_of(name, iterate)
    library = soda_cuda.compile_stencil(stencil,
                                        options=codegen.Options(**options))
    feedback = None
    if len(library.inputs) != len(library.outputs):
      src, dst = (int(v) for v in (feed or '1-0').split('-'))
      feedback = {src: dst}
    runner = cuda_slab.SlabRunner(library, dims, rank, world,
                                  feedback=feedback,
                                  exchange=exchange)
    owned = []
    for local_in in runner.inputs:
      shape = runner.owned(local_in).shape
      owned.append(torch.rand(shape, device='cuda').to(local_in.dtype)
                   if local_in.dtype.is_floating_point else
                   torch.randint(0, 30000, shape, device='cuda').to(
                       local_in.dtype))
    runner.load_local(owned)
    del owned
    for _ in range(2):
      runner.run(iterate)
    times = []
    for _ in range(reps):
      torch.cuda.synchronize()
      if world > 1:
        dist.barrier()
      torch.cuda.synchronize()
      start, stop = torch.cuda.Event(True), torch.cuda.Event(True)
      start.record()
      runner.run(iterate)
      stop.record()
      torch.cuda.synchronize()
      ms = torch.tensor([start.elapsed_time(stop)], device='cuda')
      if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
      times.append(float(ms.item()))
    ms = float(np.median(times))
    cells = float(np.prod(dims))
    bytes_per_cell = sum(np.dtype(soda_cuda.NUMPY_TYPES[t]).itemsize
                         for _, t in library.inputs + library.outputs)
    passes = len(runner.plan(iterate))
    if rank == 0:
      print(json.dumps({
          'case': text, 'n_gpus': world, 'global_dims': list(dims),
          'iterate': iterate, 'host_level_applications': applied,
          'scaling': 'weak' if weak else 'strong', 'exchange': runner.exchange,
          'ms': round(ms, 3), 'ms_all': [round(t, 3) for t in times],
          'gcell_per_s': round(cells * iterate / ms / 1e6, 1),
          'gb_per_s_per_pass_all_gpus': round(
              cells * bytes_per_cell * passes / ms / 1e6, 1),
          'passes': passes, 'depths': runner.depths,
          'ghost_planes': [runner.reach_lo, runner.reach_hi],
          'launches_per_run_rank0': runner.launches_per_run(iterate)}),
            flush=True)
    del runner
    library.release()
    torch.cuda.empty_cache()
  if world > 1:
    dist.barrier()
    dist.destroy_process_group()


if __name__ == '__main__':
  main()
