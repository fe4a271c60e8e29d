"""Multi-GPU slabs on real devices: NCCL halo exchange, bit-identical result.

Needs >= 2 GPUs (``gpurun --gpus 2``); skipped on a single-GPU box.
"""
import os
import socket

import numpy as np
import pytest
import torch

import common

pytestmark = pytest.mark.gpu


def _free_port():
  with socket.socket() as sock:
    sock.bind(('127.0.0.1', 0))
    return sock.getsockname()[1]


# program, iterate, backend options, global dims[, exchange]; the default
# exchange on GPUs is `p2p` (copy engine into peer-mapped ghost rows + flags)
CASES = [
    ('jacobi2d', 16, {'depth': 4}, (2048, 700)),
    ('jacobi2d', 7, {'depth': 4}, (1030, 333)),
    ('jacobi2d', 40, {'depth': 8}, (2048, 600)),      # 5 launches, 3 buffers
    ('heat3d', 4, {'depth': 2}, (128, 64, 90)),
    ('seidel2d', 6, {'depth': 2}, (1280, 400)),
    ('blur', 1, {}, (1037, 211)),
    ('denoise2d', 1, {}, (1024, 200)),
    ('jacobi2d', 16, {'depth': 4}, (2048, 700), 'collective'),   # NCCL
    ('heat3d', 4, {'depth': 2}, (128, 64, 90), 'collective'),
]


def _worker(rank, world, port, queue):
  """One process per GPU for ALL cases: process start-up, the first CUDA
  context and the NCCL rendezvous dominate the cost of this test."""
  import torch.distributed as dist
  os.environ['MASTER_ADDR'] = '127.0.0.1'
  os.environ['MASTER_PORT'] = str(port)
  torch.cuda.set_device(rank)
  dist.init_process_group('nccl', rank=rank, world_size=world,
                          device_id=torch.device('cuda', rank))
  try:
    from soda import cuda as soda_cuda, cuda_slab
    from soda.codegen import cuda as codegen
    for index, (name, iterate, options, dims, *how) in enumerate(CASES):
      library = soda_cuda.compile_stencil(common.stencil(name, iterate),
                                          options=codegen.Options(**options))
      orc = common.oracle(name, iterate)
      full = common.random_inputs(orc, dims, seed=23)
      runner = cuda_slab.SlabRunner(library, dims, rank, world,
                                    exchange=how[0] if how else None)
      assert runner.exchange == (how[0] if how else 'p2p')
      runner.load_local([torch.from_numpy(a[runner.begin:runner.end].copy()
                                          ).cuda() for a in full])
      for _ in range(3):          # back-to-back runs, no barrier in between
        outs = runner.run(iterate)
      torch.cuda.synchronize()
      queue.put((index, rank, runner.begin, runner.end,
                 [o.cpu().numpy().copy() for o in outs]))
      dist.barrier()
  finally:
    dist.destroy_process_group()


def test_sharded_equals_oracle():
  world = min(torch.cuda.device_count(), 4)
  if world < 2:
    pytest.skip('needs at least 2 GPUs')
  import torch.multiprocessing as mp
  from soda import cuda as soda_cuda
  from soda.codegen import cuda as codegen
  for name, iterate, options, dims, *_ in CASES:
    soda_cuda.build(common.stencil(name, iterate),
                    options=codegen.Options(**options))
  ctx = mp.get_context('spawn')
  queue = ctx.Queue()
  port = _free_port()
  procs = [ctx.Process(target=_worker, args=(rank, world, port, queue))
           for rank in range(world)]
  for proc in procs:
    proc.start()
  pieces = [queue.get(timeout=600) for _ in range(world * len(CASES))]
  for proc in procs:
    proc.join(timeout=60)
    assert proc.exitcode == 0
  for index, (name, iterate, options, dims, *_) in enumerate(CASES):
    orc = common.oracle(name, iterate)
    want = orc.run(common.random_inputs(orc, dims, seed=23))
    for k, expected in enumerate(want):
      got = np.zeros_like(expected)
      for case, _, begin, end, outs in pieces:
        if case == index:
          got[begin:end] = outs[k]
      common.assert_bit_exact(got, expected,
                              '%s on %d GPUs' % (name, world))


@pytest.mark.parametrize('name,times,dims', [
    ('denoise3d', 4, (96, 40, 33)),
    ('denoise2d', 3, (1024, 70)),
])
def test_repeated_application_one_gpu(name, times, dims):
  """BASELINE config 5 ("denoise3d iterate 16" is outside the reference,
  core.py:228-233): the program applied `times` times with u <- output equals
  the golden loop called `times` times, on the region every call keeps."""
  from soda import cuda as soda_cuda, cuda_slab
  library = soda_cuda.compile_stencil(common.stencil(name, 1))
  orc = common.oracle(name, 1)
  f, u = common.random_inputs(orc, dims, seed=31)
  runner = cuda_slab.SlabRunner(library, dims, 0, 1, feedback={1: 0})
  runner.load_local([torch.from_numpy(f).cuda(), torch.from_numpy(u).cuda()])
  got = runner.run(times)[0].cpu().numpy()
  for _ in range(times):
    u, = orc.run([f, u])
  lo, hi = library.window(1)
  keep = np.zeros(u.shape, dtype=bool)
  keep[tuple(slice(-l * times, n - h * times) for l, h, n in
             reversed(list(zip(lo, hi, dims))))] = True
  assert keep.sum() > 0
  # denoise2d's lowered form multiplies where the DSL text divides (SURVEY
  # 8a2): on noise it overflows to inf and then NaN within three applications
  common.assert_bit_exact(got, np.where(keep, u, 0).astype(u.dtype),
                          '%s x%d' % (name, times), any_nan=True)
