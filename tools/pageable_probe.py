#!/usr/bin/env python3
"""Developer probe: what does pageable host memory cost a run on host buffers,
and what would pinning it for the duration of the call cost?

  python tools/pageable_probe.py
"""
import ctypes
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, 'soda-compiler_b200')]

import numpy as np   # noqa: E402
import torch         # noqa: E402

from soda import core, cuda as soda_cuda   # noqa: E402


def main():
  stencil = core.Stencil.from_file(
      os.path.join(ROOT, 'benchmarks', 'jacobi2d.soda'), iterate=64)
  library = soda_cuda.compile_stencil(stencil)
  shape = (16384, 16384)
  pageable_in = np.random.default_rng(0).random(shape, dtype=np.float32)
  pageable_out = np.empty(shape, dtype=np.float32)
  pinned_in = torch.empty(shape, dtype=torch.float32, pin_memory=True)
  pinned_out = torch.empty(shape, dtype=torch.float32, pin_memory=True)
  pinned_in.copy_(torch.from_numpy(pageable_in))
  for what, a, b in (('pinned', pinned_in.numpy(), pinned_out.numpy()),
                     ('pageable', pageable_in, pageable_out)):
    for _ in range(2):
      library.run([a], [b])
    times = []
    for _ in range(5):
      t0 = time.perf_counter()
      library.run([a], [b])
      times.append((time.perf_counter() - t0) * 1e3)
    print('%-9s host buffers: %.1f ms per run (h2d %.1f, d2h %.1f)' % (
        what, float(np.median(times)), library.stats['h2d_ms'],
        library.stats['d2h_ms']), flush=True)
  assert np.array_equal(pageable_out, pinned_out.numpy())
  cudart = ctypes.CDLL('libcudart.so.12')
  for _ in range(3):
    t0 = time.perf_counter()
    rc1 = cudart.cudaHostRegister(ctypes.c_void_p(pageable_in.ctypes.data),
                                  ctypes.c_size_t(pageable_in.nbytes), 0)
    rc2 = cudart.cudaHostRegister(ctypes.c_void_p(pageable_out.ctypes.data),
                                  ctypes.c_size_t(pageable_out.nbytes), 0)
    t1 = time.perf_counter()
    library.run([pageable_in], [pageable_out])
    t2 = time.perf_counter()
    cudart.cudaHostUnregister(ctypes.c_void_p(pageable_in.ctypes.data))
    cudart.cudaHostUnregister(ctypes.c_void_p(pageable_out.ctypes.data))
    t3 = time.perf_counter()
    print('register 2 x 1 GiB: %.1f ms (rc %d %d), run %.1f ms, unregister '
          '%.1f ms' % ((t1 - t0) * 1e3, rc1, rc2, (t2 - t1) * 1e3,
                       (t3 - t2) * 1e3), flush=True)


if __name__ == '__main__':
  main()
