#!/usr/bin/env python3
"""Host cost of a device-resident run on small grids (developer tool): wall
time per `soda_cuda_run_device` call against the kernels' own time, for the
launch-bound configurations (BASELINE config 1: blur 2000 x 1000).

  python tools/launch_overhead.py
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, 'soda-compiler_b200')]

import numpy as np   # noqa: E402
import torch         # noqa: E402

from soda import core, cuda as soda_cuda   # noqa: E402


def main():
  for name, iterate, dims in (('blur', 1, (2000, 1000)),
                              ('jacobi2d', 64, (2000, 1000)),
                              ('heat3d', 32, (128, 128, 128))):
    stencil = core.Stencil.from_file(
        os.path.join(ROOT, 'benchmarks', name + '.soda'), iterate=iterate)
    library = soda_cuda.compile_stencil(stencil)
    shape = tuple(reversed(dims))
    dtype = {2: torch.int16, 4: torch.float32}[np.dtype(
        soda_cuda.NUMPY_TYPES[library.inputs[0][1]]).itemsize]
    ins = [torch.ones(shape, dtype=dtype, device='cuda')]
    outs = [torch.empty(shape, dtype=dtype, device='cuda')]
    stream = torch.cuda.current_stream().cuda_stream
    for _ in range(20):
      library.run_device(ins, outs, dims, 0, stream)
    torch.cuda.synchronize()
    calls = 300
    start, stop = torch.cuda.Event(True), torch.cuda.Event(True)
    t0 = time.perf_counter()
    start.record()
    for _ in range(calls):
      library.run_device(ins, outs, dims, 0, stream)
    stop.record()
    issue_us = (time.perf_counter() - t0) * 1e6 / calls
    torch.cuda.synchronize()
    wall_us = (time.perf_counter() - t0) * 1e6 / calls
    device_us = start.elapsed_time(stop) * 1e3 / calls
    launches = library.stats['launches']
    print('%-9s x%-2d %-14s %2d launch(es) per run: host issues a run in '
          '%6.1f us (%4.1f us per launch), device time %6.1f us per run, '
          'wall %6.1f us' % (name, iterate, 'x'.join(map(str, dims)), launches,
                             issue_us, issue_us / launches, device_us,
                             wall_us), flush=True)


if __name__ == '__main__':
  main()
