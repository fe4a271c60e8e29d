#!/usr/bin/env python3
"""Developer micro-benchmark: device-resident timing of compiled SODA programs.

  python tools/quick_bench.py jacobi2d:64:16384x16384:depth=8 heat3d:4:512x512x512

Prints one line per case: GCell/s (cell updates / s), GB/s of algorithmic HBM
traffic per launch pass, kernel config.  Not the contract bench (see bench.py).
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, 'soda-compiler_b200')]

import numpy as np   # noqa: E402
import torch         # noqa: E402

from soda import core, cuda as soda_cuda                # noqa: E402
from soda.codegen import cuda as codegen                # noqa: E402


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


def time_e2e(library, dims, reps):
  """ms per call of library.run on pinned host arrays (H2D + run + D2H)."""
  import time
  shape = tuple(reversed(dims))
  pin = lambda t: torch.empty(shape, dtype=torch.from_numpy(np.empty(
      0, soda_cuda.NUMPY_TYPES[t])).dtype, pin_memory=True)
  ins = [pin(t) for _, t in library.inputs]
  for x in ins:
    x.copy_(torch.rand(shape) if x.dtype.is_floating_point else
            torch.randint(0, 30000, shape).to(x.dtype))
  outs = [pin(t) for _, t in library.outputs]
  np_in, np_out = [x.numpy() for x in ins], [x.numpy() for x in outs]
  for _ in range(2):
    library.run(np_in, np_out)
  times = []
  for _ in range(reps):
    t0 = time.perf_counter()
    library.run(np_in, np_out)
    times.append((time.perf_counter() - t0) * 1e3)
  return float(np.median(times)), library.stats


def _unused():
  name = iterate = dims = options = None
  return name, iterate, dims, options


def main():
  reps = int(os.environ.get('REPS', '7'))
  for text in sys.argv[1:]:
    name, iterate, dims, options = parse_case(text)
    e2e = options.pop('e2e', 0)
    fast = bool(options.pop('fast', 0))   # -DSODA_CUDA_FAST_MATH build
    with open(os.path.join(ROOT, 'benchmarks', name + '.soda')) as handle:
      stencil = core.Stencil.from_text(handle.read(), iterate=iterate)
    try:
      library = soda_cuda.compile_stencil(stencil, fast_math=fast,
                                          options=codegen.Options(**options))
    except Exception as e:   # pylint: disable=broad-except
      print('%-50s build failed: %s' % (text, str(e)[:300]))
      continue
    if e2e:
      ms, stats = time_e2e(library, dims, reps)
      print('%-50s e2e %8.3f ms  %8.1f GCell/s  pieces %s h2d %.2f run %.2f '
            'd2h %.2f launches %d' % (
                text, ms, float(np.prod(dims)) * iterate / ms / 1e6,
                os.environ.get('SODA_CUDA_PIECES', 'auto'), stats['h2d_ms'],
                stats['kernel_ms'], stats['d2h_ms'], stats['launches']),
            flush=True)
      library.release()
      continue
    shape = tuple(reversed(dims))
    dtype = {1: torch.uint8, 2: torch.int16, 4: torch.float32,
             8: torch.float64}
    ins = []
    for _, t in library.inputs:
      size = np.dtype(soda_cuda.NUMPY_TYPES[t]).itemsize
      x = torch.rand(shape, device='cuda') if size == 4 else \
          torch.randint(0, 30000, shape, device='cuda').to(dtype[size])
      ins.append(x.to(dtype[size]))
    outs = [torch.empty(shape, dtype=dtype[np.dtype(
        soda_cuda.NUMPY_TYPES[t]).itemsize], device='cuda')
            for _, t in library.outputs]
    stream = torch.cuda.current_stream().cuda_stream
    try:
      for _ in range(2):
        library.run_device(ins, outs, dims, 0, stream)
      torch.cuda.synchronize()
      times = []
      for _ in range(reps):
        start, stop = torch.cuda.Event(True), torch.cuda.Event(True)
        start.record()
        library.run_device(ins, outs, dims, 0, stream)
        stop.record()
        torch.cuda.synchronize()
        times.append(start.elapsed_time(stop))
    except Exception as e:   # pylint: disable=broad-except
      print('%-50s run failed: %s' % (text, e))
      continue
    ms = float(np.median(times))
    cells = float(np.prod(dims))
    stats = library.stats
    bytes_per_cell = sum(
        np.dtype(soda_cuda.NUMPY_TYPES[t]).itemsize
        for _, t in library.inputs + library.outputs)
    passes = stats['launches']
    print('%-50s %8.3f ms  %8.1f GCell/s  %7.1f GB/s/pass  launches %d depth '
          '%d blocks %d thr %d smem %d tma %d' % (
              text, ms, cells * iterate / ms / 1e6,
              cells * bytes_per_cell * passes / ms / 1e6, passes,
              stats['depth'], stats['blocks'], stats['threads'],
              stats['smem_bytes'], stats['used_tma']), flush=True)
    del ins, outs
    library.release()
    torch.cuda.empty_cache()


if __name__ == '__main__':
  main()
