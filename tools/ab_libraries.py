"""A/B timing of two builds of the same programs on one box (developer tool):

  build the libraries of two source states into  tools/ab/old  and
  tools/ab/new  (soda.cuda.build(stencil, build_dir=...)), then run this file
  on a GPU.  Used for captures r3s / r3t (profiles/README.md): box-to-box
  spread is +-3 %, so small kernel changes are compared on the same box,
  alternating, twice each.
"""
import os, sys, glob
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for sub in ('soda-compiler_b200', 'tests', 'oracle'):
  sys.path.insert(0, os.path.join(ROOT, sub))
from soda import cuda as soda_cuda
def bench(path, dims, iterate):
  lib = soda_cuda.Library(path)
  shape = tuple(reversed(dims))
  ins = [torch.rand(shape, device='cuda') for _ in lib.inputs]
  outs = [torch.empty(shape, device='cuda') for _ in lib.outputs]
  stream = torch.cuda.current_stream().cuda_stream
  for _ in range(2): lib.run_device(ins, outs, dims, 0, stream)
  torch.cuda.synchronize()
  times = []
  for _ in range(9):
    a, b = torch.cuda.Event(True), torch.cuda.Event(True)
    a.record(); lib.run_device(ins, outs, dims, 0, stream); b.record(); torch.cuda.synchronize()
    times.append(a.elapsed_time(b))
  ms = float(np.median(times))
  return ms, np.prod(dims) * iterate / ms / 1e6
for app, iterate, grids in (('heat3d', 32, [(1024,1024,1024),(1024,1024,128)]), ('jacobi3d', 32, [(1024,1024,1024),(1024,1024,128)]), ('denoise3d', 1, [(768,768,768),(768,768,96)])):
  for dims in grids:
    for rep in range(2):
      for which in ('old', 'new'):
        path = glob.glob(os.path.join(ROOT, 'tools/ab/%s/%s-*/libsoda_%s.so' % (which, app, app)))[0]
        ms, rate = bench(path, dims, iterate)
        print(app, dims, which, '%.3f ms %.1f GCell/s' % (ms, rate), flush=True)
