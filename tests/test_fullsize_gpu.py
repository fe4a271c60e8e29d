"""Parity at BASELINE.json's full sizes (needs a GPU): the crop property.

A stencil result depends only on the cell's transitive window (reference
src/soda/core.py:793-830), so for ANY box B of the grid

    oracle(inputs restricted to B) == cuda(full inputs) restricted to B

on the cells of B whose window stays inside B — and, where B touches the edge
of the grid, on the border there too (both sides store 0 outside the valid
region).  The CUDA path runs the whole configuration on device-resident
arrays through the C ABI; the CPU oracle only sees boxes it finishes in
seconds: corners (grid-edge handling, first/last tile and chunk), boxes
straddling tile and chunk boundaries in the middle, and the far corner (the
last partial tile).  Every comparison is bitwise.
"""
import numpy as np
import pytest
import torch

import common
from soda import cuda as soda_cuda

pytestmark = pytest.mark.gpu


def _device_inputs(library, dims, seed):
  """Uniform noise generated on the device (the reference initialiser's ramp
  is smooth, which hides misplaced neighbours)."""
  gen = torch.Generator(device='cuda')
  gen.manual_seed(seed)
  shape = tuple(reversed(dims))
  arrays = []
  for _, haoda_type in library.inputs:
    dtype = np.dtype(soda_cuda.NUMPY_TYPES[haoda_type])
    if dtype.kind == 'f':
      arrays.append(torch.rand(shape, generator=gen, device='cuda',
                               dtype=torch.float32).to(
                                   torch.from_numpy(np.empty(0, dtype)).dtype))
    else:
      signed = {1: torch.int8, 2: torch.int16, 4: torch.int32,
                8: torch.int64}[dtype.itemsize]
      info = torch.iinfo(signed)
      arrays.append(torch.randint(info.min, info.max, shape, generator=gen,
                                  device='cuda', dtype=signed))
  return arrays


def _host(tensor, haoda_type):
  array = tensor.contiguous().cpu().numpy()
  return array.view(soda_cuda.NUMPY_TYPES[haoda_type])


def _boxes(dims, size):
  """Corners, the far corner, and boxes at odd offsets in the middle."""
  size = [min(s, n) for s, n in zip(size, dims)]
  low = [0] * len(dims)
  high = [n - s for n, s in zip(dims, size)]
  mid = [(n - s) // 2 + 37 for n, s in zip(dims, size)]
  third = [(n - s) // 3 - 11 for n, s in zip(dims, size)]
  mixed = [h if d % 2 else 0 for d, h in enumerate(high)]
  starts = [low, high, mixed,
            [max(0, min(m, h)) for m, h in zip(mid, high)],
            [max(0, min(t, h)) for t, h in zip(third, high)]]
  return [tuple((s, s + e) for s, e in zip(start, size)) for start in starts]


def _check_boxes(name, dims, library, lo, hi, inputs, outputs, reference,
                 boxes, extend_to_grid_edge=True):
  for box in boxes:
    slices = tuple(slice(b, e) for b, e in reversed(box))
    sub_in = [_host(t[slices], haoda_type)
              for t, (_, haoda_type) in zip(inputs, library.inputs)]
    want = reference(sub_in)
    region = []
    for d, (b, e) in enumerate(box):
      first, last = -lo[d], (e - b) - hi[d]
      if extend_to_grid_edge and b == 0:
        first = 0
      if extend_to_grid_edge and e == dims[d]:
        last = e - b
      assert last > first, 'box too small for the window'
      region.append(slice(first, last))
    region = tuple(reversed(region))
    for k, (_, haoda_type) in enumerate(library.outputs):
      got = _host(outputs[k][slices], haoda_type)
      common.assert_bit_exact(
          np.ascontiguousarray(got[region]),
          np.ascontiguousarray(want[k][region]),
          '%s %s box %s' % (name, 'x'.join(map(str, dims)), box))


# BASELINE.json configs 2-4: program, iterate, dims, box extents
CONFIGS = [
    ('jacobi2d', 64, (16384, 16384), (640, 512)),
    ('sobel2d', 1, (32768, 32768), (1024, 300)),
    ('denoise2d', 1, (32768, 32768), (1024, 300)),
    ('heat3d', 32, (1024, 1024, 1024), (256, 160, 144)),
    ('jacobi3d', 32, (1024, 1024, 1024), (256, 160, 144)),
]


@pytest.mark.parametrize('name,iterate,dims,size', CONFIGS,
                         ids=[c[0] for c in CONFIGS])
def test_full_size_equals_oracle_on_boxes(name, iterate, dims, size):
  library = soda_cuda.compile_stencil(common.stencil(name, iterate))
  need = sum(np.dtype(soda_cuda.NUMPY_TYPES[t]).itemsize
             for _, t in library.inputs + library.outputs * 2) * np.prod(
                 [float(n) for n in dims])
  if torch.cuda.mem_get_info()[0] < need * 1.2:
    pytest.skip('not enough device memory')
  orc = common.oracle(name, iterate)
  inputs = _device_inputs(library, dims, seed=41)
  outputs = [torch.empty(tuple(reversed(dims)), dtype=torch.from_numpy(
      np.empty(0, soda_cuda.NUMPY_TYPES[t])).dtype
      if np.dtype(soda_cuda.NUMPY_TYPES[t]).kind == 'f' else
      {1: torch.int8, 2: torch.int16, 4: torch.int32, 8: torch.int64}[
          np.dtype(soda_cuda.NUMPY_TYPES[t]).itemsize], device='cuda')
             for _, t in library.outputs]
  library.run_device(inputs, outputs, dims, 0,
                     torch.cuda.current_stream().cuda_stream)
  torch.cuda.synchronize()
  lo, hi = library.window(iterate)
  try:
    _check_boxes(name, dims, library, lo, hi, inputs, outputs, orc.run,
                 _boxes(dims, size))
  finally:
    del inputs, outputs
    library.release()
    torch.cuda.empty_cache()


def test_full_size_denoise3d_sixteen_applications():
  """BASELINE config 5: denoise3d 768^3 "iterate 16" = 16 applications with
  u <- output (the reference cannot iterate it, core.py:228-233)."""
  from soda import cuda_slab
  name, times, dims, size = 'denoise3d', 16, (768, 768, 768), (192, 160, 144)
  library = soda_cuda.compile_stencil(common.stencil(name, 1))
  orc = common.oracle(name, 1)
  inputs = _device_inputs(library, dims, seed=43)
  runner = cuda_slab.SlabRunner(library, dims, 0, 1, feedback={1: 0})
  runner.load_local(inputs)
  outputs = runner.run(times)
  torch.cuda.synchronize()
  lo1, hi1 = library.window(1)
  lo = [l * times for l in lo1]
  hi = [h * times for h in hi1]

  def reference(sub_in):
    f, u = sub_in
    for _ in range(times):
      u, = orc.run([f, u])
    return [u]
  try:
    _check_boxes(name, dims, library, lo, hi, inputs, outputs, reference,
                 _boxes(dims, size)[:4], extend_to_grid_edge=False)
  finally:
    del inputs, outputs, runner
    library.release()
    torch.cuda.empty_cache()
