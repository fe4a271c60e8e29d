"""The C ABI boundary: include/soda_cuda.h vs what a compiled library exports.

No GPU needed: the libraries are cross-compiled here and only loaded; the one
compute entry that is called must fail loudly (no CPU fallback).
"""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

import common
from soda import cuda as soda_cuda

HEADER = os.path.join(common.ROOT, 'include', 'soda_cuda.h')


@pytest.fixture(scope='module')
def library():
  return soda_cuda.compile_stencil(common.stencil('jacobi2d', 3))


def declared_functions():
  with open(HEADER) as handle:
    text = re.sub(r'/\*.*?\*/', '', handle.read(), flags=re.S)
  return sorted(set(re.findall(r'\b(soda_cuda_\w+)\s*\(', text)))


def test_every_declared_symbol_is_exported(library):
  names = declared_functions()
  assert len(names) >= 14
  symbols = subprocess.run(['nm', '-D', '--defined-only', library.path],
                           stdout=subprocess.PIPE, text=True,
                           check=True).stdout
  for name in names:
    assert re.search(r'\bT %s\b' % name, symbols), name
    assert hasattr(ctypes.CDLL(library.path), name)


def test_reference_entry_point_has_cxx_linkage(library):
  """`int jacobi2d(buffer_t*, buffer_t*, const char*)` mangles exactly like
  the function the reference header declares (header.py:57-60)."""
  symbols = subprocess.run(['nm', '-D', '--defined-only', library.path],
                           stdout=subprocess.PIPE, text=True,
                           check=True).stdout
  assert ' T _Z8jacobi2dP8buffer_tS0_PKc' in symbols


def test_buffer_t_layout():
  # reference header.py:36-48; offsets verified in SURVEY.md 8b
  assert ctypes.sizeof(soda_cuda.BufferT) == 72
  for field, offset in (('dev', 0), ('host', 8), ('extent', 16),
                        ('stride', 32), ('min', 48), ('elem_size', 64)):
    assert getattr(soda_cuda.BufferT, field).offset == offset


def test_identity_queries(library):
  assert library.app_name == 'jacobi2d'
  assert (library.dim, library.iterate) == (2, 3)
  assert library.inputs == [('t1', 'float')]
  assert library.outputs == [('t0', 'float')]
  assert library.depths and sum(library.depths) >= 1
  assert library.window() == ((-3, -3), (3, 3))
  assert library.window(1) == ((-1, -1), (1, 1))
  assert library.valid_region((100, 50)) == [(3, 97), (3, 47)]


def test_bounds_query_mode_needs_no_device(library):
  """host == NULL && dev == 0 fills in shapes and computes nothing
  (reference host.py:204-252): input extent = output extent + window - 1."""
  out = soda_cuda.BufferT()
  out.extent[0], out.extent[1] = 100, 60
  inp = soda_cuda.BufferT()
  code = library._lib.soda_cuda_run(
      (ctypes.POINTER(soda_cuda.BufferT) * 1)(ctypes.pointer(inp)),
      (ctypes.POINTER(soda_cuda.BufferT) * 1)(ctypes.pointer(out)), None)
  assert code == 0
  assert list(out.stride[:2]) == [1, 100] and out.elem_size == 4
  assert list(inp.extent[:2]) == [106, 66]
  assert list(inp.stride[:2]) == [1, 106] and inp.elem_size == 4


def test_no_cpu_fallback(library):
  import torch
  if torch.cuda.is_available():
    pytest.skip('a GPU is present')
  with pytest.raises(soda_cuda.CudaError) as info:
    library.run([np.zeros((64, 64), dtype=np.float32)])
  assert info.value.code == -19       # no_device_interface


def test_argument_checks(library):
  with pytest.raises(TypeError):
    library.run([np.zeros((64, 64), dtype=np.float64)])
  with pytest.raises(ValueError):
    library.run([np.zeros((4, 64, 64), dtype=np.float32)])
  with pytest.raises(ValueError):
    library.run([np.zeros((64, 64), dtype=np.float32)[:, ::2]])


def test_per_output_windows_are_exported(library):
  assert library.window_of(0) == ((-3, -3), (3, 3))
  assert library.valid_regions((100, 50), 1) == [[(1, 99), (1, 49)]]
  with pytest.raises(soda_cuda.CudaError):
    library.window_of(1)


@pytest.mark.parametrize('rows,n', [(1000, 4), (1001, 3), (50, 8), (16384, 8),
                                    (7, 2)])
def test_shard_plan_covers_the_grid_with_ghosts_of_the_whole_run(library, rows,
                                                                 n):
  """soda_cuda_shard_plan (no device needed): the slabs of a sharded run own
  disjoint row ranges that cover the grid, and hold ghost rows of the whole
  run's reach (3 iterations of jacobi2d: 3 rows) wherever the grid continues."""
  plan = library.shard_plan((256, rows), n)
  assert 1 <= len(plan) <= n
  lo, hi = library.window()
  ghost_lo, ghost_hi = max(0, -lo[-1]), max(0, hi[-1])
  previous = 0
  for local_begin, local_end, own_begin, own_end in plan:
    assert own_begin == previous and own_end > own_begin
    assert local_begin == max(0, own_begin - ghost_lo)
    assert local_end == min(rows, own_end + ghost_hi)
    previous = own_end
  assert previous == rows
  if len(plan) > 1:       # never more ghost than grid
    assert min(b - a for _, _, a, b in plan) >= ghost_lo + ghost_hi
  sizes = [b - a for _, _, a, b in plan]
  assert max(sizes) - min(sizes) <= 1
