"""FPGA wire-format kernels on the GPU (csrc/soda_fpga_layout.cu): bit-exact
against the fixtures produced by the reference's own generated loops and, on
larger grids, against the numpy oracle."""
import numpy as np
import pytest
import torch

import common
import fpga_layout as oracle_layout      # oracle/fpga_layout.py
import test_fpga_layout as cpu_side
from soda import core, fpga_layout

pytestmark = pytest.mark.gpu

_TORCH = {1: torch.uint8, 2: torch.int16, 4: torch.int32, 8: torch.int64}


def _to_device(array):
  return torch.from_numpy(array.view({1: np.uint8, 2: np.int16, 4: np.int32,
                                      8: np.int64}[array.itemsize])).cuda()


def _to_host(tensor, like):
  return tensor.cpu().numpy().view(like.dtype)


@pytest.mark.parametrize('path', cpu_side.FIXTURES,
                         ids=lambda p: p.split('/')[-1])
def test_kernels_match_the_reference_fixtures(path):
  data, stencil, layout = cpu_side.load(path)
  for k, name in enumerate(stencil.input_names):
    want = [data['in%d_bank%d' % (k, b)] for b in range(4)]
    banks = {b: torch.zeros_like(_to_device(want[b]))
             for b in layout.banks(name)}
    fpga_layout.pack(layout, name, _to_device(data['in%d' % k]), banks)
    torch.cuda.synchronize()
    for b in layout.banks(name):
      common.assert_bit_exact(_to_host(banks[b], want[b]), want[b],
                              '%s bank %d' % (name, b))
  for k, name in enumerate(stencil.output_names):
    want = data['out%d' % k]
    banks = {b: _to_device(data['out%d_bank%d' % (k, b)])
             for b in layout.banks(name)}
    dense = torch.zeros_like(_to_device(want))
    fpga_layout.unpack(layout, name, dense, banks)
    torch.cuda.synchronize()
    common.assert_bit_exact(_to_host(dense, want), want, name)


@pytest.mark.parametrize('name,tile,burst,dims,banks', [
    ('blur', [2000], 512, (4100, 301), [[0, 1], [2, 3]]),
    ('sobel2d', [129], 256, (1000, 77), [[3], [0, 1, 2]]),
    ('denoise3d', [32, 32], 512, (70, 61, 19), [[0], [1, 2], [3]]),
    ('jacobi3d', [40, 17], 128, (100, 40, 23), [[0, 1], [1, 3]]),
    ('heat3d', [24, 24], 256, (50, 60, 9), [[2], [0, 1, 3]]),
])
def test_kernels_match_the_oracle(name, tile, burst, dims, banks):
  stencil = core.Stencil.from_text(common.bench_text(name), tile_size=tile,
                                   burst_width=burst)
  tensors = list(stencil.input_stmts) + list(stencil.output_stmts)
  for stmt, dram in zip(tensors, banks):
    stmt.dram = tuple(dram)
  layout = fpga_layout.WireLayout(stencil, dims)
  rng = np.random.default_rng(11)
  shape = tuple(reversed(dims))
  for stmt in stencil.input_stmts:
    dtype = np.dtype({2: np.uint16, 4: np.float32}[
        layout.descriptor(stmt.name).elem_size])
    dense = (rng.random(shape) * 60000).astype(dtype)
    count = layout.bank_elems(stmt.name)
    want = {b: np.full(count, 7, dtype) for b in range(4)}
    oracle_layout.pack(layout, stmt.name, dense, want)
    got = {b: _to_device(np.full(count, 7, dtype)) for b in stmt.dram}
    fpga_layout.pack(layout, stmt.name, _to_device(dense), got)
    torch.cuda.synchronize()
    for b in stmt.dram:
      common.assert_bit_exact(_to_host(got[b], want[b]), want[b],
                              '%s bank %d' % (stmt.name, b))
  for stmt in stencil.output_stmts:
    dtype = np.dtype({2: np.uint16, 4: np.float32}[
        layout.descriptor(stmt.name).elem_size])
    count = layout.bank_elems(stmt.name)
    bank_arrays = {b: (rng.random(count) * 60000).astype(dtype)
                   for b in range(4)}
    want = np.full(shape, 9, dtype)
    oracle_layout.unpack(layout, stmt.name, want, bank_arrays)
    dense = _to_device(np.full(shape, 9, dtype))
    fpga_layout.unpack(layout, stmt.name, dense,
                       {b: _to_device(bank_arrays[b]) for b in stmt.dram})
    torch.cuda.synchronize()
    common.assert_bit_exact(_to_host(dense, want), want, stmt.name)


def test_both_mappings_run_in_both_directions():
  """The function picks the direction, the descriptor the mapping: a GPU
  stand-in for the FPGA kernel reads its inputs OUT of the input mapping and
  writes its outputs INTO the output mapping.  Round trips return the data."""
  stencil = core.Stencil.from_text(common.bench_text('heat3d'),
                                   tile_size=[24, 20], burst_width=256)
  stencil.input_stmts[0].dram = (0, 2)
  stencil.output_stmts[0].dram = (1, 3)
  dims = (61, 50, 11)
  layout = fpga_layout.WireLayout(stencil, dims)
  rng = np.random.default_rng(5)
  shape = tuple(reversed(dims))
  dense = (rng.random(shape) * 1000).astype(np.float32)
  name_in, name_out = stencil.input_names[0], stencil.output_names[0]
  # input mapping: every cell of the grid is in some tile
  banks = {b: torch.zeros(layout.bank_elems(name_in), dtype=torch.int32,
                          device='cuda') for b in layout.banks(name_in)}
  fpga_layout.pack(layout, name_in, _to_device(dense), banks)
  back = torch.zeros(shape, dtype=torch.int32, device='cuda')
  fpga_layout.unpack(layout, name_in, back, banks)
  torch.cuda.synchronize()
  common.assert_bit_exact(_to_host(back, dense), dense, 'input mapping')
  # output mapping: the cells whose window fits come back, the rest is kept
  banks = {b: torch.zeros(layout.bank_elems(name_out), dtype=torch.int32,
                          device='cuda') for b in layout.banks(name_out)}
  fpga_layout.pack(layout, name_out, _to_device(dense), banks)
  back = torch.full(shape, 77, dtype=torch.int32, device='cuda')
  fpga_layout.unpack(layout, name_out, back, banks)
  torch.cuda.synchronize()
  want = np.full(shape, 77, np.int32).view(np.float32)
  lo, hi = layout.window_offset, layout.valid_hi_margin()
  inner = tuple(slice(l, n - h) for l, h, n in
                reversed(list(zip(lo, hi, dims))))
  want[inner] = dense[inner]
  common.assert_bit_exact(_to_host(back, dense), want, 'output mapping')
