"""FPGA wire format without a GPU: the layout constants and the
numpy oracle against fixtures produced by the unmodified reference's own
generated loops (oracle/fpga_layout_ref.py --fixtures)."""
import glob
import os

import numpy as np
import pytest

import common
import fpga_layout as oracle_layout      # oracle/fpga_layout.py
from soda import core, fpga_layout

FIXTURES = sorted(glob.glob(os.path.join(common.GOLDEN_DIR, 'fpga_layout',
                                         '*.npz')))


def load(path):
  data = np.load(path)
  name = os.path.basename(path).split('_')[0]
  banks = [[int(b) for b in row if b >= 0] for row in data['banks']]
  stencil = core.Stencil.from_text(
      common.bench_text(name), tile_size=[int(t) for t in data['tile_size']],
      burst_width=int(data['burst_width']))
  tensors = list(stencil.input_stmts) + list(stencil.output_stmts)
  for stmt, dram in zip(tensors, banks):
    stmt.dram = tuple(dram)
  dims = tuple(int(n) for n in data['dims'])
  return data, stencil, fpga_layout.WireLayout(stencil, dims)


def test_fixtures_exist():
  assert len(FIXTURES) >= 4


@pytest.mark.parametrize('path', FIXTURES, ids=os.path.basename)
def test_layout_constants_match_the_reference(path):
  data, stencil, layout = load(path)
  assert list(layout.tile_num) == data['tile_num'].tolist()
  assert [layout.tile_size_linearized_i, layout.tile_size_linearized_o] == \
      data['tile_size_linearized'].tolist()
  names = list(stencil.input_names) + list(stencil.output_names)
  widths = [layout.descriptor(n).elem_size for n in names]
  assert [layout.bank_elems(n) * w for n, w in zip(names, widths)] == \
      data['bank_bytes'].tolist()


@pytest.mark.parametrize('path', FIXTURES, ids=os.path.basename)
def test_oracle_pack_and_unpack_match_the_reference(path):
  data, stencil, layout = load(path)
  for k, name in enumerate(stencil.input_names):
    want = [data['in%d_bank%d' % (k, b)] for b in range(4)]
    got = [np.zeros_like(w) for w in want]
    oracle_layout.pack(layout, name, data['in%d' % k], got)
    for b in range(4):
      common.assert_bit_exact(got[b], want[b], '%s bank %d' % (name, b))
    assert any(g.any() for g in got)
  for k, name in enumerate(stencil.output_names):
    banks = [data['out%d_bank%d' % (k, b)] for b in range(4)]
    got = np.zeros_like(data['out%d' % k])
    oracle_layout.unpack(layout, name, got, banks)
    common.assert_bit_exact(got, data['out%d' % k], name)
    assert got.any()


def test_tile_smaller_than_the_window_is_rejected():
  from haoda import util
  stencil = core.Stencil.from_text(common.bench_text('blur'), tile_size=[2])
  with pytest.raises(util.SemanticError):
    fpga_layout.WireLayout(stencil, (64, 8))


def test_configuration_the_reference_overruns_is_rejected():
  """4 input banks, 1 output bank: the output tiles (sized from the input's
  burst count, host.py:334-347) are shorter than their cell count and the
  last one ends beyond the buffer of host.py:399-415."""
  from haoda import util
  stencil = core.Stencil.from_text(common.bench_text('jacobi3d'),
                                   tile_size=[40, 17], burst_width=128)
  stencil.input_stmts[0].dram = (0, 1, 2, 3)
  stencil.output_stmts[0].dram = (1,)
  layout = fpga_layout.WireLayout(stencil, (100, 40, 23))
  layout.descriptor(stencil.input_names[0])
  with pytest.raises(util.SemanticError, match='beyond the bank buffer'):
    layout.descriptor(stencil.output_names[0])


def test_library_exports_the_declared_entry_points():
  import ctypes
  import re
  header = os.path.join(common.ROOT, 'include', 'soda_fpga_layout.h')
  with open(header) as handle:
    text = re.sub(r'/\*.*?\*/', '', handle.read(), flags=re.S)
  names = sorted(set(re.findall(r'\b(soda_fpga_\w+)\s*\(', text)))
  assert names == ['soda_fpga_pack', 'soda_fpga_unpack']
  lib = ctypes.CDLL(fpga_layout.build())
  for name in names:
    assert hasattr(lib, name)
  assert ctypes.sizeof(fpga_layout.TensorLayout) == 144


@pytest.mark.parametrize('path', FIXTURES, ids=os.path.basename)
def test_kernel_index_arithmetic_model_matches_the_reference(path):
  """tests/wire_kernel_model.py restates the CUDA kernel's decomposition and
  index arithmetic; it must reproduce the reference's fixtures too."""
  import wire_kernel_model as model
  data, stencil, layout = load(path)
  for k, name in enumerate(stencil.input_names):
    want = [data['in%d_bank%d' % (k, b)] for b in range(4)]
    got = {b: np.zeros_like(want[b]) for b in range(4)}
    model.run(layout.descriptor(name), data['in%d' % k].ravel(), got, True)
    for b in range(4):
      common.assert_bit_exact(got[b], want[b], '%s bank %d' % (name, b))
  for k, name in enumerate(stencil.output_names):
    banks = {b: data['out%d_bank%d' % (k, b)] for b in range(4)}
    got = np.zeros_like(data['out%d' % k])
    model.run(layout.descriptor(name), got.reshape(-1), banks, False)
    common.assert_bit_exact(got, data['out%d' % k], name)


def test_kernel_model_round_trips_both_mappings():
  import wire_kernel_model as model
  stencil = core.Stencil.from_text(common.bench_text('heat3d'),
                                   tile_size=[24, 20], burst_width=256)
  stencil.input_stmts[0].dram = (0, 2)
  stencil.output_stmts[0].dram = (1, 3)
  dims = (61, 50, 11)
  layout = fpga_layout.WireLayout(stencil, dims)
  shape = tuple(reversed(dims))
  dense = (np.random.default_rng(5).random(shape) * 1000).astype(np.float32)
  name_in, name_out = stencil.input_names[0], stencil.output_names[0]
  banks = {b: np.zeros(layout.bank_elems(name_in), np.float32)
           for b in range(4)}
  model.run(layout.descriptor(name_in), dense.ravel(), banks, True)
  back = np.zeros(shape, np.float32)
  model.run(layout.descriptor(name_in), back.reshape(-1), banks, False)
  common.assert_bit_exact(back, dense, 'input mapping')
  banks = {b: np.zeros(layout.bank_elems(name_out), np.float32)
           for b in range(4)}
  model.run(layout.descriptor(name_out), dense.ravel(), banks, True)
  back = np.full(shape, 77, np.float32)
  model.run(layout.descriptor(name_out), back.reshape(-1), banks, False)
  want = np.full(shape, 77, np.float32)
  inner = tuple(slice(l, n - h) for l, h, n in reversed(list(zip(
      layout.window_offset, layout.valid_hi_margin(), dims))))
  want[inner] = dense[inner]
  common.assert_bit_exact(back, want, 'output mapping')


@pytest.mark.skipif(not common.have_reference(),
                    reason='needs /root/reference')
def test_committed_fixtures_are_what_the_reference_loops_produce():
  """Re-runs oracle/fpga_layout_ref.py's extraction (the unmodified
  reference's generated pack / unpack nests, compiled) on the committed
  fixture inputs: the committed outputs must come out again."""
  import fpga_layout_ref as ref_tool
  name, tile, burst, dims, banks = ref_tool.FIXTURES[1]     # two banks each
  data = np.load(os.path.join(common.GOLDEN_DIR, 'fpga_layout',
                              '%s_%s.npz' % (name, 'x'.join(map(str, tile)))))
  ref = ref_tool.ReferenceLayout(
      os.path.join(common.REFERENCE_DIR, 'tests', 'src', name + '.soda'),
      tile, burst)
  packed = ref.pack(dims, banks, [data['in0']])
  for b in range(4):
    common.assert_bit_exact(packed[0][b], data['in0_bank%d' % b])
  unpacked = ref.unpack(dims, banks,
                        [[data['out0_bank%d' % b] for b in range(4)]])
  common.assert_bit_exact(unpacked[0], data['out0'])


def _random_layout(rng):
  from haoda import util
  name = rng.choice(['blur', 'sobel2d', 'denoise2d', 'heat3d', 'jacobi3d'])
  dim = 3 if name.endswith('3d') else 2
  tile = [int(rng.integers(8, 40)) for _ in range(dim - 1)]
  burst = int(rng.choice([64, 128, 256, 512]))
  stencil = core.Stencil.from_text(common.bench_text(name), tile_size=tile,
                                   burst_width=burst, iterate=1)
  n_banks = int(rng.integers(1, 5))            # same count on both sides
  for stmt in list(stencil.input_stmts) + list(stencil.output_stmts):
    stmt.dram = tuple(int(b) for b in rng.permutation(4)[:n_banks])
  dims = tuple(int(rng.integers(12, 70)) for _ in range(dim - 1)) + (
      int(rng.integers(6, 12)),)
  try:
    layout = fpga_layout.WireLayout(stencil, dims)
    for stmt in list(stencil.input_stmts) + list(stencil.output_stmts):
      layout.descriptor(stmt.name)
  except util.SemanticError:
    return None
  if min(layout.tile_num) < 1:
    return None
  return stencil, layout, dims


def test_kernel_model_equals_the_oracle_on_random_layouts():
  """The CUDA kernel's index arithmetic (tests/wire_kernel_model.py) against
  the numpy restatement of the reference loops, over random tile sizes,
  burst widths, bank assignments and grids."""
  import wire_kernel_model as model
  rng = np.random.default_rng(2024)
  done = 0
  for _ in range(60):
    made = _random_layout(rng)
    if made is None:
      continue
    stencil, layout, dims = made
    shape = tuple(reversed(dims))
    for stmt in stencil.input_stmts:
      width = layout.descriptor(stmt.name).elem_size
      dtype = {2: np.uint16, 4: np.float32}[width]
      dense = (rng.random(shape) * 60000).astype(dtype)
      count = layout.bank_elems(stmt.name)
      want = {b: np.full(count, 7, dtype) for b in range(4)}
      got = {b: np.full(count, 7, dtype) for b in range(4)}
      oracle_layout.pack(layout, stmt.name, dense, want)
      model.run(layout.descriptor(stmt.name), dense.ravel(), got, True)
      for b in range(4):
        common.assert_bit_exact(got[b], want[b], '%s %s' % (stmt.name, dims))
    for stmt in stencil.output_stmts:
      width = layout.descriptor(stmt.name).elem_size
      dtype = {2: np.uint16, 4: np.float32}[width]
      count = layout.bank_elems(stmt.name)
      banks = {b: (rng.random(count) * 60000).astype(dtype) for b in range(4)}
      want = np.full(shape, 9, dtype)
      got = np.full(shape, 9, dtype)
      oracle_layout.unpack(layout, stmt.name, want, banks)
      model.run(layout.descriptor(stmt.name), got.reshape(-1), banks, False)
      common.assert_bit_exact(got, want, '%s %s' % (stmt.name, dims))
    done += 1
  assert done >= 25
