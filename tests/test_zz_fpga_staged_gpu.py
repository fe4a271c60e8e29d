"""The staged variants of the wire-format kernels (a tile row goes through
shared memory so that both global sides use wide accesses; for unpack also
the software-pipelined kernel, SODA_FPGA_PIPELINED) must move exactly the
bytes the element-wise variant moves.  They are the default for rows of 512
bytes and more; the environment switches pin each variant here.
"""
import numpy as np
import pytest
import torch

import common
import test_fpga_layout as cpu_side
import test_fpga_layout_gpu as gpu_side
from soda import core, fpga_layout

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('pipelined', ['0', '1'])
@pytest.mark.parametrize('path', cpu_side.FIXTURES,
                         ids=lambda p: p.split('/')[-1])
def test_staged_kernels_match_the_reference_fixtures(path, pipelined,
                                                     monkeypatch):
  """Both unpack kernels of the staged family: the warp-per-row one and the
  software-pipelined one (the default where a row fits its registers)."""
  monkeypatch.setenv('SODA_FPGA_STAGED', '1')
  monkeypatch.setenv('SODA_FPGA_PIPELINED', pipelined)
  gpu_side.test_kernels_match_the_reference_fixtures(path)


@pytest.mark.parametrize('name,tile,burst,dims,banks', [
    ('blur', [2000], 512, (4100, 301), [[0, 1], [2, 3]]),
    ('sobel2d', [129], 256, (1000, 77), [[3], [0, 1, 2]]),
    ('denoise3d', [32, 32], 512, (70, 61, 19), [[0], [1, 2], [3]]),
    ('heat3d', [24, 24], 256, (50, 60, 9), [[2], [0, 1, 3]]),
    ('jacobi3d', [40, 17], 128, (101, 40, 23), [[0, 1, 2, 3], [0, 1, 2, 3]]),
    # rows longer than the pipelined kernel's registers hold (falls back)
    ('blur', [4000], 512, (9000, 40), [[0, 1], [2, 3]]),
    ('jacobi2d', [1500], 512, (4000, 50), [[0, 1, 2], [1, 2, 3]]),
    ('jacobi2d', [700], 512, (2000, 50), [[1], [0, 2, 3]]),
])
def test_staged_equals_default(name, tile, burst, dims, banks, monkeypatch):
  stencil = core.Stencil.from_text(common.bench_text(name), tile_size=tile,
                                   burst_width=burst)
  tensors = list(stencil.input_stmts) + list(stencil.output_stmts)
  for stmt, dram in zip(tensors, banks):
    stmt.dram = tuple(dram)
  layout = fpga_layout.WireLayout(stencil, dims)
  gen = torch.Generator(device='cuda')
  gen.manual_seed(3)
  shape = tuple(reversed(dims))
  for stmt in tensors:
    width = layout.descriptor(stmt.name).elem_size
    dtype = {2: torch.int16, 4: torch.int32}[width]
    info = torch.iinfo(dtype)
    dense = torch.randint(info.min, info.max, shape, generator=gen,
                          device='cuda', dtype=dtype)
    count = layout.bank_elems(stmt.name)
    results = []
    # element-wise; a row per warp; the same with the pipelined unpack
    for staged, pipelined in (('0', '0'), ('1', '0'), ('1', '1')):
      monkeypatch.setenv('SODA_FPGA_STAGED', staged)
      monkeypatch.setenv('SODA_FPGA_PIPELINED', pipelined)
      filled = {b: torch.full((count,), 5, dtype=dtype, device='cuda')
                for b in stmt.dram}
      fpga_layout.pack(layout, stmt.name, dense, filled)
      back = torch.full(shape, 9, dtype=dtype, device='cuda')
      fpga_layout.unpack(layout, stmt.name, back, filled)
      torch.cuda.synchronize()
      results.append(([filled[b].cpu().numpy() for b in stmt.dram],
                      back.cpu().numpy()))
    for other in results[1:]:
      for plain, staged in zip(results[0][0], other[0]):
        common.assert_bit_exact(staged, plain, '%s pack' % stmt.name)
      common.assert_bit_exact(other[1], results[0][1], '%s unpack' % stmt.name)
    assert (results[0][1] != 9).any()
