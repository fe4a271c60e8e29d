"""The kernels BASELINE.json's configurations run (same program signature, so
the same tuned schedule and the same compiled library as bench.py uses) on
small grids whose rows are NOT 16-byte aligned (needs a GPU).

At the full sizes every row is aligned and the kernels take their TMA loads
and 128-bit stores; an odd width sends the same binaries through plain loads
and cell-by-cell stores instead.  Those paths share the unrolled step bodies
with the fast ones, and a compiler defect there shows
only as wrong cells — so each production binary is compared bit for bit with
the oracle on such a grid too, over several chunks of the streamed dimension.
"""
import pytest

import common
from soda import cuda as soda_cuda

pytestmark = pytest.mark.gpu

CASES = [
    ('jacobi2d', 64, (1061, 333)),
    ('jacobi2d', 64, (2050, 280)),
    ('blur', 1, (2001, 97)),
    ('sobel2d', 1, (1027, 131)),
    ('seidel2d', 2, (1061, 97)),
    ('denoise2d', 1, (1061, 97)),
    ('heat3d', 32, (131, 100, 90)),
    ('jacobi3d', 32, (197, 101, 83)),
    ('denoise3d', 1, (131, 45, 37)),
]


@pytest.mark.parametrize('name,iterate,dims', CASES,
                         ids=['%s-x%d' % (c[0], c[1]) for c in CASES])
def test_production_binary_on_an_unaligned_grid(name, iterate, dims,
                                                monkeypatch):
  orc = common.oracle(name, iterate)
  library = soda_cuda.compile_stencil(common.stencil(name, iterate))
  inputs = common.random_inputs(orc, dims, seed=11)
  want = orc.run(inputs)
  for chunks in (None, '3'):
    if chunks:
      monkeypatch.setenv('SODA_CUDA_CHUNKS', chunks)
    got = library.run(inputs)
    assert library.stats['used_tma'] == 0
    for k, (g, w) in enumerate(zip(got, want)):
      common.assert_bit_exact(g, w, '%s x%d %s output %d chunks %s' % (
          name, iterate, dims, k, chunks), any_nan=True)
