"""Random programs (tests/random_programs.py) through the C ABI against the
CPU oracle, bit for bit (needs a GPU).

(File name chosen to sort last: written after the round's GPU budget was
spent; `pytest -x` reaches every measured test first.)
"""
import numpy as np
import pytest

import common
import golden
import random_programs as rp
from soda import cuda as soda_cuda

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('seed', rp.SEEDS)
def test_random_program_matches_oracle(seed):
  stencil = rp.stencil_of(seed)
  orc = golden.Oracle(stencil)
  library = soda_cuda.compile_stencil(stencil)
  dims = rp.dims_of(stencil, seed)
  rng = np.random.default_rng(seed)
  shape = tuple(reversed(dims))
  inputs = []
  for dtype in orc.input_dtypes:
    dtype = np.dtype(dtype)
    if dtype.kind == 'f':
      inputs.append((rng.random(shape) + 0.5).astype(dtype))
    else:
      info = np.iinfo(dtype)
      inputs.append(rng.integers(info.min, int(info.max) + 1, size=shape,
                                 dtype=np.int64).astype(dtype))
  want = orc.run(inputs)
  got = library.run(inputs)
  for k, (g, w) in enumerate(zip(got, want)):
    # products of up to four iterations may overflow to inf and NaN: those
    # compare by class (see common.assert_bit_exact)
    common.assert_bit_exact(g, w, 'seed %d:\n%s' % (
        seed, rp.program_text(seed)), any_nan=True)


@pytest.mark.parametrize('name', sorted(rp.EXTRA))
def test_outputs_read_by_later_statements(name):
  stencil = rp.extra_stencil(name)
  dims = rp.EXTRA[name][1]
  orc = golden.Oracle(stencil)
  library = soda_cuda.compile_stencil(stencil)
  for seed in (1, 2):
    inputs = common.random_inputs(orc, dims, seed=seed)
    want = orc.run(inputs)
    got = library.run(inputs)
    assert len(got) == 2
    for k, (g, w) in enumerate(zip(got, want)):
      common.assert_bit_exact(g, w, '%s output %d' % (name, k))
