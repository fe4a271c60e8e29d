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


# every 3rd seed: 20 programs whose outputs are defined on different boxes
MULTI_GPU_SEEDS = rp.MULTI_SEEDS[::3]


@pytest.mark.parametrize('seed', MULTI_GPU_SEEDS)
def test_multi_output_program_matches_oracle(seed):
  """Each output is zeroed outside ITS OWN valid box (reference
  host.py:1082-1091) and equals the golden loop inside it — whole arrays, bit
  for bit; host buffers (pipelined or not) and one-launch device runs."""
  stencil = rp.multi_stencil(seed)
  orc = golden.Oracle(stencil)
  library = soda_cuda.compile_stencil(stencil)
  dims = rp.multi_dims(stencil, seed)
  inputs = common.random_inputs(orc, dims, seed=seed)
  want = orc.run(inputs)
  got = library.run(inputs)
  regions = library.valid_regions(dims)
  assert len(got) == len(want) == len(regions)
  for k, (g, w) in enumerate(zip(got, want)):
    common.assert_bit_exact(g, w, 'seed %d output %d:\n%s' % (
        seed, k, rp.multi_program_text(seed)), any_nan=True)
    # the oracle's border is 0 exactly outside the reference's box
    box = tuple(slice(lo, hi) for lo, hi in reversed(regions[k]))
    outside = np.ones(w.shape, dtype=bool)
    outside[box] = False
    assert not w[outside].any() and not g[outside].any()
