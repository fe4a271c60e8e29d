"""`half` tensors through the C ABI against the CPU oracle, bit for bit
(needs a GPU).  Semantics: tests/test_half.py."""
import numpy as np
import pytest

import common
import golden
import half_programs as hp
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
      # magnitudes that exercise binary16 rounding, subnormals and overflow
      scale = rng.choice([1e-6, 1e-3, 1.0, 300.0, 3e4], size=shape)
      arrays.append(((rng.random(shape) - 0.3) * scale).astype(dtype))
    else:
      # sums stay below binary16's 65504: float -> int of an infinity is
      # undefined in C++ (x86 gives INT_MIN, the GPU saturates)
      arrays.append(rng.integers(-15000, 15000, size=shape).astype(dtype))
  return arrays


@pytest.mark.parametrize('name,dims,options', hp.CASES,
                         ids=[c[0] for c in hp.CASES])
def test_half_program_matches_oracle(name, dims, options):
  stencil = hp.stencil_of(name)
  orc = golden.Oracle(stencil)
  library = soda_cuda.compile_stencil(stencil,
                                      options=codegen.Options(**options))
  for seed in (1, 2):
    inputs = _inputs(orc, dims, seed)
    want = orc.run(inputs)
    got = library.run(inputs)
    for k, (g, w) in enumerate(zip(got, want)):
      common.assert_bit_exact(g, w, '%s %s output %d' % (name, dims, k),
                              any_nan=True)
