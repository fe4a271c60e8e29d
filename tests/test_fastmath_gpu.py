"""The optional fast build (-DSODA_CUDA_FAST_MATH: FMA contraction, float
overloads of the math calls) against the oracle, within the tolerance
BASELINE.json's north_star states for non-exact float builds: 1e-6 relative
or 2 ulp, whichever is larger.  (The reference's own comparator accepts 1e-5
relative, src/soda/codegen/xilinx/host.py:1118-1146.)  The default build is
bit-exact and is what every other parity test runs.
"""
import numpy as np
import pytest

import common
from soda import cuda as soda_cuda

pytestmark = pytest.mark.gpu

REL_TOL = 1e-6
ULP_TOL = 2

CASES = [
    ('jacobi2d', 8, (2048, 260)),
    ('seidel2d', 4, (1280, 160)),
    ('heat3d', 4, (192, 48, 40)),
    ('jacobi3d', 4, (128, 64, 40)),
    ('denoise2d', 1, (1024, 300)),
    ('denoise3d', 1, (128, 48, 40)),
]


@pytest.mark.parametrize('name,iterate,dims', CASES,
                         ids=[c[0] for c in CASES])
def test_fast_build_within_tolerance(name, iterate, dims):
  library = soda_cuda.compile_stencil(common.stencil(name, iterate),
                                      fast_math=True)
  orc = common.oracle(name, iterate)
  for inputs in (orc.reference_inputs(dims),
                 common.random_inputs(orc, dims, seed=5)):
    want = orc.run(inputs)
    got = library.run(inputs)
    for g, w in zip(got, want):
      assert np.isfinite(w).all() and np.isfinite(g).all()
      err = np.abs(g.astype(np.float64) - w.astype(np.float64))
      bound = np.maximum(REL_TOL * np.abs(w).astype(np.float64),
                         ULP_TOL * np.spacing(np.abs(w)).astype(np.float64))
      worst = float((err / np.maximum(bound, 1e-300)).max())
      assert worst <= 1.0, (
          '%s: worst error is %.2f x the tolerance (max rel %.3g)' % (
              name, worst, float((err / np.maximum(np.abs(w), 1e-30)).max())))
