"""Float quotients whose operands leave the quick division's range (needs a
GPU).  The exact build runs `a / b` and `a / sqrt(x)` at the top of a float
statement through soda::div_try — the quick sequence for every cell of a
vector, the plain quotient for the whole vector if any cell raised `rare`
(DESIGN.md section 3).  The benchmarks' data never raise it; here most
vectors do: operands spread over 2^-70 .. 2^70 with zeros, denormals,
infinities, NaNs and negative radicands mixed in, so both paths and every
mixture of them within a vector are compared with the oracle bit for bit
(a NaN matches a NaN of any payload: that is the machine's, not the
program's)."""
import numpy as np
import pytest

import common
import golden
from soda import core
from soda import cuda as soda_cuda
from soda.codegen import cuda as codegen

pytestmark = pytest.mark.gpu

HEADER = 'kernel: %s\nburst width: 64\nunroll factor: 1\niterate: 1\n'
PROGRAMS = {
    'quot2d': ('''input float: a(32, *)
input float: b(32, *)
local float: q(0, 0) = a(0, 0) / b(1, 0)
local float: g(0, 0) = 1.0f / sqrt(a(0, 1) + a(0, 0) * 0.5f)
output float: o(0, 0) = q(0, 0) * b(0, 0) / g(-1, 0)
''', [(1024, 70), (1061, 45)]),
    'quot3d': ('''input float: a(16, 8, *)
input float: b(16, 8, *)
local float: g(0, 0, 0) = 2.5f / sqrt(a(0, 0, 0) * a(0, 0, 0) + b(0, 1, 0))
output float: o(0, 0, 0) = (a(0, 0, 1) + g(1, 0, 0)) / (b(0, 0, 0) + g(0, 0, -1))
''', [(128, 40, 24), (131, 35, 19)]),
}


def _operands(shape, rng, wild):
  """Magnitudes 2^-70 .. 2^70 with random signs; ``wild``: one cell in eight
  is a zero, a denormal, an infinity, a NaN or the largest/smallest normal."""
  exponent = rng.uniform(-70.0, 70.0, size=shape)
  values = (np.exp2(exponent) * rng.uniform(1.0, 2.0, size=shape) *
            rng.choice((-1.0, 1.0), size=shape)).astype(np.float32)
  if wild:
    specials = np.array([0.0, -0.0, 1e-41, -3e-45, np.inf, -np.inf, np.nan,
                         3.4028235e38, 1.1754944e-38, 1.0, -1.0],
                        dtype=np.float32)
    pick = rng.random(shape) < 0.125
    values[pick] = rng.choice(specials, size=int(pick.sum()))
  return values


@pytest.mark.parametrize('name', sorted(PROGRAMS))
@pytest.mark.parametrize('wild', [False, True], ids=['wide', 'specials'])
@pytest.mark.parametrize('options', [{}, {'style': 'ring'}],
                         ids=['default', 'ring'])
def test_quotients_match_the_oracle_on_wild_operands(name, wild, options):
  body, cases = PROGRAMS[name]
  stencil = core.Stencil.from_text(HEADER % name + body)
  orc = golden.Oracle(stencil)
  library = soda_cuda.compile_stencil(stencil,
                                      options=codegen.Options(**options))
  rng = np.random.default_rng(17 if wild else 5)
  for dims in cases:
    shape = tuple(reversed(dims))
    inputs = [_operands(shape, rng, wild) for _ in orc.input_dtypes]
    with np.errstate(all='ignore'):
      want = orc.run(inputs)
    got = library.run(inputs)
    for k, (g, w) in enumerate(zip(got, want)):
      common.assert_bit_exact(g, w, '%s %s output %d' % (name, dims, k),
                              any_nan=True)
    finite = np.isfinite(want[0])
    assert 0.05 < finite.mean() < 1.0 or not wild     # both kinds occur
