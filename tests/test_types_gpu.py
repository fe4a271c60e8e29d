"""Integer programs outside the benchmark set, through the C ABI (needs a GPU).

The register kernels keep 8- and 16-bit integer tensors in 32-bit registers
with unspecified upper bits and narrow them only where a consumer can tell
(kernel_reg.py, "lazy truncation").  These programs mix polynomial stages
(read the loose registers as they are) with stages that divide, compare or
widen (must read through a cast), signed and unsigned, 2-D and 3-D (shared
planes hold the declared type).  Bar: bit-exact against the CPU oracle, whole
array, random inputs covering the full range of the type.
"""
import numpy as np
import pytest

import common
import golden
from soda import core, cuda as soda_cuda
from soda.codegen import cuda as codegen

pytestmark = pytest.mark.gpu

HEADER = 'kernel: %s\nburst width: 64\nunroll factor: 1\niterate: %d\n'

PROGRAMS = {
    'mix8': (1, '''input uint8: a(32, *)
local uint8: s(0, 0) = a(0, 0) + a(1, 0) * 3 - a(0, 1)
local uint8: d(0, 0) = (s(0, 0) + s(-1, 0)) / 2
output uint8: o(0, 0) = d(0, 0) * d(0, 1) - s(1, 0)
'''),
    'mix16': (1, '''input int16: a(32, *)
local int32: w(0, 0) = a(0, 0) * a(1, 0) - a(0, 1)
local int16: n(0, 0) = w(0, 0) + a(0, 0) * 7
output int16: o(0, 0) = n(0, 0) - n(-1, 0) * n(0, -1) + a(0, 0) / 4
'''),
    'iter16': (3, '''input uint16: a(32, *)
local uint16: s(0, 0) = a(-1, 0) * a(1, 0) + a(0, -1) - a(0, 1) * 65533
output uint16: b(0, 0) = s(0, 0) - (a(0, 0) % 7)
'''),
    'cmp8': (1, '''input int8: a(32, *)
input int8: b(32, *)
local int8: m(0, 0) = a(0, 0) * b(0, 0) - a(1, 0)
output int8: o(0, 0) = m(0, 0) + (m(-1, 0) < b(0, 1)) * 100 - m(0, 1) * a(0, -1)
'''),
    'vol16': (1, '''input uint16: a(16, 8, *)
local uint16: s(0, 0, 0) = a(0, 0, 0) + a(0, 1, 0) + a(0, 0, 1) * 5 - a(1, 0, 0)
output uint16: o(0, 0, 0) = s(0, 0, 0) * s(0, -1, 0) + s(-1, 0, 0) - s(0, 0, -1) / 3
'''),
}

# program, dims, backend options
CASES = [
    ('mix8', (1061, 97), {}), ('mix8', (2048, 160), {}),
    ('mix8', (2048, 160), {'vec': 4}), ('mix8', (2048, 160), {'vec': 16}),
    ('mix16', (777, 131), {}), ('mix16', (2048, 96), {}),
    ('iter16', (1536, 200), {}), ('iter16', (1001, 75), {'depth': 1}),
    ('iter16', (2048, 75), {'depth': 3}),
    ('cmp8', (2048, 128), {}), ('cmp8', (999, 64), {}),
    ('vol16', (256, 48, 40), {}), ('vol16', (131, 35, 29), {}),
    ('mix8', (2048, 160), {'style': 'ring'}),
]


def stencil_of(name):
  iterate, body = PROGRAMS[name]
  return core.Stencil.from_text(HEADER % (name, iterate) + body)


@pytest.mark.parametrize('name,dims,options', CASES)
def test_integer_program_matches_oracle(name, dims, options):
  stencil = stencil_of(name)
  orc = golden.Oracle(stencil)
  library = soda_cuda.compile_stencil(stencil,
                                      options=codegen.Options(**options))
  for seed in (1, 2):
    inputs = common.random_inputs(orc, dims, seed=seed)
    want = orc.run(inputs)
    got = library.run(inputs)
    for k, (g, w) in enumerate(zip(got, want)):
      common.assert_bit_exact(g, w, '%s %s output %d' % (name, dims, k))
