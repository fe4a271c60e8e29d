"""Four-dimensional programs through the C ABI against the CPU oracle, bit
for bit (needs a GPU): TMA boxes of rank 4, two tiled dimensions read through
shared planes, host buffers cut into pieces and into slabs."""
import pytest

import common
import dim4_programs as d4
import golden
from soda import cuda as soda_cuda

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('name,dims', d4.CASES)
def test_four_dimensional_program_matches_oracle(name, dims, monkeypatch):
  stencil = d4.stencil_of(name)
  orc = golden.Oracle(stencil)
  library = soda_cuda.compile_stencil(stencil)
  inputs = common.random_inputs(orc, dims, seed=2)
  want = orc.run(inputs)
  for pieces, devices in (('1', None), ('3', None), ('2', '0,0')):
    monkeypatch.setenv('SODA_CUDA_PIECES', pieces)
    got = library.run(inputs, devices=devices)
    for k, (g, w) in enumerate(zip(got, want)):
      common.assert_bit_exact(g, w, '%s %s pieces %s devices %s' % (
          name, dims, pieces, devices))
  # ragged extents: the plain-load instance
  ragged = tuple(d + 1 for d in dims)
  inputs = common.random_inputs(orc, ragged, seed=3)
  monkeypatch.setenv('SODA_CUDA_PIECES', '1')
  for g, w in zip(library.run(inputs), orc.run(inputs)):
    common.assert_bit_exact(g, w, '%s %s' % (name, ragged))
