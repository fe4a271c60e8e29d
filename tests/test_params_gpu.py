"""Programs with `param` statements on the GPU, through the C ABI: bit-exact
against the oracle (see tests/test_params.py for what pins the oracle)."""
import numpy as np
import pytest
import torch

import common
import golden
import param_programs as pp
from soda import cuda as soda_cuda
from soda.codegen import cuda as codegen

pytestmark = pytest.mark.gpu


def _params(orc, seed):
  rng = np.random.default_rng(seed)
  arrays = []
  for _, dtype, size in orc.params:
    if np.dtype(dtype).kind == 'f':
      arrays.append(rng.random(size, dtype=np.float32).astype(dtype))
    else:
      values = rng.integers(-9, 10, size=size).astype(dtype)
      values[values == 0] = 3          # params also appear as divisors
      arrays.append(values)
  return arrays


@pytest.mark.parametrize('name,dims,options', pp.CASES)
def test_param_program_matches_oracle(name, dims, options):
  stencil = pp.stencil_of(name)
  orc = golden.Oracle(stencil)
  library = soda_cuda.compile_stencil(stencil,
                                      options=codegen.Options(**options))
  cases = [(orc.reference_inputs(dims), orc.reference_params())]
  cases += [(common.random_inputs(orc, dims, seed=s), _params(orc, s))
            for s in (1, 2)]
  for inputs, params in cases:
    want = orc.run(inputs, params=params)
    got = library.run(inputs, params=params)
    for k, (g, w) in enumerate(zip(got, want)):
      common.assert_bit_exact(g, w, '%s %s output %d' % (name, dims, k))


def test_device_entry_uses_the_params_last_set():
  stencil = pp.stencil_of('relax')
  orc = golden.Oracle(stencil)
  library = soda_cuda.compile_stencil(stencil)
  dims = (1024, 120)
  x, = common.random_inputs(orc, dims, seed=7)
  dev_in = torch.from_numpy(x).cuda()
  dev_out = torch.empty_like(dev_in)
  library.release()                     # forget params of earlier tests
  with pytest.raises(soda_cuda.CudaError) as info:
    library.run_device([dev_in], [dev_out], dims)
  assert info.value.code == -12         # params not set
  for seed in (1, 2):
    params = _params(orc, seed)
    library.set_params(params)
    library.run_device([dev_in], [dev_out], dims, 0,
                       torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    common.assert_bit_exact(dev_out.cpu().numpy(),
                            orc.run([x], params=params)[0],
                            'relax, params of seed %d' % seed)
