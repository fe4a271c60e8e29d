"""One small run of a compiled SODA program, for compute-sanitizer
(tests/test_sanitizer_gpu.py runs this file under memcheck and racecheck).

  python tests/sanitizer_case.py <program> <iterate> <dims: AxBxC> [devices]

The libraries are prebuilt (tools/prebuild.py); the result is compared with
the CPU oracle so that a sanitised run that computes nonsense also fails.
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for sub in ('tests', 'oracle', 'soda-compiler_b200'):
  sys.path.insert(0, os.path.join(ROOT, sub))

import common                                   # noqa: E402
from soda import cuda as soda_cuda              # noqa: E402


def main():
  name, iterate = sys.argv[1], int(sys.argv[2])
  dims = tuple(int(x) for x in sys.argv[3].split('x'))
  devices = sys.argv[4] if len(sys.argv) > 4 else None
  orc = common.oracle(name, iterate)
  library = soda_cuda.compile_stencil(common.stencil(name, iterate))
  inputs = common.random_inputs(orc, dims, seed=3)
  want = orc.run(inputs)
  got = library.run(inputs, devices=devices)
  for g, w in zip(got, want):
    common.assert_bit_exact(g, w, name, any_nan=True)
  print('SANITIZER_CASE_OK %s x%d %s %s' % (name, iterate, dims, devices))


if __name__ == '__main__':
  main()
