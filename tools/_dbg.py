import sys, os
ROOT='/root/repo'
sys.path[:0]=[ROOT+'/soda-compiler_b200', ROOT+'/tests', ROOT+'/oracle', ROOT+'/tools']
import numpy as np
import common, quick_bench
from soda import cuda as soda_cuda
from soda.codegen import cuda as codegen
for text in sys.argv[1:]:
  name, it, dims, opts = quick_bench.parse_case(text)
  orc = common.oracle(name, it)
  inputs = orc.reference_inputs(dims)
  want = orc.run(inputs)
  try:
    lib = soda_cuda.compile_stencil(common.stencil(name, it), options=codegen.Options(**opts))
    got = lib.run(inputs)
    print(text, 'equal', all(np.array_equal(g.view(np.uint8), w.view(np.uint8)) for g,w in zip(got,want)), 'tma', lib.stats['used_tma'], flush=True)
  except Exception as e:
    print(text, 'FAILED', str(e)[:200], flush=True)
    break
