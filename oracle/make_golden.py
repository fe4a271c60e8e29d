#!/usr/bin/env python3
"""Generate the committed fixtures under tests/golden/ — TEST TOOLING.

Needs /root/reference (run in the build container; the fixtures travel).

reference_stages.json
    For every reference benchmark (/root/reference/tests/src/*.soda) at its own
    ``iterate`` and at iterate 1 and 3: what the UNMODIFIED reference frontend
    makes of it — tensor names, stage order, store indices, golden-loop bounds
    and each stage's lowered C expression (``oracle/ref_tool.py describe``).
    Pins this repo's frontend and expression lowering.
outputs/<app>_it<N>_<dims>.npz
    Output arrays for the reference initialiser's inputs at a small size,
    computed by the CPU oracle and accepted with ZERO mismatches by the
    reference's own ``<app>_test`` harness (host.print_test, compiled from the
    reference's emitted code; oracle/ref_harness.py) at generation time.  Pins
    the oracle, and through it the CUDA path, wherever the reference is absent.
"""
import glob
import json
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
import golden        # noqa: E402
import ref_harness   # noqa: E402

REF_SRC = '/root/reference/tests/src'
OUT = os.path.join(ROOT, 'tests', 'golden')
VECTORS = [('blur', 1, (64, 40)), ('blur', 1, (2000, 1000)),
           ('sobel2d', 1, (48, 40)),
           ('jacobi2d', 2, (48, 40)), ('jacobi2d', 5, (64, 48)),
           ('seidel2d', 2, (48, 40)), ('denoise2d', 1, (48, 40)),
           ('jacobi3d', 2, (20, 18, 16)), ('heat3d', 2, (20, 18, 16)),
           ('heat3d', 3, (24, 20, 18)), ('denoise3d', 1, (20, 18, 16))]


def main():
  os.makedirs(os.path.join(OUT, 'outputs'), exist_ok=True)
  stages = {}
  for path in sorted(glob.glob(os.path.join(REF_SRC, '*.soda'))):
    name = os.path.basename(path)[:-5]
    for iterate in (None, 1, 3):
      command = [sys.executable, os.path.join(HERE, 'ref_tool.py'), 'describe',
                 path] + (['--iterate', str(iterate)] if iterate else [])
      done = subprocess.run(command, stdout=subprocess.PIPE,
                            stderr=subprocess.PIPE, text=True, check=False)
      key = '%s@%s' % (name, iterate or 'default')
      if done.returncode == 0:
        stages[key] = json.loads(done.stdout)
      else:   # e.g. denoise with iterate 3: record the error class + message
        stages[key] = {'error': done.stderr.strip().splitlines()[-1]}
  with open(os.path.join(OUT, 'reference_stages.json'), 'w') as handle:
    json.dump(stages, handle, indent=1, sort_keys=True)
    handle.write('\n')

  for name, iterate, dims in VECTORS:
    soda_file = os.path.join(REF_SRC, name + '.soda')
    stencil = golden.stencil_from_file(soda_file, iterate)
    oracle = golden.Oracle(stencil)
    harness = ref_harness.RefHarness(
        ref_harness.build_ref(soda_file, iterate), stencil)
    kept = {}

    def implementation(inputs, kept=kept, oracle=oracle):
      kept['outputs'] = oracle.run(inputs)
      return kept['outputs']
    errors = harness.test(dims, implementation)
    assert errors == 0, (name, iterate, dims, errors)
    target = os.path.join(OUT, 'outputs', '%s_it%d_%s.npz' % (
        name, iterate, 'x'.join(map(str, dims))))
    np.savez_compressed(target, **{
        'out%d' % k: a for k, a in enumerate(kept['outputs'])})
    print('wrote', os.path.relpath(target, ROOT),
          os.path.getsize(target), 'bytes')


if __name__ == '__main__':
  main()
