#!/usr/bin/env python3
"""Tune the kernel configuration of SODA programs on the GPU at hand.

  python tools/autotune.py --build-only jacobi2d:64:16384x16384 ...   # no GPU
  python tools/autotune.py [--record] jacobi2d:64:16384x16384 ...    # on a GPU

A case is ``program:iterate:dims`` (benchmarks/<program>.soda).  Candidates
come from soda.cuda_tune.candidates; ``--build-only`` compiles them (cached
in-tree, so a GPU box only has to time them); ``--record`` writes the winners
to soda/codegen/cuda/tuned.json, which the backend then uses by default.
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, 'soda-compiler_b200')]

os.environ['SODA_CUDA_TUNED'] = '0'      # candidates are the planner's, not
                                         # a previous winner's
from soda import core, cuda_tune                    # noqa: E402
from soda.codegen.cuda import plan                  # noqa: E402


def main():
  args = sys.argv[1:]
  build_only = '--build-only' in args
  record = '--record' in args
  cases = [a for a in args if not a.startswith('--')]
  summary = []
  for text in cases:
    name, iterate, dims = text.split(':')[:3]
    dims = tuple(int(x) for x in dims.split('x'))
    stencil = core.Stencil.from_file(
        os.path.join(ROOT, 'benchmarks', name + '.soda'), iterate=int(iterate))
    program = plan.extract_program(stencil)
    option_sets = cuda_tune.candidates(program)
    if build_only:
      built = cuda_tune.build_all(stencil, option_sets)
      bad = [(o, e) for o, e in built if isinstance(e, Exception)]
      print('%-40s %d candidates built, %d failed' % (
          text, len(built) - len(bad), len(bad)), flush=True)
      for options, error in bad:
        print('   %s: %s' % (options, str(error)[:200]))
      continue
    import torch
    print('== %s' % text, flush=True)
    results = cuda_tune.tune(stencil, dims, option_sets,
                             log=lambda line: print('  ' + line, flush=True))
    if not results:
      continue
    cells = 1.0
    for n in dims:
      cells *= n
    default_ms = next((ms for ms, o in results if not o), None)
    best_ms, best = results[0]
    summary.append({
        'case': text, 'best': best, 'best_ms': round(best_ms, 4),
        'best_gcell_per_s': round(cells * int(iterate) / best_ms / 1e6, 1),
        'planner_ms': default_ms and round(default_ms, 4),
        'planner_gcell_per_s': default_ms and round(
            cells * int(iterate) / default_ms / 1e6, 1),
        'candidates': len(results)})
    print(json.dumps(summary[-1]), flush=True)
    # a winner inside the noise of the planner's choice is not recorded
    if record and best and default_ms and best_ms < 0.98 * default_ms:
      cuda_tune.record(program, dims, best_ms, best,
                       torch.cuda.get_device_name())
    torch.cuda.empty_cache()


if __name__ == '__main__':
  main()
