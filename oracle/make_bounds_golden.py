#!/usr/bin/env python3
"""Generate tests/golden/multi_output_bounds.json — TEST TOOLING.

Needs /root/reference (run in the build container; the fixture travels).

For the seeded multi-output programs of tests/random_programs.py
(``MULTI_SEEDS``) and its hand-written ``EXTRA`` programs: the golden-loop
bounds the UNMODIFIED reference frontend gives every non-input tensor
(``oracle/ref_tool.py describe``: ``lo`` and ``hi_margin`` per stage, i.e.
reference src/soda/codegen/xilinx/host.py:1082-1091 with the window from all
inputs to that tensor, src/soda/core.py:793-835).  Outputs of one program are
defined on different boxes; this pins the per-output valid regions of the
CUDA backend (plan.Program.window_of) to the reference.
"""
import json
import os
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path[:0] = [os.path.join(ROOT, 'tests'),
                os.path.join(ROOT, 'soda-compiler_b200')]
import random_programs as rp   # noqa: E402


def main():
  texts = {'multi%d' % seed: rp.multi_program_text(seed)
           for seed in rp.MULTI_SEEDS}
  texts.update({name: text for name, (text, _) in rp.EXTRA.items()})
  fixture = {}
  with tempfile.TemporaryDirectory() as tmp:
    for name, text in sorted(texts.items()):
      path = os.path.join(tmp, name + '.soda')
      with open(path, 'w') as handle:
        handle.write(text)
      done = subprocess.run(
          [sys.executable, os.path.join(HERE, 'ref_tool.py'), 'describe',
           path], stdout=subprocess.PIPE, text=True, check=True)
      described = json.loads(done.stdout)
      fixture[name] = {
          'text': text,
          'outputs': [n for n, _ in described['outputs']],
          'bounds': {stage['name']: [stage['lo'], stage['hi_margin']]
                     for stage in described['stages']}}
  out = os.path.join(ROOT, 'tests', 'golden', 'multi_output_bounds.json')
  with open(out, 'w') as handle:
    json.dump(fixture, handle, indent=1, sort_keys=True)
    handle.write('\n')
  print('wrote %s: %d programs' % (out, len(fixture)))


if __name__ == '__main__':
  main()
