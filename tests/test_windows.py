"""Every output is defined on the box of its OWN window.

The reference bounds each non-input tensor's golden loop by the window from
all inputs to that tensor (src/soda/codegen/xilinx/host.py:1082-1091,
src/soda/core.py:793-835).  Outputs of one program therefore differ: in
`chain2`, o0 is defined on [0, W-1) x [0, H) and o1 on [1, W) x [0, H-1).
The planner's per-output windows (plan.Program.window_of), which reach the
kernels as one valid box per output (StreamArgs::valid_lo/hi), must equal

* this repo's Stencil IR (`Stencil.valid_bounds`, pinned to the reference by
  tests/test_frontend.py), and
* tests/golden/multi_output_bounds.json: the bounds the UNMODIFIED reference
  frontend computes for 62 multi-output programs (oracle/make_bounds_golden.py).
"""
import json
import os

import pytest

import common
import random_programs as rp
from soda import core
from soda.codegen import cuda as codegen
from soda.codegen.cuda import plan

with open(os.path.join(common.GOLDEN_DIR, 'multi_output_bounds.json')) as _h:
  FIXTURE = json.load(_h)


def _texts():
  texts = {'multi%d' % seed: rp.multi_program_text(seed)
           for seed in rp.MULTI_SEEDS}
  texts.update({name: text for name, (text, _) in rp.EXTRA.items()})
  return texts


def test_fixture_covers_the_generated_programs():
  texts = _texts()
  assert set(texts) == set(FIXTURE) and len(FIXTURE) >= 52
  for name, text in texts.items():
    assert FIXTURE[name]['text'] == text, name
  # the point of the fixture: outputs of one program have different boxes
  differing = sum(
      1 for entry in FIXTURE.values()
      if len({json.dumps(entry['bounds'][out])
              for out in entry['outputs']}) > 1)
  assert differing >= 50


@pytest.mark.parametrize('name', sorted(FIXTURE))
def test_output_windows_match_the_reference(name):
  entry = FIXTURE[name]
  stencil = core.Stencil.from_text(entry['text'])
  program = plan.extract_program(stencil)
  codegen.check_supported(program)
  dims = (97, 61) if program.dim == 2 else (45, 33, 29)
  regions = program.valid_regions(dims)
  assert [n for n, _ in program.outputs] == entry['outputs']
  for k, out_name in enumerate(entry['outputs']):
    low, margin = entry['bounds'][out_name]
    # the Stencil IR agrees with the reference ...
    assert stencil.valid_bounds(stencil.tensors[out_name]) == (
        tuple(low), tuple(margin)), out_name
    # ... and so does the planner, per output
    assert regions[k] == [(l, d - m) for l, m, d in zip(low, margin, dims)], \
        out_name
    lo, hi = program.window_of(k)
    assert tuple(max(0, -x) for x in lo) == tuple(low)
    assert tuple(max(0, x) for x in hi) == tuple(margin)
  # the union window (halo sizing) covers every output's own
  ulo, uhi = program.window()
  for k in range(len(entry['outputs'])):
    lo, hi = program.window_of(k)
    assert all(a <= b for a, b in zip(ulo, lo))
    assert all(a >= b for a, b in zip(uhi, hi))


def test_chain2_regions_are_the_verdicts():
  """The case the round-1 GPU run failed on (1170 cells of o0 zeroed)."""
  program = plan.extract_program(rp.extra_stencil('chain2'))
  assert program.valid_regions((1024, 150)) == [
      [(0, 1022), (0, 149)], [(1, 1023), (0, 148)]]
  program = plan.extract_program(rp.extra_stencil('chain3d'))
  assert program.valid_regions((128, 40, 36)) == [
      [(0, 126), (2, 40), (2, 36)], [(0, 126), (2, 39), (2, 35)]]


def test_intermediate_iteration_windows_match_the_ir():
  """Per-output windows after n < iterate iterations (remainder launches and
  the slab runner use them) against the IR's `<input>_iter<n>` replicas."""
  for name in ('chain2', 'chain3d', 'multi101', 'multi104', 'multi110'):
    stencil = core.Stencil.from_text(FIXTURE[name]['text'])
    if stencil.iterate < 2:
      continue
    program = plan.extract_program(stencil)
    for n in range(1, stencil.iterate):
      for k, in_name in enumerate(stencil.input_names):
        # output k of iteration n-1 is called <input k>_iter<n>
        tensor = stencil.tensors['%s_iter%d' % (in_name, n)]
        low, margin = stencil.valid_bounds(tensor)
        lo, hi = program.window_of(k, n)
        assert tuple(max(0, -x) for x in lo) == tuple(low), (name, n, k)
        assert tuple(max(0, x) for x in hi) == tuple(margin), (name, n, k)


@pytest.mark.parametrize('seed', rp.MULTI_SEEDS[::4])
def test_multi_output_programs_schedule_per_output(seed, monkeypatch):
  """Both kernel families, executed on identities by the CPU models, leave
  every output right on its own box and 0 outside it."""
  import reg_schedule_sim as reg_sim
  import schedule_sim as sim
  monkeypatch.setenv('SODA_CUDA_TUNED', '0')
  stencil = rp.multi_stencil(seed)
  program = plan.extract_program(stencil)
  dims = (150, 31) if program.dim == 2 else (70, 24, 13)
  for style in ('reg', 'ring'):
    try:
      schedules = codegen.make_schedules(program, codegen.Options(style=style))
    except Exception as e:   # pylint: disable=broad-except
      assert any(w in str(e) for w in ('halo', 'shared memory', 'tile',
                                       'register', 'pair')), str(e)
      continue
    for sched in schedules:
      runner = reg_sim if sched.style == 'reg' else sim
      outs = runner.run_schedule(sched, dims, 9)
      runner.check_outputs(sched, dims, outs)


def test_deeper_depth_that_fits_no_kernel_keeps_the_last_one():
  """the 2-D depth search must not fail the whole compile when a
  deeper depth fits neither kernel family."""
  text = ('kernel: wide\nburst width: 64\nunroll factor: 1\niterate: 2\n'
          'input float: a(32, *)\n'
          'output float: b(0, 0) = a(-150, 0) + a(150, 0) + a(0, 0)\n')
  program = plan.extract_program(core.Stencil.from_text(text))
  schedules = codegen.make_schedules(program)
  assert [s.depth for s in schedules] == [1]
