"""Single-use locals spliced into their readers (plan.inline_single_use).

The transformation must not change a single bit: a spliced local is its own
expression on the same operands in the same order, rounded through its
declared type.  Checked here on the CPU by evaluating both versions of a
program with tests/program_eval.py (periodic boundaries, every cell defined)
— benchmarks, random multi-output programs and hand-written corner cases:
a narrowing integer local, a chain of single-use locals, a local read once
at a non-zero offset.  The schedules of spliced programs run through the same
CPU schedule models as any other (tests/test_plan.py style) below; GPU
parity of spliced kernels: the `inline` cases of tests/test_parity_gpu.py.
"""
import numpy as np
import pytest

import common
import program_eval
import random_programs as rp
import reg_schedule_sim as reg_sim
import schedule_sim as sim
from soda import core
from soda.codegen import cuda as codegen
from soda.codegen.cuda import plan

HEADER = 'kernel: %s\nburst width: 64\nunroll factor: 1\niterate: 1\n'
CORNERS = {
    # int16 local narrows (wraps) before it is widened again
    'narrow': HEADER % 'narrow' + 'input int16: a(32, *)\n'
              'local int16: s(0, 0) = a(0, 0) * 300 + a(1, 0)\n'
              'output int32: o(0, 0) = s(0, 0) * 2 + a(0, 1)\n',
    # chain of single-use locals, the last read at a non-zero offset
    'chain': HEADER % 'chain' + 'input float: a(32, *)\n'
             'local float: p(0, 0) = a(0, 0) * 0.3f + a(1, 0)\n'
             'local float: q(0, 0) = p(0, 0) * p(0, 0) + a(0, 1)\n'
             'output float: o(0, 0) = q(1, -1) / 3 + a(0, 0)\n',
    # float local inside double arithmetic: the rounding to float must stay
    'mixed': HEADER % 'mixed' + 'input double: a(32, *)\n'
             'local float: m(0, 0) = a(0, 0) * 0.1 + a(0, 1)\n'
             'output double: o(0, 0) = m(0, 0) * 3.0 + a(1, 0)\n',
    # a local read by two statements stays; one read twice in ONE expression
    # at one offset goes
    'shared': HEADER % 'shared' + 'input float: a(32, *)\n'
              'local float: s(0, 0) = a(0, 0) + a(1, 0)\n'
              'local float: t(0, 0) = a(0, 0) - a(0, 1)\n'
              'output float: o0(0, 0) = s(0, 0) * t(0, 0) * t(0, 0)\n'
              'output float: o1(0, 0) = s(0, 1) + a(0, 0)\n',
}


def _inputs(program, dims, seed):
  import golden
  rng = np.random.default_rng(seed)
  shape = tuple(reversed(dims))
  arrays = []
  for _, haoda_type in program.inputs:
    dtype = np.dtype(golden.NUMPY_TYPES[haoda_type])
    if dtype.kind == 'f':
      arrays.append((rng.random(shape) * 4 - 1).astype(dtype))
    else:
      arrays.append(rng.integers(-300, 300, size=shape).astype(dtype))
  return arrays


def _programs():
  cases = [(name, common.stencil(name)) for name in
           ('denoise2d', 'denoise3d', 'sobel2d')]
  cases += [(name, core.Stencil.from_text(text))
            for name, text in sorted(CORNERS.items())]
  cases += [('multi%d' % seed, rp.multi_stencil(seed))
            for seed in rp.MULTI_SEEDS[:24]]
  return cases


def test_which_locals_are_spliced():
  spliced = lambda name: [s.name for s in plan.inline_single_use(
      plan.extract_program(core.Stencil.from_text(CORNERS[name]))).stages]
  assert spliced('narrow') == ['o']
  assert spliced('chain') == ['o']
  assert spliced('shared') == ['s', 'o0', 'o1']
  denoise = plan.inline_single_use(plan.extract_program(
      common.stencil('denoise3d')))
  assert [s.name for s in denoise.stages] == ['g', 'output']   # g: sqrt stays
  assert plan.inline_single_use(plan.extract_program(
      common.stencil('blur'))) is None


@pytest.mark.parametrize('name,stencil', _programs(),
                         ids=[n for n, _ in _programs()])
def test_spliced_program_computes_the_same_bits(name, stencil):
  program = plan.extract_program(stencil)
  spliced = plan.inline_single_use(program)
  if spliced is None:
    pytest.skip('nothing to splice')
  assert len(spliced.stages) < len(program.stages)
  # same windows: valid regions and halos do not move
  for k in range(len(program.outputs)):
    assert spliced.window_of(k, 1) == program.window_of(k, 1)
  dims = (24, 18) if program.dim == 2 else (12, 10, 9)
  inputs = _inputs(program, dims, seed=len(name))
  want = program_eval.evaluate(program, inputs)
  got = program_eval.evaluate(spliced, inputs)
  for g, w in zip(got, want):
    common.assert_bit_exact(g, w, name, any_nan=True)
    assert np.isfinite(w.astype(np.float64)).mean() > 0.9


@pytest.mark.parametrize('name', ['denoise2d', 'denoise3d', 'sobel2d'])
def test_spliced_schedules_run_on_the_cpu_models(name, monkeypatch):
  monkeypatch.setenv('SODA_CUDA_TUNED', '0')
  program = plan.extract_program(common.stencil(name))
  dims = (300, 23) if program.dim == 2 else (140, 20, 9)
  for inline in (True, False):
    for style in ('reg', 'ring'):
      sched = codegen.make_schedules(
          program, codegen.Options(inline=inline, style=style))[0]
      assert (len(sched.program.stages) < len(program.stages)) == inline
      runner = reg_sim if sched.style == 'reg' else sim
      outs = runner.run_schedule(sched, dims, 7)
      runner.check_outputs(sched, dims, outs)


def test_splicing_is_on_request_and_holds_fewer_registers(monkeypatch):
  monkeypatch.setenv('SODA_CUDA_TUNED', '0')
  program = plan.extract_program(common.stencil('denoise2d'))
  plain = codegen.make_schedules(program)[0]
  assert len(plain.program.stages) == len(program.stages)    # default: as written
  spliced = codegen.make_schedules(program, codegen.Options(inline=True))[0]
  assert len(spliced.program.stages) == 2
  assert codegen.history_registers(spliced) < codegen.history_registers(plain)
