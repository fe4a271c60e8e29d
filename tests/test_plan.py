"""The streaming schedule, checked on the CPU by executing it on identities.

schedule_sim.run_schedule runs a plan.Schedule block by block with the index
arithmetic of the generated kernel (ring slots, wrap-around plane offsets,
overlapping tiles, chunk lead-in); a cell keeps its identity only if computed
from exactly the prescribed operands.  Every cell of the reference's valid
region must come out right, everything else must be 0, nothing may be stored
twice or left unwritten.
"""
import math

import pytest
from hypothesis import given, settings, strategies as st

import common
import schedule_sim as sim
from soda import core
from soda.codegen import cuda as codegen
from soda.codegen.cuda import plan

CASES = [
    ('blur', 1, (32,), 4, (70, 37), 16), ('blur', 1, (64,), 8, (130, 21), 7),
    ('sobel2d', 1, (32,), 8, (50, 41), 16),
    ('jacobi2d', 1, (32,), 4, (75, 40), 13),
    ('jacobi2d', 3, (32,), 4, (75, 40), 13),
    ('jacobi2d', 5, (64,), 2, (100, 30), 30),
    ('seidel2d', 2, (64,), 4, (100, 33), 11),
    ('denoise2d', 1, (32,), 4, (61, 29), 8),
    ('heat3d', 1, (16, 8), 4, (21, 13, 17), 6),
    ('heat3d', 2, (16, 8), 4, (21, 13, 17), 6),
    ('jacobi3d', 2, (32, 8), 4, (40, 13, 9), 9),
    ('denoise3d', 1, (16, 8), 2, (19, 11, 13), 5),
]


@pytest.mark.parametrize('name,depth,tile,vec,dims,chunk', CASES)
def test_schedule_produces_the_valid_region(name, depth, tile, vec, dims,
                                            chunk):
  program = plan.extract_program(common.stencil(name, depth))
  sched = plan.Schedule(program, depth, tile, vec, math.prod(tile) // vec,
                        prefetch=2)
  outs = sim.run_schedule(sched, dims, chunk)
  sim.check_outputs(sched, dims, outs)


def test_non_final_launch_stores_everything():
  program = plan.extract_program(common.stencil('jacobi2d', 2))
  sched = plan.Schedule(program, 2, (32,), 4, 8, prefetch=1)
  outs = sim.run_schedule(sched, (50, 20), 9, final=False)
  sim.check_outputs(sched, (50, 20), outs, final=False)


def test_window_matches_reference_bounds():
  """plan.Program.window (bounding boxes) agrees with the exact window of the
  Stencil IR, which tests/test_frontend.py pins to the reference."""
  for name in common.BENCHMARKS:
    for iterate in (1, 2, 5):
      try:
        stencil = common.stencil(name, iterate)
      except Exception:      # denoise: iterate > 1 is rejected
        continue
      program = plan.extract_program(stencil)
      lo, hi = program.window(iterate)
      for out_name in stencil.output_names:
        low, margin = stencil.valid_bounds(stencil.tensors[out_name])
        assert tuple(max(0, -l) for l in lo) == tuple(low)
        assert tuple(max(0, h) for h in hi) == tuple(margin)


def test_default_schedules_fit_shared_memory():
  from soda.codegen.cuda import kernel
  for name, iterate in [(n, None) for n in common.BENCHMARKS] + [
      ('jacobi2d', 64), ('heat3d', 32)]:
    program = plan.extract_program(common.stencil(name, iterate))
    schedules = codegen.make_schedules(program)
    assert sum(s.depth for s in schedules) <= program.iterate or True
    for sched in schedules:
      assert kernel.Layout(sched).total <= codegen.SMEM_LIMIT
    depths = [s.depth for s in schedules]
    assert program.iterate % depths[0] == (depths[1] if len(depths) > 1 else 0)


# --- random programs -----------------------------------------------------------

@st.composite
def random_program(draw):
  dim = draw(st.integers(2, 3))
  n_local = draw(st.integers(0, 3))
  names = ['a']
  lines = ['kernel: rnd', 'burst width: 64', 'unroll factor: 1',
           'input float: a(%s*)' % ''.join('8, ' for _ in range(dim - 1))]
  zero = ', '.join('0' for _ in range(dim))
  for k in range(n_local + 1):
    target = 'l%d' % k if k < n_local else 'out'
    terms = []
    for _ in range(draw(st.integers(1, 3))):
      parent = draw(st.sampled_from(names))
      off = [draw(st.integers(-2, 2)) for _ in range(dim)]
      terms.append('%s(%s)' % (parent, ', '.join(map(str, off))))
    # every window must contain the store point (Program.check_windows)
    terms.append('%s(%s)' % (names[-1], zero))
    lines.append('%s float: %s(%s) = %s' % (
        'local' if k < n_local else 'output', target, zero, ' + '.join(terms)))
    names.append(target)
  iterate = draw(st.integers(1, 3))
  lines.append('iterate: %d' % iterate)
  return '\n'.join(lines) + '\n', dim, iterate


@settings(max_examples=25, deadline=None)
@given(random_program(), st.integers(1, 2))
def test_random_programs_schedule_correctly(generated, prefetch):
  text, dim, iterate = generated
  stencil = core.Stencil.from_text(text)
  program = plan.extract_program(stencil)
  program.check_windows()
  tile = (64,) if dim == 2 else (32, 16)
  dims = (90, 23) if dim == 2 else (41, 21, 11)
  try:
    sched = plan.Schedule(program, iterate, tile, 4, math.prod(tile) // 4,
                          prefetch=prefetch)
  except Exception as e:   # halo larger than the tile: a legal refusal
    assert 'halo' in str(e)
    return
  outs = sim.run_schedule(sched, dims, 8)
  sim.check_outputs(sched, dims, outs)


# --- the register-streaming family (the default kernels) ----------------------

REG_CASES = [
    ('blur', 1, (300, 23)), ('sobel2d', 1, (300, 23)),
    ('jacobi2d', 1, (300, 23)), ('jacobi2d', 8, (300, 60)),   # paired f32x2
    ('jacobi2d', 6, (300, 47)), ('jacobi2d', 3, (300, 30)),   # odd: unpaired
    ('seidel2d', 2, (300, 23)), ('denoise2d', 1, (300, 23)),
    ('heat3d', 2, (140, 40, 12)), ('heat3d', 1, (140, 40, 9)),
    ('jacobi3d', 2, (140, 40, 12)), ('denoise3d', 1, (140, 20, 9)),
]


@pytest.mark.parametrize('name,depth,dims', REG_CASES)
def test_register_schedule_produces_the_valid_region(name, depth, dims,
                                                     monkeypatch):
  """reg_schedule_sim executes plan.RegSchedule — register histories, warp
  shuffles, shared planes for y neighbours, paired lanes, trips — as the
  emitter lays it out."""
  import reg_schedule_sim as reg_sim
  monkeypatch.setenv('SODA_CUDA_TUNED', '0')
  program = plan.extract_program(common.stencil(name, depth))
  sched = codegen.make_schedules(program, codegen.Options(depth=depth))[0]
  assert sched.style == 'reg' and sched.depth == depth
  for chunk in (7, dims[-1]):
    outs = reg_sim.run_schedule(sched, dims, chunk)
    reg_sim.check_outputs(sched, dims, outs)


def test_register_schedule_sim_notices_a_wrong_schedule(monkeypatch):
  import reg_schedule_sim as reg_sim
  monkeypatch.setenv('SODA_CUDA_TUNED', '0')
  for name, depth, dims, attr in (('jacobi2d', 8, (300, 60), 'pair_lag'),
                                  ('blur', 1, (300, 23), 'lead')):
    program = plan.extract_program(common.stencil(name, depth))
    sched = codegen.make_schedules(program, codegen.Options(depth=depth))[0]
    setattr(sched, attr, getattr(sched, attr) + (1 if attr == 'pair_lag'
                                                 else -1))
    with pytest.raises(AssertionError):
      outs = reg_sim.run_schedule(sched, dims, 7)
      reg_sim.check_outputs(sched, dims, outs)


@settings(max_examples=20, deadline=None)
@given(random_program())
def test_random_programs_schedule_correctly_in_registers(generated):
  import reg_schedule_sim as reg_sim
  text, dim, iterate = generated
  stencil = core.Stencil.from_text(text)
  program = plan.extract_program(stencil)
  program.check_windows()
  dims = (300, 29) if dim == 2 else (140, 40, 13)
  try:
    sched = codegen.make_schedule(program, iterate,
                                  codegen.Options(style='reg'))
  except Exception as e:   # pylint: disable=broad-except
    # a legal refusal: halo larger than the tile, histories too large
    assert any(word in str(e) for word in ('halo', 'shared memory', 'tile',
                                           'register')), str(e)
    return
  assert sched.style == 'reg'
  outs = reg_sim.run_schedule(sched, dims, 9)
  reg_sim.check_outputs(sched, dims, outs)


@st.composite
def multi_io_program(draw):
  """One or two inputs and as many outputs (later outputs may read earlier
  ones), up to two locals, 1-3 iterations."""
  dim = draw(st.integers(2, 3))
  n_io = draw(st.integers(1, 2))
  n_local = draw(st.integers(0, 2))
  inputs = ['a', 'b'][:n_io]
  names = list(inputs)
  lines = ['kernel: rnd', 'burst width: 64', 'unroll factor: 1']
  for name in inputs:
    lines.append('input float: %s(%s*)' % (
        name, ''.join('8, ' for _ in range(dim - 1))))
  zero = ', '.join('0' for _ in range(dim))
  targets = ['l%d' % k for k in range(n_local)] + [
      'o%d' % k for k in range(n_io)]
  for k, target in enumerate(targets):
    terms = []
    for _ in range(draw(st.integers(1, 3))):
      parent = draw(st.sampled_from(names))
      off = [draw(st.integers(-2, 2)) for _ in range(dim)]
      terms.append('%s(%s)' % (parent, ', '.join(map(str, off))))
    # every window must contain the store point (Program.check_windows)
    terms += ['%s(%s)' % (name, zero) for name in [names[-1]] + inputs]
    lines.append('%s float: %s(%s) = %s' % (
        'local' if k < n_local else 'output', target, zero,
        ' + '.join(terms)))
    names.append(target)
  iterate = draw(st.integers(1, 3))
  lines.append('iterate: %d' % iterate)
  return '\n'.join(lines) + '\n', dim, iterate


@settings(max_examples=20, deadline=None)
@given(multi_io_program())
def test_random_multi_output_programs_schedule_correctly(generated):
  import os
  import reg_schedule_sim as reg_sim
  text, dim, iterate = generated
  program = plan.extract_program(core.Stencil.from_text(text))
  program.check_windows()
  dims = (300, 29) if dim == 2 else (140, 40, 13)
  before = os.environ.get('SODA_CUDA_TUNED')
  os.environ['SODA_CUDA_TUNED'] = '0'
  try:
    sched = codegen.make_schedule(program, iterate, codegen.Options())
  except Exception as e:   # pylint: disable=broad-except
    assert any(word in str(e) for word in ('halo', 'shared memory')), str(e)
    return
  finally:
    if before is None:
      del os.environ['SODA_CUDA_TUNED']
    else:
      os.environ['SODA_CUDA_TUNED'] = before
  runner = reg_sim if sched.style == 'reg' else sim
  outs = runner.run_schedule(sched, dims, 9)
  runner.check_outputs(sched, dims, outs)


@pytest.mark.parametrize('name', ['chain2', 'chain3d'])
def test_output_read_by_a_later_statement(name, monkeypatch):
  """Under iterate > 1 the Stencil IR renames output k of the first
  iteration `<input k>_iter1` (reference core.py:347-351), also inside the
  statements that read it; the program extraction must undo that."""
  import random_programs as rp
  import reg_schedule_sim as reg_sim
  monkeypatch.setenv('SODA_CUDA_TUNED', '0')
  stencil = rp.extra_stencil(name)
  program = plan.extract_program(stencil)
  readers = [s for s in program.stages if s.name == 'o1']
  assert any(load.parent == 'o0' for load in readers[0].loads)
  dims = (300, 31) if program.dim == 2 else (140, 40, 13)
  schedules = list(codegen.make_schedules(program))
  try:      # all iterations in one launch too, where that fits
    schedules.append(codegen.make_schedule(program, program.iterate,
                                           codegen.Options()))
  except Exception:   # pylint: disable=broad-except
    pass
  assert schedules
  for sched in schedules:
    runner = reg_sim if sched.style == 'reg' else sim
    outs = runner.run_schedule(sched, dims, 9)
    runner.check_outputs(sched, dims, outs)


def test_window_without_store_point_is_rejected():
  text = ('kernel: k\nburst width: 64\nunroll factor: 1\niterate: 1\n'
          'input float: a(8, *)\nlocal float: l(0, 0) = a(0, -1)\n'
          'output float: o(0, 0) = l(0, 1)\n')
  program = plan.extract_program(core.Stencil.from_text(text))
  with pytest.raises(Exception) as info:
    codegen.check_supported(program)
  assert 'window must include 0' in str(info.value)
