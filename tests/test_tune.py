"""The autotuner's host logic (soda/cuda_tune.py, codegen/cuda/tuned.py): the
candidate space, the choice, the tuned table and its use by the planner.  The
GPU timing itself is replaced by a stand-in here; tests/test_tune_gpu.py runs
the real thing.
"""
import json

import pytest

import common
from soda import cuda_tune
from soda.codegen import cuda as codegen
from soda.codegen.cuda import plan, tuned


@pytest.mark.parametrize('name,iterate', [
    ('blur', 1), ('jacobi2d', 64), ('seidel2d', 2), ('denoise2d', 1),
    ('heat3d', 32), ('denoise3d', 1)])
def test_candidates_are_plannable_and_distinct(name, iterate, monkeypatch):
  monkeypatch.setenv('SODA_CUDA_TUNED', '0')
  program = plan.extract_program(common.stencil(name, iterate))
  sets = cuda_tune.candidates(program)
  assert sets and sets[0] == {}            # the planner's own choice first
  assert len(sets) >= 3
  kernels = set()
  for options in sets:
    schedules = codegen.make_schedules(program, codegen.Options(**options))
    assert all(codegen.layout_of(s).total <= codegen.SMEM_LIMIT
               for s in schedules)
    kernels.add(tuple(s.describe() + str(s.min_blocks) for s in schedules))
  assert len(kernels) == len(sets)         # no two resolve to the same kernels
  if program.feedback and iterate > 1:
    assert any('depth' in options for options in sets)


def test_tuned_table_steers_the_planner(tmp_path, monkeypatch):
  table = tmp_path / 'tuned.json'
  monkeypatch.setattr(tuned, 'TABLE_PATH', str(table))
  monkeypatch.delenv('SODA_CUDA_TUNED', raising=False)
  program = plan.extract_program(common.stencil('jacobi2d', 64))
  planner = codegen.make_schedules(program)[0]
  assert planner.depth == 8
  cuda_tune.record(program, (4096, 4096), 1.25, {'depth': 4, 'threads': 64},
                   'test device')
  entry = json.loads(table.read_text())[tuned.signature(program)]
  assert entry['options'] == {'depth': 4, 'threads': 64}
  assert entry['gcell_per_s'] == round(4096 * 4096 * 64 / 1.25 / 1e6, 1)
  tuned_sched = codegen.make_schedules(program)[0]
  assert (tuned_sched.depth, tuned_sched.threads) == (4, 64)
  # while tuning, an empty option set is the planner's choice again
  with cuda_tune.untuned():
    assert codegen.make_schedules(program)[0].depth == 8
  assert codegen.make_schedules(program)[0].depth == 4
  assert {'depth': 4} in cuda_tune.candidates(program)   # distinct from {}
  # explicit options win over the table; the switch turns it off
  assert codegen.make_schedules(
      program, codegen.Options(depth=2))[0].depth == 2
  monkeypatch.setenv('SODA_CUDA_TUNED', '0')
  assert codegen.make_schedules(program)[0].depth == 8
  monkeypatch.delenv('SODA_CUDA_TUNED')
  # another iteration count is another program: no entry
  other = plan.extract_program(common.stencil('jacobi2d', 16))
  assert tuned.signature(other) != tuned.signature(program)
  assert tuned.lookup(other) is None


def test_tune_picks_the_fastest_and_discards_wrong_results(monkeypatch):
  stencil = common.stencil('blur', 1)
  option_sets = [{}, {'threads': 64}, {'threads': 256}, {'prefetch': 36}]
  speed = {'{}': 3.0, "{'threads': 64}": 2.0, "{'threads': 256}": 1.0,
           "{'prefetch': 36}": 2.5}

  class FakeLibrary:
    def __init__(self, options):
      self.options = options

    def release(self):
      pass
  monkeypatch.setattr(cuda_tune, 'build_all', lambda st, sets, jobs=8: [
      (o, RuntimeError('nvcc failed') if o == {'prefetch': 36} else str(o))
      for o in sets])
  monkeypatch.setattr(cuda_tune.soda_cuda, 'load',
                      lambda path: FakeLibrary(path))

  def measure(library, dims, reps):
    # the fastest candidate computes something else: it must not win
    outputs = ['other' if library.options == "{'threads': 256}" else 'same']
    return speed[library.options], outputs
  lines = []
  results = cuda_tune.tune(stencil, (64, 64), option_sets, measure=measure,
                           log=lines.append)
  assert results == [(2.0, {'threads': 64}), (3.0, {})]
  assert any('DISCARDED' in line for line in lines)
  assert any('build failed' in line for line in lines)


def test_an_entry_may_tune_the_fast_build_separately(tmp_path, monkeypatch):
  """``options_fast``: the fast-math kernels of a program have other register
  needs than the exact ones (shipped: denoise3d)."""
  table = tmp_path / 'tuned.json'
  monkeypatch.setattr(tuned, 'TABLE_PATH', str(table))
  monkeypatch.delenv('SODA_CUDA_TUNED', raising=False)
  program = plan.extract_program(common.stencil('jacobi2d', 64))
  cuda_tune.record(program, (4096, 4096), 1.25, {'depth': 4}, 'test device')
  data = json.loads(table.read_text())
  data[tuned.signature(program)]['options_fast'] = {'depth': 2}
  table.write_text(json.dumps(data))
  assert codegen.make_schedules(program)[0].depth == 4
  assert codegen.make_schedules(program, fast_math=True)[0].depth == 2
  assert tuned.lookup(program) == {'depth': 4}
  assert tuned.lookup(program, fast_math=True) == {'depth': 2}


def test_shipped_denoise3d_entry_distinguishes_the_builds(monkeypatch):
  monkeypatch.delenv('SODA_CUDA_TUNED', raising=False)
  program = plan.extract_program(common.stencil('denoise3d'))
  exact = codegen.make_schedules(program)[0]
  fast = codegen.make_schedules(program, fast_math=True)[0]
  assert (tuple(exact.tile), exact.threads) == ((128, 28), 896)
  assert (tuple(fast.tile), fast.threads) == ((128, 32), 512)
