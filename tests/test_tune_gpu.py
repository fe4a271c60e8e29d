"""The autotuner on a real GPU: candidates are timed, all of them compute the
same bits, and the winner is one of them."""
import pytest

import common
from soda import cuda_tune

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('name,iterate,dims,option_sets', [
    ('blur', 1, (2048, 600), [{}, {'threads': 64}, {'prefetch': 36}]),
    ('jacobi2d', 64, (2048, 700), [{}, {'depth': 4}, {'depth': 2}]),
    ('heat3d', 32, (256, 96, 80), [{}, {'tile': [128, 16]}, {'depth': 1}]),
])
def test_tune_times_every_candidate(name, iterate, dims, option_sets,
                                    monkeypatch):
  monkeypatch.setenv('SODA_CUDA_TUNED', '0')
  lines = []
  results = cuda_tune.tune(common.stencil(name, iterate), dims, option_sets,
                           reps=3, log=lines.append)
  assert len(results) == len(option_sets), lines   # none failed or differed
  assert all(ms > 0 for ms, _ in results)
  assert [ms for ms, _ in results] == sorted(ms for ms, _ in results)
  assert not any('DISCARDED' in line for line in lines)
