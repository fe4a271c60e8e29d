import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for sub in ('tests', 'oracle', 'soda-compiler_b200'):
  path = os.path.join(ROOT, sub)
  if path not in sys.path:
    sys.path.insert(0, path)


def pytest_configure(config):
  config.addinivalue_line('markers', 'gpu: needs a CUDA device (B200)')


def _cuda_device_count():
  """Devices the CUDA runtime sees (0 without a driver), without torch."""
  import ctypes
  for name in ('libcudart.so', 'libcudart.so.12',
               '/usr/local/cuda/lib64/libcudart.so'):
    try:
      runtime = ctypes.CDLL(name)
    except OSError:
      continue
    count = ctypes.c_int(0)
    if runtime.cudaGetDeviceCount(ctypes.byref(count)) != 0:
      return 0
    return count.value
  return 0


def pytest_collection_modifyitems(config, items):
  """A plain `pytest tests` on a machine without a GPU skips the gpu-marked
  tests.  When the run SELECTS them (`-m gpu`) nothing is skipped: on a GPU
  box a missing device must fail loudly, not pass as skipped."""
  import pytest
  selected = config.getoption('-m') or ''
  if 'gpu' in selected and 'not gpu' not in selected:
    return
  if _cuda_device_count() > 0:
    return
  skip = pytest.mark.skip(reason='no CUDA device (run with -m gpu on a B200)')
  for item in items:
    if 'gpu' in item.keywords:
      item.add_marker(skip)
