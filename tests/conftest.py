import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for sub in ('tests', 'oracle', 'soda-compiler_b200'):
  path = os.path.join(ROOT, sub)
  if path not in sys.path:
    sys.path.insert(0, path)


def pytest_configure(config):
  config.addinivalue_line('markers', 'gpu: needs a CUDA device (B200)')
