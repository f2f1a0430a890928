import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
  sys.path.insert(0, ROOT)


def pytest_configure(config):
  config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on a B200)')


def has_gpu() -> bool:
  try:
    import torch
    return torch.cuda.is_available()
  except Exception:  # pylint: disable=broad-except
    return False


def pytest_collection_modifyitems(config, items):
  if has_gpu():
    return
  skip = pytest.mark.skip(reason='no CUDA device in this container')
  for item in items:
    if 'gpu' in item.keywords:
      item.add_marker(skip)
