import os
import sys

import pytest

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def pytest_addoption(parser):
  parser.addoption("--runslow", action="store_true", default=False, help="also run the tests marked `slow` (or set AGB_RUN_SLOW=1)")


def pytest_configure(config):
  config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `pytest -m gpu`)")
  config.addinivalue_line("markers", "timeout: per-test limit (pytest-timeout; 900 s by default, see pytest_collection_modifyitems)")
  config.addinivalue_line("markers", "slow: multi-second CPU test (finite differences over whole networks); run with --runslow / AGB_RUN_SLOW=1")


def pytest_collection_modifyitems(config, items):
  if config.pluginmanager.hasplugin("timeout"):   # pytest-timeout: a blocked test (dead producer thread, lost rank) must end the run, not hang it
    for item in items:
      if item.get_closest_marker("timeout") is None:
        item.add_marker(pytest.mark.timeout(900, method="thread"))
  if not (config.getoption("--runslow") or os.environ.get("AGB_RUN_SLOW")):
    skip_slow = pytest.mark.skip(reason="slow test: --runslow / AGB_RUN_SLOW=1")
    for item in items:
      if "slow" in item.keywords:
        item.add_marker(skip_slow)
  try:
    import torch
    has_gpu = torch.cuda.is_available()
  except Exception:
    has_gpu = False
  if has_gpu:
    return
  skip = pytest.mark.skip(reason="no CUDA device")
  for item in items:
    if "gpu" in item.keywords:
      item.add_marker(skip)
