import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
  sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
  config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
  config.addinivalue_line("markers", "slow: longer CPU test")


def load_golden(name):
  import json
  with open(os.path.join(GOLDEN, name)) as fh:
    return json.load(fh)


@pytest.fixture(scope="session")
def oracle():
  sys.path.insert(0, os.path.join(ROOT, "oracle"))
  import oracle as orc
  orc.build()
  return orc
