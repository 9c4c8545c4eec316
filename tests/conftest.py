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
  config.addinivalue_line("markers", "multigpu: gpu test that launches torchrun over >= 2 devices of the box")


def _cuda_devices():
  """Number of CUDA devices the built library can open (0: no library, no driver or no GPU)."""
  try:
    import ctypes
    from starks_b200 import _lib
    lib = _lib.load()
    n = 0
    for dev in range(16):
      ctx = ctypes.c_void_p()
      if lib.stk_init(dev, ctypes.byref(ctx)) != 0:
        break
      lib.stk_destroy(ctx)
      n += 1
    return n
  except Exception:
    return 0


def pytest_collection_modifyitems(config, items):
  """`gpu` tests are skipped (not failed) on a host without a CUDA device, so that a plain
  `pytest tests` on a CPU box separates regressions from a missing GPU.  `multigpu` tests need
  at least two devices."""
  gpu_items = [it for it in items if "gpu" in it.keywords or "multigpu" in it.keywords]
  if not gpu_items:
    return
  ndev = _cuda_devices()
  if ndev == 0 and (os.path.exists("/dev/nvidiactl") or os.path.exists("/dev/nvidia0")):
    return  # a GPU box whose library cannot open the device must FAIL loudly, not skip
  for it in gpu_items:
    if ndev == 0:
      it.add_marker(pytest.mark.skip(reason="needs a CUDA device (libstarks_b200 has no CPU fallback)"))
    elif "multigpu" in it.keywords and ndev < 2:
      it.add_marker(pytest.mark.skip(reason="needs at least two CUDA devices"))


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
