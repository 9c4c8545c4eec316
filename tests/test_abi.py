"""CPU-side checks of the drop-in boundary: the shared library loads and exports every
symbol include/starks_b200.h declares; no compute calls (no GPU here)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
  text = open(os.path.join(ROOT, "include", "starks_b200.h")).read()
  text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
  return sorted(set(re.findall(r"\b(stk_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported():
  from starks_b200 import _lib
  if not os.path.exists(_lib.LIB_PATH):
    import __graft_entry__
    __graft_entry__.build()
  lib = ctypes.CDLL(_lib.LIB_PATH)
  syms = declared_symbols()
  assert len(syms) >= 20
  for s in syms:
    assert hasattr(lib, s), "library does not export %s" % s
  # the Python binding table covers exactly the header
  assert sorted(_lib.SIGNATURES) == syms
  assert lib.stk_version() >= 1


def test_no_cpu_fallback():
  """Without a CUDA device the engine must fail loudly (no silent CPU path)."""
  import torch
  if torch.cuda.is_available():
    pytest.skip("a GPU is present")
  from starks_b200 import Engine, StarksB200Error
  with pytest.raises(StarksB200Error):
    Engine(0)


def test_product_never_imports_oracle():
  pkg = os.path.join(ROOT, "starks_b200")
  for dirpath, _, files in os.walk(pkg):
    for f in files:
      if f.endswith((".py", ".cu", ".cuh", ".h")):
        src = open(os.path.join(dirpath, f)).read()
        assert "import oracle" not in src and "liboracle" not in src and "oracle/" not in src, f
