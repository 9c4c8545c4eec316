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


def test_host_fiat_shamir_indices_match_the_reference_rule():
  """stk_pseudorandom_indices needs no device: the library's host BLAKE2s chain and index rule
  (used inside stk_fri_prove) against the Python mirror of starks/utils.py:60-90."""
  import ctypes
  import hashlib
  import numpy as np
  from starks_b200 import _lib
  from starks_b200.utils import get_pseudorandom_indices
  lib = _lib.load()
  for seed, modulus, count, excl in ((b"a", 1 << 21, 40, 8), (b"b", 512, 80, 0), (b"c", (1 << 24) - 1, 7, 3),
                                     (b"d", 32, 200, 4), (b"e", 2048, 1, 0)):
    entropy = hashlib.blake2s(seed).digest()
    out = np.zeros(count, dtype=np.uint64)
    buf = (ctypes.c_uint8 * 32).from_buffer_copy(entropy)
    assert lib.stk_pseudorandom_indices(None, buf, modulus, count, excl, out.ctypes.data) == 0
    assert out.tolist() == get_pseudorandom_indices(entropy, modulus, count, exclude_multiples_of=excl)
  out = np.zeros(4, dtype=np.uint64)
  assert lib.stk_pseudorandom_indices(None, buf, 1 << 24, 4, 0, out.ctypes.data) != 0   # utils.py:69 assert
