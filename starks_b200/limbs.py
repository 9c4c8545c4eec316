"""Conversions between Python ints and the ABI's element layout (8 little-endian uint32
limbs, canonical residue; include/starks_b200.h)."""
import numpy as np


def ints_to_limbs(ints) -> np.ndarray:
  """iterable of ints in [0, 2^256) -> (n, 8) uint32."""
  buf = b"".join(int(x).to_bytes(32, "little") for x in ints)
  return np.frombuffer(buf, dtype="<u4").reshape(-1, 8).copy()


def limbs_to_ints(arr) -> list:
  b = np.ascontiguousarray(arr, dtype="<u4").tobytes()
  return [int.from_bytes(b[i:i + 32], "little") for i in range(0, len(b), 32)]


def int_to_limbs(x: int) -> np.ndarray:
  return np.frombuffer(int(x).to_bytes(32, "little"), dtype="<u4").copy()


def limbs_to_be_bytes(arr) -> np.ndarray:
  """(n, 8) uint32 limbs -> (n, 32) uint8, the 32-byte big-endian form of
  IntegerModP.to_bytes (starks/modp.py:94-95)."""
  a = np.ascontiguousarray(arr, dtype="<u4").reshape(-1, 8)
  return a[:, ::-1].astype(">u4").view(np.uint8).reshape(-1, 32)


def be_bytes_to_limbs(b: np.ndarray) -> np.ndarray:
  a = np.ascontiguousarray(b, dtype=np.uint8).reshape(-1, 32)
  return a.view(">u4").astype("<u4")[:, ::-1].copy()
