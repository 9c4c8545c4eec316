"""starks_b200 -- B200-native drop-in for the STARK prover hot path of computablelabs/starks
(NTT / LDE over the modp field, BLAKE2s Merkle commitment, FRI folding).

Host side: Python mirrors of the reference modules (same names and semantics) over the C ABI
of libstarks_b200.so.  No CPU fallback: using an engine without the built library or without
a CUDA device raises."""
from ._lib import StarksB200Error, LIB_PATH  # noqa: F401
from .engine import Engine, default_engine, P_STARK  # noqa: F401

__all__ = ["Engine", "default_engine", "StarksB200Error", "P_STARK", "LIB_PATH"]
