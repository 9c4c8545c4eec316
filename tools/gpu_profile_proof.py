"""Full 2^20-step Fibonacci proof, twice (the second is the profiled one)."""
import sys, time
sys.path.insert(0, '.')
import numpy as np
from starks_b200 import Engine
from starks_b200.limbs import ints_to_limbs
from starks_b200.modp import IntegersModP
from starks_b200.stark import STARK
P = 2**256 - 351*2**32 + 1
steps = 1 << 20
a, b, c0, c1 = 0, 1, [], []
for _ in range(steps):
    c0.append(a); c1.append(b); a, b = b, (a + b) % P
witness = np.stack([ints_to_limbs(c0), ints_to_limbs(c1)])
eng = Engine(0)
S = STARK(IntegersModP(P), steps, 8, 2, [{(0, 1): 1}, {(1, 0): 1, (0, 1): 1}], engine=eng)
for i in range(3):
    t0 = time.time(); proof = S.mk_proof(witness, [(0, 0, 0), (0, 1, 1)]); print("proof", i, time.time() - t0, {k: round(v, 2) for k, v in S.timings.items()})
