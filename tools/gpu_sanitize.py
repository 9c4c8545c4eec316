"""Small end-to-end pass over every kernel family, for compute-sanitizer (memcheck / racecheck)."""
import sys
sys.path.insert(0, '.')
import numpy as np
from starks_b200 import Engine
from starks_b200.modp import IntegersModP
from starks_b200.stark import STARK
from starks_b200.fri import FRI
P = 2**256 - 351*2**32 + 1
eng = Engine(0)
rng = np.random.default_rng(0)
def cols(b, n):
    a = rng.integers(0, 2**32, size=(b, n, 8), dtype=np.uint64).astype(np.uint32); a[:, :, 7] &= 0x7FFFFFFF; return a
for logn, b in ((3, 2), (6, 3), (10, 3), (13, 2), (14, 1)):
    n = 1 << logn; w = pow(7, (P-1)//n, P)
    x = cols(b, n); y = eng.ntt_host(x, n, w); z = eng.ntt_host(y, n, w, inverse=True); assert (z == x).all()
eng.set_field(31); eng.ntt_host(np.zeros((1, 4, 8), np.uint32), 6, pow(3, 5, 31)); eng.set_field(P)
F = IntegersModP(P)
sp = [{(0, 1): 1}, {(1, 0): 1, (0, 2): 1}]
steps = 64
tr = [[2], [3]]
for _ in range(steps - 1):
    a, b = tr[0][-1], tr[1][-1]; tr[0].append(b); tr[1].append((a + b * b) % P)
S = STARK(F, steps, 8, 2, sp, engine=eng)
proof = S.mk_proof(tr, [(0, 0, 2), (0, 1, 3)])
assert S.verify_proof(proof, tr, [(0, 0, 2), (0, 1, 3)])
prf = FRI(F, engine=eng).generate_proximity_proof(list(range(1, 129)), F(pow(7, (P-1)//1024, P)), 128, exclude_multiples_of=8)
assert len(prf) == 3
print("sanitize workload ok")
