import sys, hashlib
sys.path.insert(0, '.'); sys.path.insert(0, 'oracle')
import numpy as np
import oracle as orc
from starks_b200.modp import IntegersModP
import starks_b200.stark as st
P = orc.P_STARK
F = IntegersModP(P)
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 8
sp = [{(0, 1): 1}, {(1, 0): 1, (0, 1): 1}]
witness = orc.computational_trace(P, [0, 1], steps, sp)
boundary = [(0, 0, 0), (0, 1, 1)]
O = orc.StarkOracle(steps, 8, 2, sp)
want, inter = O.mk_proof(witness, boundary, return_intermediates=True)
S = st.STARK(F, steps, 8, 2, sp)
got = S.mk_proof(witness, boundary, keep_device=True)
N = steps * 8
cols = S.device["cols"].download((6, N, 8))
names = ["P1", "P2", "D1", "D2", "B1", "B2"]
for k in range(6):
    g = orc.from_limbs(cols[k])
    ok = g == inter["evals"][k]
    print(names[k], "OK" if ok else "MISMATCH", "" if ok else [i for i in range(N) if g[i] != inter["evals"][k][i]][:8])
pc = orc.from_limbs(S.device["pcoef"].download((2, steps, 8)).reshape(-1, 8))
print("pcoef", pc[:steps] == inter["trace_polys"][0] + [0] * (steps - len(inter["trace_polys"][0])))
l = orc.from_limbs(S.device["l"].download((N, 8)))
print("l evals", l == inter["l_evals"])
print("roots", got[0] == want[0], got[1] == want[1], "proof", got == want)
