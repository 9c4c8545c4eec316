import sys, time, cProfile, pstats, io
sys.path.insert(0, '.')
import numpy as np, torch
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
mode = sys.argv[1]
if mode in ("stream", "both"):
    stream = torch.cuda.Stream(); eng.set_stream(stream.cuda_stream)
if mode in ("mem", "both"):
    big = torch.empty((4 << 30,), dtype=torch.uint8, device="cuda")
    pin = eng.pinned((64, 1 << 20, 8))
S = STARK(IntegersModP(P), steps, 8, 2, [{(0, 1): 1}, {(1, 0): 1, (0, 1): 1}], engine=eng)
S.mk_proof(witness, [(0, 0, 0), (0, 1, 1)])
pr = cProfile.Profile(); pr.enable()
S.mk_proof(witness, [(0, 0, 0), (0, 1, 1)])
pr.disable()
print(mode, 1, {k: round(v, 1) for k, v in S.timings.items()}, flush=True)
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(30); print(s.getvalue()[:2500])
