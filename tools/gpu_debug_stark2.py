import sys, ctypes
sys.path.insert(0, '.'); sys.path.insert(0, 'oracle')
import numpy as np
import oracle as orc
from starks_b200 import Engine
from starks_b200.limbs import ints_to_limbs, int_to_limbs
P = orc.P_STARK
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 8
ext, w = 8, 2
N = steps * ext
sp = [{(0, 1): 1}, {(1, 0): 1, (0, 1): 1}]
witness = orc.computational_trace(P, [0, 1], steps, sp)
O = orc.StarkOracle(steps, ext, w, sp)
tp, ds, bs = O.intermediates(witness, [(0, 0, 0), (0, 1, 1)])
G2, last = O.G2, O.last_step_position
Pev = [orc.fft_1d(P, t, G2, order=N) for t in tp]
Cev = [[(Pev[j][(i + ext) % N] - orc.eval_step_poly_on_ints(P, sp[j], [Pev[0][i], Pev[1][i]])) % P for i in range(N)] for j in range(w)]
Ccoef = [orc.fft_1d(P, c, G2, inv=True, order=N) for c in Cev]
eng = Engine(0)
E = 32
d_p = eng.alloc(w * N * E).upload(np.stack([ints_to_limbs(c) for c in Pev]))
d_c = eng.alloc(w * N * E); d_cc = eng.alloc(w * N * E); d_d = eng.alloc(w * N * E)
h_out = np.asarray([0, 1, 1], dtype=np.uint32)
h_coef = ints_to_limbs([1, 1, 1])
h_exp = np.asarray([[0, 1], [0, 1], [1, 0]], dtype=np.uint8)
eng._check(eng.lib.stk_constraint_eval(eng.ctx, d_p.ptr, N, ext, w, N, h_out.ctypes.data, h_coef.ctypes.data, h_exp.ctypes.data, 3, d_c.ptr, N))
got = d_c.download((w, N, 8))
for j in range(w):
    g = orc.from_limbs(got[j]); print("Cev", j, g == Cev[j], [i for i in range(N) if g[i] != Cev[j][i]][:6])
eng.ntt(d_c.ptr, N, N, d_cc.ptr, N, N, w, G2, inverse=True)
got = d_cc.download((w, N, 8))
for j in range(w):
    g = orc.from_limbs(got[j]); print("Ccoef", j, g == Ccoef[j], "deg", max([i for i in range(N) if Ccoef[j][i]] or [-1]))
bad = ctypes.c_uint32(0)
last_l = int_to_limbs(last)
for j in range(w):
    eng._check(eng.lib.stk_quotient_z(eng.ctx, d_cc.at(j * N * E), N, steps, last_l.ctypes.data_as(ctypes.POINTER(ctypes.c_uint32)), d_d.at(j * N * E), ctypes.byref(bad)))
    g = orc.from_limbs(d_d.download((N, 8), byte_offset=j * N * E))
    want = ds[j] + [0] * (N - len(ds[j]))
    print("Dcoef", j, "bad", bad.value, g == want, [i for i in range(N) if g[i] != want[i]][:6], "len ds", len(ds[j]))
print("---- full")
got = orc.from_limbs(d_c.download((w, N, 8)).reshape(-1, 8))
for j in range(w):
    for i in (1, 2, 9):
        g = got[j * N + i]
        print(j, i, "got  %064x" % g)
        print(j, i, "want %064x" % Cev[j][i])
        print(j, i, "diff %064x" % ((g - Cev[j][i]) % P), "-diff %064x" % ((Cev[j][i] - g) % P))
