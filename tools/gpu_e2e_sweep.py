import os, sys, time
sys.path.insert(0, '.')
import numpy as np
from starks_b200 import Engine
P = 2**256 - 351*2**32 + 1
eng = Engine(0)
N, cols = 1 << 20, 64
w = pow(7, (P-1)//N, P)
hin = eng.pinned((cols, N, 8)); hout = eng.pinned((cols, N, 8))
rng = np.random.default_rng(0)
hin.array[...] = rng.integers(0, 2**31, size=(cols, N, 8), dtype=np.int64).astype(np.uint32)
for mb in (32, 64, 128, 256, 512):
    os.environ['STK_HOST_CHUNK_MB'] = str(mb)
    eng.ntt_host(hin.array, N, w, out=hout.array)
    t0 = time.perf_counter()
    for _ in range(3): eng.ntt_host(hin.array, N, w, out=hout.array)
    dt = (time.perf_counter() - t0) / 3
    print("chunk %d MiB: %.2f ms  %.0f Melem/s  (%.1f GB/s each way)" % (mb, dt*1e3, cols*N/dt/1e6, cols*N*32/dt/1e9), flush=True)
