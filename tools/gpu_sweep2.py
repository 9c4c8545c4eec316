"""Tuning sweep 2: LDE-shaped transforms and register-capped variants."""
import os, sys
sys.path.insert(0, '.')
import torch
from starks_b200 import Engine
P = 2**256 - 351*2**32 + 1
eng = Engine(0)
stream = torch.cuda.Stream(); eng.set_stream(stream.cuda_stream)
def run(logn, cols, n_in=None, reps=5):
    N = 1 << logn
    n_in = n_in or N
    w = pow(7, (P-1)//N, P)
    d_in = torch.randint(0, 2**31-1, (cols, n_in, 8), dtype=torch.int32, device='cuda')
    d_out = torch.empty((cols, N, 8), dtype=torch.int32, device='cuda')
    for _ in range(2): eng.ntt(d_in.data_ptr(), n_in, n_in, d_out.data_ptr(), N, N, cols, w)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        e0.record(stream)
        for _ in range(reps): eng.ntt(d_in.data_ptr(), n_in, n_in, d_out.data_ptr(), N, N, cols, w)
        e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)/reps
    del d_in, d_out
    return ms, cols*N/ms/1e3
for (logn, cols, n_in) in ((20, 64, None), (21, 64, 1 << 18), (18, 64, None), (23, 6, 1 << 20)):
    for (logt, kmax, minb) in ((10, 10, 4), (10, 10, 5), (10, 10, 6), (11, 11, 4), (12, 11, 4), (10, 7, 4), (10, 8, 4), (9, 9, 4)):
        os.environ['STK_NTT_LOGT']=str(logt); os.environ['STK_NTT_KMAX']=str(kmax); os.environ['STK_NTT_MINB']=str(minb)
        ms, rate = run(logn, cols, n_in)
        print("logn=%d cols=%d n_in=%s logT=%d kmax=%d minb=%d : %.3f ms  %.0f Melem/s" % (logn, cols, n_in, logt, kmax, minb, ms, rate), flush=True)
