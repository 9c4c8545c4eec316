"""Tuning sweep of the NTT plan knobs on the GPU (not a test)."""
import os, sys, time
sys.path.insert(0, '.')
import numpy as np, torch
from starks_b200 import Engine
P = 2**256 - 351*2**32 + 1
eng = Engine(0)
stream = torch.cuda.Stream(); eng.set_stream(stream.cuda_stream)
def run(logn, cols, reps=5):
    N = 1 << logn
    w = pow(7, (P-1)//N, P)
    d_in = torch.randint(0, 2**31-1, (cols, N, 8), dtype=torch.int32, device='cuda')
    d_out = torch.empty_like(d_in)
    for _ in range(2): eng.ntt(d_in.data_ptr(), N, N, d_out.data_ptr(), N, N, cols, w)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        e0.record(stream)
        for _ in range(reps): eng.ntt(d_in.data_ptr(), N, N, d_out.data_ptr(), N, N, cols, w)
        e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)/reps
    return ms, cols*N/ms/1e3
configs = [(11,3,11),(10,3,11),(10,3,7),(9,3,9),(9,3,7),(8,3,8),(8,3,7),(10,3,5)]
for logn, cols in ((20,64),):
    for (logt, radix, kmax) in configs:
        os.environ['STK_NTT_LOGT']=str(logt); os.environ['STK_NTT_RADIX']=str(radix); os.environ['STK_NTT_KMAX']=str(kmax)
        try:
            ms, rate = run(logn, cols)
            print("logn=%d cols=%d logT=%d radix=%d kmax=%d : %.3f ms  %.0f Melem/s" % (logn, cols, logt, radix, kmax, ms, rate), flush=True)
        except Exception as ex:
            print("logn=%d logT=%d radix=%d kmax=%d FAILED %s" % (logn, logt, radix, kmax, ex), flush=True)
os.environ['STK_NTT_LOGT']='10'; os.environ['STK_NTT_RADIX']='3'; os.environ['STK_NTT_KMAX']='11'
for logn, cols in ((10,4096),(12,2048),(14,1024),(16,256),(18,128),(22,16),(24,4),(20,1),(24,1)):
    ms, rate = run(logn, cols)
    print("default plan logn=%d cols=%d : %.3f ms  %.0f Melem/s" % (logn, cols, ms, rate), flush=True)
