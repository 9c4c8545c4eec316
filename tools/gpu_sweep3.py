import os, sys
sys.path.insert(0, '.')
import torch
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
    return e0.elapsed_time(e1)/reps
for (radix, minb) in ((3, 4), (2, 2), (2, 3), (2, 4), (3, 4)):
    os.environ['STK_NTT_RADIX']=str(radix); os.environ['STK_NTT_MINB']=str(minb)
    for logn, cols in ((20, 64), (12, 2048)):
        ms = run(logn, cols)
        print("radix=%d minb=%d logn=%d cols=%d: %.3f ms %.0f Melem/s" % (radix, minb, logn, cols, ms, cols*(1<<logn)/ms/1e3), flush=True)
