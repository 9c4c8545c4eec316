"""BASELINE config 2: standalone NTT / iNTT sweep 2^10..2^24 on one B200 (CUDA events, data
resident, best of 3 x 10 launches).  batch=1 and a batch that fills HBM-sized work (2^26 elements)."""
import sys
sys.path.insert(0, '.')
import torch
from starks_b200 import Engine
P = 2**256 - 351*2**32 + 1
eng = Engine(0)
stream = torch.cuda.Stream(); eng.set_stream(stream.cuda_stream)
def run(logn, cols, inverse, reps=10):
    N = 1 << logn
    w = pow(7, (P-1)//N, P)
    d_in = torch.randint(0, 2**31-1, (cols, N, 8), dtype=torch.int32, device='cuda')
    d_out = torch.empty_like(d_in)
    for _ in range(3): eng.ntt(d_in.data_ptr(), N, N, d_out.data_ptr(), N, N, cols, w, inverse=inverse)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            e0.record(stream)
            for _ in range(reps): eng.ntt(d_in.data_ptr(), N, N, d_out.data_ptr(), N, N, cols, w, inverse=inverse)
            e1.record(stream)
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1)/reps)
    return best
print("# log2N  batch  fwd_us  fwd_Melem/s  inv_us  inv_Melem/s")
for logn in range(10, 25):
    for cols in (1, max(1, (1 << 26) >> logn)):
        f = run(logn, cols, False); i = run(logn, cols, True)
        n = cols << logn
        print("%6d %6d %9.1f %10.0f %9.1f %10.0f" % (logn, cols, f*1e3, n/f/1e3, i*1e3, n/i/1e3), flush=True)
