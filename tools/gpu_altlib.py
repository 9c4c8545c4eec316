"""A/B of an alternative kernel build (STARKS_B200_LIB): in-register multiply / butterfly rates,
NTT 64 x 2^20 time and a digest of its output (must match between builds)."""
import hashlib, os, sys
sys.path.insert(0, '.')
import torch
from starks_b200 import Engine
P = 2**256 - 351*2**32 + 1
eng = Engine(0)
print('lib', os.environ.get('STARKS_B200_LIB', 'default'))
for which, name in ((5, 'field_mul'), (6, 'butterfly')):
    best = 0
    for _ in range(3):
        ms, ops = eng.microbench(which, 1000)
        best = max(best, ops / (ms * 1e-3) / 1e9)
    print('%-10s %.1f G/s' % (name, best))
N, cols = 1 << 20, 64
w = pow(7, (P-1)//N, P)
g = torch.Generator(device='cuda'); g.manual_seed(1)
d_in = torch.randint(0, 2**31-1, (cols, N, 8), dtype=torch.int32, device='cuda', generator=g)
d_out = torch.empty_like(d_in)
stream = torch.cuda.Stream(); eng.set_stream(stream.cuda_stream)
for _ in range(3): eng.ntt(d_in.data_ptr(), N, N, d_out.data_ptr(), N, N, cols, w)
torch.cuda.synchronize()
best = 1e9
for _ in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        e0.record(stream)
        for _ in range(10): eng.ntt(d_in.data_ptr(), N, N, d_out.data_ptr(), N, N, cols, w)
        e1.record(stream)
    torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1) / 10)
print('ntt 64x2^20 %.3f ms' % best, 'digest', hashlib.blake2s(d_out[:2].cpu().numpy().tobytes()).hexdigest()[:16])
