"""Timing of the headline transforms under env-knob combinations (CUDA events, data resident).
usage: gpu_knobs.py "A=1,B=2" "A=3" ...   ('-' = defaults)"""
import os, sys
sys.path.insert(0, '.')
import torch
from starks_b200 import Engine
P = 2**256 - 351*2**32 + 1
eng = Engine(0)
stream = torch.cuda.Stream(); eng.set_stream(stream.cuda_stream)
def timeit(fn, reps=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            e0.record(stream)
            for _ in range(reps): fn()
            e1.record(stream)
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / reps)
    return best
bufs = {}
def ntt(logn, cols, inverse=False):
    N = 1 << logn
    w = pow(7, (P-1)//N, P)
    if (logn, cols) not in bufs:
        bufs.clear(); torch.cuda.empty_cache()
        d_in = torch.randint(0, 2**31-1, (cols, N, 8), dtype=torch.int32, device='cuda')
        bufs[(logn, cols)] = (d_in, torch.empty_like(d_in))
    d_in, d_out = bufs[(logn, cols)]
    return timeit(lambda: eng.ntt(d_in.data_ptr(), N, N, d_out.data_ptr(), N, N, cols, w, inverse=inverse))
def lde(logs, cols, ext=8):
    steps = 1 << logs; N = steps * ext
    g2 = pow(7, (P-1)//N, P)
    bufs.clear(); torch.cuda.empty_cache()
    tr = torch.randint(0, 2**31-1, (cols, steps, 8), dtype=torch.int32, device='cuda')
    ev = torch.empty((cols, N, 8), dtype=torch.int32, device='cuda')
    return timeit(lambda: eng.lde(tr.data_ptr(), steps, steps, ext, cols, g2, ev.data_ptr(), N))
def zp(logn, cols):
    """zero-padded forward transform: n_in = N/8 coefficients per column (the proof's batched evaluation)"""
    N = 1 << logn; nin = N >> 3
    w = pow(7, (P-1)//N, P)
    bufs.clear(); torch.cuda.empty_cache()
    d_in = torch.randint(0, 2**31-1, (cols, nin, 8), dtype=torch.int32, device='cuda')
    d_out = torch.empty((cols, N, 8), dtype=torch.int32, device='cuda')
    return timeit(lambda: eng.ntt(d_in.data_ptr(), nin, nin, d_out.data_ptr(), N, N, cols, w))
cases = [('zp23x6', lambda: zp(23, 6)), ('zp21x64', lambda: zp(21, 64)), ('zp16x512', lambda: zp(16, 512)), ('ntt20x64', lambda: ntt(20, 64)), ('intt20x64', lambda: ntt(20, 64, True)), ('ntt24x4', lambda: ntt(24, 4)),
         ('ntt16x1024', lambda: ntt(16, 1024)), ('ntt12x16384', lambda: ntt(12, 16384)), ('ntt23x6', lambda: ntt(23, 6)),
         ('lde18x64', lambda: lde(18, 64)), ('lde20x2', lambda: lde(20, 2))]
if os.environ.get('KNOB_CASES'):
    keep = os.environ['KNOB_CASES'].split(',')
    cases = [c for c in cases if c[0] in keep]
combos = sys.argv[1:] or ['-']
print('%-40s' % 'knobs' + ''.join('%13s' % c[0] for c in cases), flush=True)
for combo in combos:
    sets = [kv.split('=') for kv in combo.split(',')] if combo != '-' else []
    for k, v in sets: os.environ[k] = v
    row = [fn() for _, fn in cases]
    for k, _ in sets: del os.environ[k]
    print('%-40s' % combo + ''.join('%10.3f ms' % t for t in row), flush=True)
