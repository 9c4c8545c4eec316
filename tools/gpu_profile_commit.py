"""One LDE + Merkle commit of BASELINE config 3 (64 x 2^18 -> 2^21), three times: the profiled
launches are the last call's (merkle_leaf_pairs_cols_kernel, the pass kernels)."""
import sys
sys.path.insert(0, '.')
import torch
from starks_b200 import Engine
P = 2**256 - 351 * 2**32 + 1
steps, ext, ncols = 1 << 18, 8, 64
n = steps * ext
g2 = pow(7, (P - 1) // n, P)
eng = Engine(0)
tr = torch.randint(0, 2**31 - 1, (ncols, steps, 8), dtype=torch.int32, device='cuda')
ev = torch.empty((ncols, n, 8), dtype=torch.int32, device='cuda')
nodes = torch.empty((n, 32), dtype=torch.uint8, device='cuda')
torch.cuda.synchronize()
for i in range(3):
    print(eng.lde_commit(tr.data_ptr(), steps, steps, ext, ncols, g2, ev.data_ptr(), n, nodes.data_ptr()).hex())
