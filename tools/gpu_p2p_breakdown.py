"""Per-phase timing of the two four-step variants (torchrun, N GPUs)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from starks_b200 import Engine
from starks_b200 import dist as sd
P = 2**256 - 351 * 2**32 + 1
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
eng = Engine(local)
n = 1 << 26; L = n // world
w = pow(7, (P - 1) // n, P)
x = torch.randint(0, 2**31 - 1, (L, 8), dtype=torch.int32, device=dev)
sd._adopt_stream(eng, x)
fs = sd.FourStepP2P(eng, L, dev)
def ev(): return torch.cuda.Event(enable_timing=True)
def timeit(fn, reps=5):
    for _ in range(2): fn()
    torch.cuda.synchronize(); dist.barrier()
    a, b = ev(), ev(); a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b) / reps], dtype=torch.float64, device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0])
y = x.clone(); out = torch.empty_like(x)
res = {"world": world}
res["phase0_local_ms"] = timeit(lambda: eng.ntt_dist_phase(0, y.data_ptr(), y.data_ptr(), L, 1, L, w, world, rank, False))
res["all_to_all_plus_transpose_ms"] = timeit(lambda: sd.transpose_exchange(y))
z = sd.transpose_exchange(y)
res["phase1_contiguous_ms"] = timeit(lambda: eng.ntt_dist_phase(1, z.data_ptr(), out.data_ptr(), L, 1, L, w, world, rank, False))
def p0():
    fs.hdl.barrier(channel=0)
    eng.ntt_dist_phase0_p2p(y.data_ptr(), L, w, world, rank, fs.ptrs, False)
    fs.hdl.barrier(channel=1)
res["phase0_p2p_with_barriers_ms"] = timeit(p0)
res["phase1_rotated_ms"] = timeit(lambda: eng.ntt_dist_phase(2, fs.recv.data_ptr(), out.data_ptr(), L, 1, L, w, world, rank, False))
res["barrier_pair_ms"] = timeit(lambda: (fs.hdl.barrier(channel=0), fs.hdl.barrier(channel=1)))
if rank == 0: print(json.dumps(res), flush=True)
dist.barrier(); dist.destroy_process_group()
