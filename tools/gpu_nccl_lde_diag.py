"""Why is the NCCL sharded commit slower when the LDE takes the coset route?  Per-step CUDA-event
times of ShardedCommit.lde_commit's pieces with STK_LDE_R0 on / off (torchrun, 2+ ranks)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from starks_b200 import Engine
from starks_b200 import dist as sd
P = 2**256 - 351 * 2**32 + 1
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
eng = Engine(local)
steps, ext, ncols = 1 << 18, 8, 64
n = steps * ext
g2 = pow(7, (P - 1) // n, P)
cl = ncols // world
mine = torch.randint(0, 2**31 - 1, (cl, steps, 8), dtype=torch.int32, device=dev)
sd._adopt_stream(eng, mine)
res = {"world": world}
def ev():
  e = torch.cuda.Event(enable_timing=True); e.record(); return e
for tag, env in (("r0", None), ("full", "0"), ("r0_again", None)):
  if env is not None: os.environ["STK_LDE_R0"] = env
  acc = {}
  for it in range(5):
    torch.cuda.synchronize(); dist.barrier()
    t0 = time.perf_counter()
    e0 = ev()
    evals = torch.empty((cl, n, 8), dtype=torch.int32, device=dev)
    eng.lde(mine.data_ptr(), steps, steps, ext, cl, g2, evals.data_ptr(), n)
    e1 = ev()
    send = sd.pack_rows_for_leaf_owners(evals, world)
    e2 = ev()
    recv = torch.empty_like(send)
    dist.all_to_all_single(recv, send)
    e3 = ev()
    rows = recv.view(world * cl, n // world, 8)
    nodes = torch.empty((n // world, 32), dtype=torch.uint8, device=dev)
    eng.merkle_commit(rows.data_ptr(), n // world, ncols, n // world, nodes.data_ptr(), want_root=False)
    e4 = ev()
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) * 1e3
    if it >= 2:
      for k, v in (("lde", e0.elapsed_time(e1)), ("pack", e1.elapsed_time(e2)), ("a2a", e2.elapsed_time(e3)),
                   ("commit", e3.elapsed_time(e4)), ("wall", wall)):
        acc.setdefault(k, []).append(v)
    del evals, send, recv, rows, nodes
  res[tag] = {k: round(sum(v) / len(v), 3) for k, v in acc.items()}
  if env is not None: del os.environ["STK_LDE_R0"]
if rank == 0:
  print("DIAG " + json.dumps(res), flush=True)
dist.barrier(); dist.destroy_process_group()
