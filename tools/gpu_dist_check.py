"""Multi-GPU check + timing (run under torchrun on N GPUs of one box):
   torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/gpu_dist_check.py
Parity: dist_ntt vs the single-GPU transform; sharded LDE+commit root vs single-GPU root.
Timing: BASELINE config 3 (64 columns x 2^18 -> 2^21, columns sharded) and config 4
(one 2^26-point NTT, four-step)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from starks_b200 import Engine
from starks_b200 import dist as sd

P = 2**256 - 351 * 2**32 + 1
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
eng = Engine(local)
res = {"world": world}


def rand_limbs(shape, seed):
  g = torch.Generator(device="cpu"); g.manual_seed(seed)
  a = torch.randint(0, 2**31 - 1, shape + (8,), dtype=torch.int32, generator=g)
  return a


def tmax(ms):
  t = torch.tensor([ms], dtype=torch.float64, device=dev)
  dist.all_reduce(t, op=dist.ReduceOp.MAX)
  return float(t[0])

# ---------------- parity: four-step NTT
for logn in (8, 12, 16, 20, 22):
  n = 1 << logn
  L = n // world
  w = pow(7, (P - 1) // n, P)
  x = rand_limbs((n,), 100 + logn).to(dev)            # same on every rank
  ref = torch.empty_like(x)
  sd._adopt_stream(eng, x)
  eng.ntt(x.data_ptr(), n, n, ref.data_ptr(), n, n, 1, w)
  mine = x[rank::world].contiguous()                   # cyclic shard
  out = sd.dist_ntt(eng, mine, w)
  torch.cuda.synchronize()
  g = world.bit_length() - 1
  rho = int(format(rank, "0%db" % g)[::-1], 2) if g else 0
  want = ref[rho::world]
  ok = torch.equal(out, want)
  inv = sd.dist_ntt(eng, mine, w, inverse=True)
  refi = torch.empty_like(x)
  eng.ntt(x.data_ptr(), n, n, refi.data_ptr(), n, n, 1, w, inverse=True)
  torch.cuda.synchronize()
  ok = ok and torch.equal(inv, refi[rho::world])
  res["dist_ntt_2^%d_ok" % logn] = bool(ok)
  assert ok, "dist_ntt mismatch at 2^%d on rank %d" % (logn, rank)

# ---------------- parity + timing: fused exchange (P2P stores from the phase-0 kernel)
try:
  for logn in (12, 20, 26):
    n = 1 << logn
    L = n // world
    w = pow(7, (P - 1) // n, P)
    fs = sd.FourStepP2P(eng, L, dev)
    if logn < 26:
      x = rand_limbs((n,), 300 + logn).to(dev)
      ref = torch.empty_like(x)
      sd._adopt_stream(eng, x)
      eng.ntt(x.data_ptr(), n, n, ref.data_ptr(), n, n, 1, w)
      mine = x[rank::world].contiguous()
      out = fs.ntt(mine, w)
      torch.cuda.synchronize()
      g = world.bit_length() - 1
      rho = int(format(rank, "0%db" % g)[::-1], 2) if g else 0
      ok = torch.equal(out, ref[rho::world])
      res["p2p_ntt_2^%d_ok" % logn] = bool(ok)
      assert ok, "fused-exchange NTT mismatch at 2^%d on rank %d" % (logn, rank)
    else:
      mine = torch.randint(0, 2**31 - 1, (L, 8), dtype=torch.int32, device=dev)
      for _ in range(2):
        fs.ntt(mine, w)
      torch.cuda.synchronize(); dist.barrier()
      e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
      e0.record()
      for _ in range(3):
        fs.ntt(mine, w)
      e1.record()
      torch.cuda.synchronize()
      res["cfg4_p2p_ntt_2^26_ms"] = tmax(e0.elapsed_time(e1) / 3)
    del fs
except Exception as ex:
  import traceback
  res["p2p_error"] = traceback.format_exc()[-600:]

# ---------------- parity: sharded LDE + commit
steps, ext, ncols = 1 << 12, 8, 8
n = steps * ext
g2 = pow(7, (P - 1) // n, P)
trace = rand_limbs((ncols, steps), 7).to(dev)
sc = sd.ShardedCommit(eng)
cl = ncols // world
root, top, evals, rows, nodes = sc.lde_commit(trace[rank * cl:(rank + 1) * cl].contiguous(), ext, g2)
d_ev = torch.empty((ncols, n, 8), dtype=torch.int32, device=dev)
d_nodes = torch.empty((n, 32), dtype=torch.uint8, device=dev)
want_root = eng.lde_commit(trace.data_ptr(), steps, steps, ext, ncols, g2, d_ev.data_ptr(), n, d_nodes.data_ptr())
res["sharded_commit_ok"] = bool(root == want_root)
assert root == want_root, "sharded commit root mismatch"
# subtree nodes are the global nodes (G + rank) * 2^d + o
full_nodes = d_nodes.cpu().numpy()
loc = nodes.cpu().numpy()
for i in (1, 2, 5, n // world - 1):
  d = i.bit_length() - 1
  gi = (world + rank) * (1 << d) + (i - (1 << d))
  assert (loc[i] == full_nodes[gi]).all()

# ---------------- timing: config 3, columns sharded (strong scaling over 64 columns)
steps, ext, ncols = 1 << 18, 8, 64
n = steps * ext
g2 = pow(7, (P - 1) // n, P)
cl = ncols // world
trace = torch.randint(0, 2**31 - 1, (cl, steps, 8), dtype=torch.int32, device=dev)
for _ in range(2):
  sc.lde_commit(trace, ext, g2)
torch.cuda.synchronize(); dist.barrier()
reps = 3
t0 = time.perf_counter()
for _ in range(reps):
  r_ = sc.lde_commit(trace, ext, g2)
torch.cuda.synchronize()
res["cfg3_lde_merkle_commit_ms"] = tmax((time.perf_counter() - t0) / reps * 1e3)
del r_, trace
torch.cuda.empty_cache()

# ---------------- config 3 with the exchange fused into the LDE (P2P row scatter)
try:
  steps, ext, ncols = 1 << 12, 8, 8
  n = steps * ext
  g2 = pow(7, (P - 1) // n, P)
  tr_small = rand_limbs((ncols, steps), 7).to(dev)
  cl = ncols // world
  scp = sd.ShardedCommitP2P(eng, ncols, n, dev)
  root_p, _ = scp.lde_commit(tr_small[rank * cl:(rank + 1) * cl].contiguous(), ext, g2)
  res["sharded_commit_p2p_ok"] = bool(root_p == want_root)
  assert root_p == want_root, "fused sharded commit root mismatch"
  del scp
  steps, ncols = 1 << 18, 64
  n = steps * ext
  g2 = pow(7, (P - 1) // n, P)
  cl = ncols // world
  trace = torch.randint(0, 2**31 - 1, (cl, steps, 8), dtype=torch.int32, device=dev)
  scp = sd.ShardedCommitP2P(eng, ncols, n, dev)
  for _ in range(2):
    scp.lde_commit(trace, ext, g2)
  torch.cuda.synchronize(); dist.barrier()
  t0 = time.perf_counter()
  for _ in range(reps):
    scp.lde_commit(trace, ext, g2)
  torch.cuda.synchronize()
  res["cfg3_p2p_lde_merkle_commit_ms"] = tmax((time.perf_counter() - t0) / reps * 1e3)
  del scp, trace
  torch.cuda.empty_cache()
except Exception as ex:
  import traceback
  res["p2p_commit_error"] = traceback.format_exc()[-600:]

# ---------------- timing: config 4, one 2^26-point NTT
logn = 26
n = 1 << logn
L = n // world
w = pow(7, (P - 1) // n, P)
mine = torch.randint(0, 2**31 - 1, (L, 8), dtype=torch.int32, device=dev)
for _ in range(2):
  out = sd.dist_ntt(eng, mine, w)
torch.cuda.synchronize(); dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
  out = sd.dist_ntt(eng, mine, w)
e1.record()
torch.cuda.synchronize()
res["cfg4_ntt_2^26_ms"] = tmax(e0.elapsed_time(e1) / reps)
res["cfg4_melem_per_s"] = n / (res["cfg4_ntt_2^26_ms"] * 1e-3) / 1e6
# ---------------- timing: config 5 shape, replicas only (one full 2^20-step proof per rank):
# Fiat-Shamir serialises the proof and one GPU finishes it in ~30 ms, so ranks do not split it
from starks_b200.limbs import ints_to_limbs
from starks_b200.modp import IntegersModP
from starks_b200.stark import STARK
psteps = 1 << 20
a, b, c0, c1 = 0, 1 + rank, [], []
for _ in range(psteps):
  c0.append(a); c1.append(b); a, b = b, (a + b) % P
witness = np.stack([ints_to_limbs(c0), ints_to_limbs(c1)])
S = STARK(IntegersModP(P), psteps, 8, 2, [{(0, 1): 1}, {(1, 0): 1, (0, 1): 1}], engine=eng)
bnd = [(0, 0, 0), (0, 1, 1 + rank)]
for _ in range(2):
  proof = S.mk_proof(witness, bnd)
assert S.verify_proof(proof, witness, bnd)
torch.cuda.synchronize(); dist.barrier()
t0 = time.perf_counter()
for _ in range(3):
  S.mk_proof(witness, bnd)
torch.cuda.synchronize()
ms = tmax((time.perf_counter() - t0) / 3 * 1e3)
res["cfg5_proof_ms_per_rank"] = ms
res["cfg5_proofs_per_s_all_ranks"] = world / (ms * 1e-3)
if rank == 0:
  print(json.dumps(res), flush=True)
dist.barrier()
dist.destroy_process_group()
