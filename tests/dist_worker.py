"""Worker of tests/test_gpu_multi.py: run under torchrun, one rank per GPU of one box.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 \
        --master-port P tests/dist_worker.py [--full]

Parity of every multi-GPU path against the ONE-GPU kernels on the same inputs (which the
single-GPU tests pin to the oracle), bit for bit:
  * four-step NTT with the NCCL all-to-all (dist_ntt), forward and inverse, 2^8 .. 2^22;
  * four-step NTT with the exchange fused into the phase-0 kernel (FourStepP2P), 2^12, 2^20
    and -- with --full -- the BASELINE config 4 size 2^26;
  * sharded LDE + Merkle commit (ShardedCommit: NCCL all-to-all; ShardedCommitP2P: P2P stores
    from the transform's final pass): root and subtree nodes against the one-GPU tree, at
    8 x 2^12 and -- with --full -- BASELINE config 3 (64 x 2^18 -> 2^21);
  * one proof sharded over the ranks (ShardedProver) equals the one-GPU proof object.
Rank 0 prints one JSON line of flags; any mismatch raises on the rank that sees it."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist
from starks_b200 import Engine
from starks_b200 import dist as sd

P = 2**256 - 351 * 2**32 + 1
full = "--full" in sys.argv
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
eng = Engine(local)
res = {"world": world, "full": full}
g = world.bit_length() - 1
rho = int(format(rank, "0%db" % g)[::-1], 2) if g else 0


def same(shape, seed):
  """The same pseudo-random canonical elements on every rank (made on rank 0, broadcast)."""
  gen = torch.Generator(device=dev)
  gen.manual_seed(seed)
  t = torch.randint(0, 2**31 - 1, tuple(shape) + (8,), dtype=torch.int32, device=dev, generator=gen)
  dist.broadcast(t, src=0)
  return t


# ---------------- four-step NTT, NCCL exchange
for logn in (8, 12, 16, 20, 22):
  n = 1 << logn
  if n // world < 8:
    continue
  w = pow(7, (P - 1) // n, P)
  x = same((n,), 100 + logn)
  ref, refi = torch.empty_like(x), torch.empty_like(x)
  sd._adopt_stream(eng, x)
  eng.ntt(x.data_ptr(), n, n, ref.data_ptr(), n, n, 1, w)
  eng.ntt(x.data_ptr(), n, n, refi.data_ptr(), n, n, 1, w, inverse=True)
  mine = x[rank::world].contiguous()
  out = sd.dist_ntt(eng, mine, w)
  inv = sd.dist_ntt(eng, mine, w, inverse=True)
  torch.cuda.synchronize()
  assert torch.equal(out, ref[rho::world]), "dist_ntt forward mismatch at 2^%d on rank %d" % (logn, rank)
  assert torch.equal(inv, refi[rho::world]), "dist_ntt inverse mismatch at 2^%d on rank %d" % (logn, rank)
  for k in (0, 1, n // 2 + 3, n - 1):
    assert sd.output_owner(k, world) == (int(format(k % world, "0%db" % g)[::-1], 2) if g else 0, k // world)
res["dist_ntt_nccl_ok"] = True

# ---------------- four-step NTT, exchange fused into the kernel (P2P stores over NVLink)
for logn in (12, 20) + ((26,) if full else ()):
  n = 1 << logn
  L = n // world
  w = pow(7, (P - 1) // n, P)
  x = same((n,), 300 + logn)
  ref = torch.empty_like(x)
  sd._adopt_stream(eng, x)
  eng.ntt(x.data_ptr(), n, n, ref.data_ptr(), n, n, 1, w)
  mine = x[rank::world].contiguous()
  want = ref[rho::world].contiguous()
  del x, ref
  fs = sd.FourStepP2P(eng, L, dev)
  for rep in range(3):          # repeated use of the same exchange buffer
    out = fs.ntt(mine, w)
    torch.cuda.synchronize()
    assert torch.equal(out, want), "fused-exchange NTT mismatch at 2^%d on rank %d (rep %d)" % (logn, rank, rep)
  if logn == 26:
    o2 = sd.dist_ntt(eng, mine, w)
    torch.cuda.synchronize()
    assert torch.equal(o2, want), "NCCL four-step mismatch at 2^26 on rank %d" % rank
    del o2
  del fs, out, want, mine
  torch.cuda.empty_cache()
  res["dist_ntt_p2p_2^%d_ok" % logn] = True

# ---------------- sharded LDE + commit
def subtree_to_global(n_local):
  """Heap index i of rank r's subtree (root at 1) -> index in the global heap: the subtree root
  is global node G + r, so node i at depth d, offset o is (G + r) * 2^d + o."""
  i = np.arange(n_local, dtype=np.int64)
  i[0] = 1
  d = np.floor(np.log2(i)).astype(np.int64)
  return (world + rank) * (1 << d) + (i - (1 << d))


for steps, ncols in ((1 << 12, 8),) + (((1 << 18, 64),) if full else ()):
  if ncols % world:
    continue
  ext = 8
  n = steps * ext
  g2 = pow(7, (P - 1) // n, P)
  trace = same((ncols, steps), 7 + ncols)
  cl = ncols // world
  mine = trace[rank * cl:(rank + 1) * cl].contiguous()
  d_ev = torch.empty((ncols, n, 8), dtype=torch.int32, device=dev)
  d_nodes = torch.empty((n, 32), dtype=torch.uint8, device=dev)
  sd._adopt_stream(eng, trace)
  want_root = eng.lde_commit(trace.data_ptr(), steps, steps, ext, ncols, g2, d_ev.data_ptr(), n, d_nodes.data_ptr())
  full_nodes = d_nodes.cpu().numpy()
  sc = sd.ShardedCommit(eng)
  root, top, evals, rows, nodes = sc.lde_commit(mine, ext, g2)
  assert root == want_root, "sharded commit (NCCL) root mismatch"
  assert torch.equal(evals, d_ev[rank * cl:(rank + 1) * cl]), "column-sharded evaluations differ"
  loc = nodes.cpu().numpy()
  assert (loc[1:] == full_nodes[subtree_to_global(n // world)][1:]).all(), "subtree nodes differ from the one-GPU tree"
  for i, dg in top.items():
    if i >= 1 and i < 2 * world:
      assert dg == full_nodes[i].tobytes(), "replicated top level node %d" % i
  del evals, rows, nodes
  scp = sd.ShardedCommitP2P(eng, ncols, n, dev)
  for rep in range(2):
    root_p, _ = scp.lde_commit(mine, ext, g2)
    assert root_p == want_root, "sharded commit (fused P2P) root mismatch (rep %d)" % rep
  # the same commit from a pinned host trace, upload pipelined with the per-group transforms
  h_mine = mine.cpu().pin_memory()
  stage = torch.empty_like(mine)
  for rep in range(2):
    root_h, _ = scp.lde_commit_host(h_mine, stage, ext, g2)
    assert root_h == want_root, "sharded commit from a host trace: root mismatch (rep %d)" % rep
  torch.cuda.synchronize()
  # the rows a rank received are exactly a local tree's rows in permute4 order
  q = n // 4
  lq = q // world
  rows_want = torch.stack([d_ev[:, j * q + rank * lq:j * q + (rank + 1) * lq] for j in range(4)], dim=1).reshape(ncols, n // world, 8)
  assert torch.equal(scp.rows, rows_want), "P2P row scatter landed rows in the wrong place"
  assert (scp.nodes.cpu().numpy()[1:] == full_nodes[subtree_to_global(n // world)][1:]).all(), "P2P subtree nodes"
  del scp, d_ev, d_nodes, trace
  torch.cuda.empty_cache()
  res["sharded_commit_%dx2^%d_ok" % (ncols, steps.bit_length() - 1)] = True

# ---------------- one proof over all ranks
if hasattr(sd, "ShardedProver"):
  from starks_b200.modp import IntegersModP
  from starks_b200.stark import STARK
  cases = [(1 << 12, 2, [{(0, 1): 1}, {(1, 0): 1, (0, 1): 1}], [0, 1]),
           (1 << 11, 8, None, None)]
  if full:
    cases.append((1 << 20, 2, [{(0, 1): 1}, {(1, 0): 1, (0, 1): 1}], [0, 1]))
    cases.append((1 << 15, 9, None, None))   # 27 columns: uneven split (get_pseudorandom_ks serves widths below 10, stark.py:118-126)
  for steps, width, sp, inp in cases:
    if sp is None:  # a wide affine AIR: x_j' = x_j + x_(j+1 mod w)
      unit = lambda k: tuple(1 if i == k else 0 for i in range(width))
      sp = [{unit(j): 1, unit((j + 1) % width): 1} for j in range(width)]
      inp = list(range(1, width + 1))
    from starks_b200.air import witness_limbs
    wit = witness_limbs(IntegersModP(P), inp, steps, width, sp, engine=eng)
    bnd = [(0, j, inp[j]) for j in range(width)]
    eng.set_stream(0)
    want = STARK(IntegersModP(P), steps, 8, width, sp, engine=eng).mk_proof(wit, bnd)
    sp_ = sd.ShardedProver(eng, IntegersModP(P), steps, 8, width, sp, dev)
    got = sp_.mk_proof(wit, bnd)
    if rank == 0:
      assert got == want, "sharded proof differs from the one-GPU proof (steps %d, width %d)" % (steps, width)
    res["sharded_proof_%d_w%d_ok" % (steps, width)] = True
    del sp_
    torch.cuda.empty_cache()

torch.cuda.synchronize()
dist.barrier()
if rank == 0:
  print("DIST_WORKER_RESULT " + json.dumps(res), flush=True)
dist.destroy_process_group()
