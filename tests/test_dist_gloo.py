"""world_size-2 (and 4) gloo tests, on CPU, of the multi-GPU host logic in starks_b200/dist.py:
the leaf-row exchange + subtree-root combination of the sharded Merkle commit, the
block/cyclic redistributions and the four-step transpose.  The device kernels are replaced
by the CPU oracle here (test stand-in only); the data movement and index maps are the code
under test."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P = 2**256 - 351 * 2**32 + 1


def _free_port():
  s = socket.socket()
  s.bind(("127.0.0.1", 0))
  port = s.getsockname()[1]
  s.close()
  return port


def _init(rank, world, port):
  sys.path.insert(0, ROOT)
  sys.path.insert(0, os.path.join(ROOT, "oracle"))
  os.environ["MASTER_ADDR"] = "127.0.0.1"
  os.environ["MASTER_PORT"] = str(port)
  dist.init_process_group("gloo", rank=rank, world_size=world)


def _global_matrix(ncols, n):
  rng = np.random.default_rng(7)
  a = rng.integers(0, 2**32, size=(ncols, n, 8), dtype=np.uint64).astype(np.uint32)
  a[:, :, 7] &= 0x7FFFFFFF
  return a


def _commit_worker(rank, world, port, ncols, n, q):
  try:
    _init(rank, world, port)
    import oracle as orc
    from starks_b200 import dist as sd
    from starks_b200.limbs import limbs_to_be_bytes
    full = _global_matrix(ncols, n)
    cl = ncols // world
    mine = torch.from_numpy(full[rank * cl:(rank + 1) * cl].view(np.int32).copy())
    rows = sd.exchange_leaf_rows(mine, None)
    rows_np = rows.numpy().view(np.uint32)
    # expected rows: all columns, rows {j*q4 + rank*q4/G + i}
    q4 = n // 4
    idx = [j * q4 + rank * (q4 // world) + i for j in range(4) for i in range(q4 // world)]
    assert rows_np.shape == (ncols, n // world, 8)
    assert (rows_np == full[:, idx, :]).all()
    # local subtree (oracle stands in for stk_merkle_commit), roots, top levels
    leaves = np.concatenate([limbs_to_be_bytes(c) for c in rows_np], axis=1)
    _, nodes = orc.merkelize_bytes(leaves, 1)
    roots = sd.allgather_roots(nodes[1].tobytes(), None)
    top = sd.combine_subtree_roots(roots)
    # reference: one tree over everything
    all_leaves = np.concatenate([limbs_to_be_bytes(c) for c in full], axis=1)
    _, want = orc.merkelize_bytes(all_leaves, 1)
    assert top[1] == want[1].tobytes()
    for i in range(1, 2 * world):
      assert top[i] == want[i].tobytes()
    # local node i at depth d, offset o is global node (G + rank) * 2^d + o
    for i in (1, 2, 3, n // world - 1):
      d = i.bit_length() - 1
      gi = (world + rank) * (1 << d) + (i - (1 << d))
      assert nodes[i].tobytes() == want[gi].tobytes()
    q.put((rank, "ok"))
  except Exception as e:  # pragma: no cover
    import traceback
    q.put((rank, "FAIL: " + traceback.format_exc()))
  finally:
    if dist.is_initialized():
      dist.destroy_process_group()


def _ntt_worker(rank, world, port, logn, q):
  try:
    _init(rank, world, port)
    import oracle as orc
    from starks_b200 import dist as sd
    n = 1 << logn
    g = world.bit_length() - 1
    L = n // world
    w = pow(7, (P - 1) // n, P)
    x = [orc.synth(1, i) for i in range(n)]
    # block <-> cyclic
    blk = torch.from_numpy(orc.to_limbs(x[rank * L:(rank + 1) * L]).view(np.int32).copy())
    cyc = sd.block_to_cyclic(blk, None)
    assert orc.from_limbs(cyc.numpy().view(np.uint32)) == [x[rank + world * m] for m in range(L)]
    assert torch.equal(sd.cyclic_to_block(cyc, None), blk)
    # phase 0 on ints: DIF levels with half-size H >= G act inside one residue class
    y = [x[rank + world * m] for m in range(L)]
    H = n // 2
    while H >= world:
      hl = H // world
      for m in range(L):
        if (m // hl) % 2 == 0:
          J = (m * world) | rank
          a, b = y[m], y[m + hl]
          y[m] = (a + b) % P
          y[m + hl] = (a - b) * pow(w, (J % H) * (n // (2 * H)), P) % P
      H //= 2
    z = sd.transpose_exchange(torch.from_numpy(orc.to_limbs(y).view(np.int32).copy()), None)
    zz = orc.from_limbs(z.numpy().view(np.uint32))
    # phase 1 on ints: remaining levels on the contiguous block, then bit-reversed store
    H = world // 2
    while H >= 1:
      for jl in range(L):
        if (jl // H) % 2 == 0:
          J = rank * L + jl
          a, b = zz[jl], zz[jl + H]
          zz[jl] = (a + b) % P
          zz[jl + H] = (a - b) * pow(w, (J % H) * (n // (2 * H)), P) % P
      H //= 2
    out = [None] * L
    for jl in range(L):
      K = int(format(rank * L + jl, "0%db" % logn)[::-1], 2)
      out[K >> g] = (K, zz[jl])
    want = orc.fft_1d(P, x, w, order=n)
    for pos, (K, v) in enumerate(out):
      assert sd.output_owner(K, world) == (rank, pos)
      assert v == want[K]
    q.put((rank, "ok"))
  except Exception as e:  # pragma: no cover
    import traceback
    q.put((rank, "FAIL: " + traceback.format_exc()))
  finally:
    if dist.is_initialized():
      dist.destroy_process_group()


def _comm_worker(rank, world, port, q):
  """The sharded prover's exchanges (dist.NcclComm) over gloo on CPU tensors: subtree roots from
  the ranks' node buffers, the one-owner-per-record byte sum, the column all-gather."""
  try:
    _init(rank, world, port)
    from starks_b200 import dist as sd
    comm = sd.NcclComm(torch.device("cpu"))
    assert (comm.world, comm.rank) == (world, rank)
    nodes = torch.zeros((16, 32), dtype=torch.uint8)
    nodes[1] = torch.arange(32, dtype=torch.uint8) + rank
    roots = sd.allgather_roots_from_nodes(nodes, None)
    assert roots == [bytes((i + r) % 256 for i in range(32)) for r in range(world)]
    top = sd.combine_subtree_roots(roots)
    assert sorted(top) == list(range(1, 2 * world))
    rec, k = 96, 10
    buf = np.zeros(rec * k, dtype=np.uint8)
    want = np.zeros(rec * k, dtype=np.uint8)
    for i in range(k):
      owner = (3 * i + 1) % world
      val = np.arange(rec, dtype=np.uint8) * (i + 1) % 251 + 1
      want[i * rec:(i + 1) * rec] = val
      if owner == rank:
        buf[i * rec:(i + 1) * rec] = val
    got = comm.allreduce_bytes(buf)
    assert (got == want).all()
    lq = 8
    col_local = (torch.arange(lq * 8, dtype=torch.int32).view(lq, 8) + 1000 * rank)
    column = torch.empty((lq * world, 8), dtype=torch.int32)
    comm.allgather_column(None, column, col_local)
    for r in range(world):
      assert torch.equal(column[r * lq:(r + 1) * lq], torch.arange(lq * 8, dtype=torch.int32).view(lq, 8) + 1000 * r)
    q.put((rank, "ok"))
  except Exception as e:  # pragma: no cover
    import traceback
    q.put((rank, "FAIL: " + traceback.format_exc()))
  finally:
    if dist.is_initialized():
      dist.destroy_process_group()


def _run(worker, world, *args):
  ctx = mp.get_context("spawn")
  q = ctx.Queue()
  port = _free_port()
  procs = [ctx.Process(target=worker, args=(r, world, port) + args + (q,)) for r in range(world)]
  for p in procs:
    p.start()
  res = [q.get(timeout=180) for _ in procs]
  for p in procs:
    p.join(timeout=60)
  for r, msg in res:
    assert msg == "ok", "rank %d: %s" % (r, msg)


@pytest.mark.parametrize("world", [2, 4])
def test_sharded_commit_plumbing(world):
  _run(_commit_worker, world, 4, 64)


@pytest.mark.parametrize("world,logn", [(2, 6), (4, 7)])
def test_four_step_layouts(world, logn):
  _run(_ntt_worker, world, logn)


@pytest.mark.parametrize("world", [2, 4])
def test_sharded_prover_exchanges(world):
  _run(_comm_worker, world)


def test_pack_rows_single_process():
  sys.path.insert(0, ROOT)
  from starks_b200.dist import pack_rows_for_leaf_owners, combine_subtree_roots, output_owner
  n, cl, world = 32, 3, 4
  ev = torch.arange(cl * n * 8, dtype=torch.int32).view(cl, n, 8)
  packed = pack_rows_for_leaf_owners(ev, world)
  assert packed.shape == (world, cl, n // world, 8)
  q4 = n // 4
  for d in range(world):
    idx = [j * q4 + d * (q4 // world) + i for j in range(4) for i in range(q4 // world)]
    assert torch.equal(packed[d], ev[:, idx, :])
  roots = [bytes([r]) * 32 for r in range(4)]
  top = combine_subtree_roots(roots)
  import hashlib
  assert top[2] == hashlib.blake2s(roots[0] + roots[1]).digest() and set(top) == set(range(1, 8))
  assert output_owner(5, 4) == (2, 1) and output_owner(6, 4) == (1, 1) and output_owner(7, 1) == (0, 7)


def test_sharded_tree_branches_equal_global_branches(oracle):
  """Index logic of the sharded prover (dist.leaf_owner / extend_branch / split_columns): a branch
  cut from the owner's subtree and extended through the replicated top levels is exactly
  mk_branch of the one big tree (starks/merkle_tree.py:59-68)."""
  from starks_b200.dist import leaf_owner, extend_branch, split_columns, combine_subtree_roots
  n = 64
  leaves = [oracle.blake(bytes([i])) for i in range(n)]
  tree = oracle.merkelize(leaves)
  for world in (2, 4, 8):
    q, lq = n // 4, n // 4 // world
    local_rows = [[leaves[j * q + d * lq + t] for j in range(4) for t in range(lq)] for d in range(world)]
    local_trees = [oracle.merkelize(r) for r in local_rows]
    top = combine_subtree_roots([t[1] for t in local_trees])
    assert top[1] == tree[1]
    for i in range(1, 2 * world):
      assert top[i] == tree[i]
    seen = set()
    for x in range(n):
      d, loc = leaf_owner(x, n, world)
      assert local_rows[d][loc] == leaves[x]
      seen.add((d, loc))
      br = extend_branch(oracle.mk_branch(local_trees[d], loc), top, world, d)
      assert br == oracle.mk_branch(tree, x)
      assert oracle.verify_branch(tree[1], x, br) == leaves[x]
    assert len(seen) == n
  assert split_columns(6, 8) == [(0, 1), (1, 2), (2, 3), (3, 4), (4, 5), (5, 6), (6, 6), (6, 6)]
  assert split_columns(6, 4) == [(0, 2), (2, 4), (4, 5), (5, 6)]
  assert split_columns(48, 8)[3] == (18, 24)
