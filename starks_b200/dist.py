"""Multi-GPU plumbing on one 8 x B200 box: one process per GPU, torch.distributed (NCCL over
NVLink/NVSwitch on GPUs, gloo in the CPU tests).  Only the two steps of the path that need
an exchange use a collective; everything else shards with none.

  * ShardedCommit   LDE of a column shard + Merkle commit over ALL columns' leaves:
                    column-sharded evaluations -> (all-to-all) -> leaf-range-sharded rows,
                    one subtree per rank, all-gather of the 32-byte subtree roots, top
                    log2(G) levels replicated (merkle_tree.py:36-56 heap layout is kept:
                    rank r's subtree root is global node G + r).
  * dist_ntt        one large transform over G ranks, four-step with ONE all-to-all
                    (stk_ntt_dist_phase), cyclic in / cyclic out.

The data-movement helpers take and return torch tensors and never compute field arithmetic,
so they are exercised on CPU tensors with gloo (tests/test_dist_gloo.py); the compute hooks
are the engine's device kernels."""
from hashlib import blake2s

import torch
import torch.distributed as dist


def _world(group=None):
  return dist.get_world_size(group), dist.get_rank(group)


def _adopt_stream(engine, t: torch.Tensor):
  """Run the engine's kernels on torch's current stream so that they are ordered with the
  collectives and tensor ops torch enqueues there (no host synchronisation needed)."""
  if t.is_cuda:
    handle = torch.cuda.current_stream(t.device).cuda_stream
    # torch's default stream has handle 0, which stk_set_stream reads as "the context's own
    # stream"; name the legacy default stream explicitly (cudaStreamLegacy == 1)
    engine.set_stream(handle if handle else 1)


# ---------------------------------------------------------------- leaf exchange (commit)

def pack_rows_for_leaf_owners(evals: torch.Tensor, world: int) -> torch.Tensor:
  """evals: (cols_local, N, 8) column-sharded evaluations.  Returns (world, cols_local, N/world, 8):
  block d holds, for every local column, the rows rank d needs for its N/world permuted
  leaves: rows {j*q + d*q/G + i : j < 4, i < q/G}, q = N/4, ordered (j, i) -- which is
  exactly the row order of a local tree of N/world rows under permute4
  (starks/merkle_tree.py:11-23)."""
  cl, n, limbs = evals.shape
  q = n // 4
  assert n % (4 * world) == 0
  v = evals.view(cl, 4, world, q // world, limbs)      # [c][j][d][i]
  return v.permute(2, 0, 1, 3, 4).contiguous().view(world, cl, n // world, limbs)


def exchange_leaf_rows(evals: torch.Tensor, group=None) -> torch.Tensor:
  """All-to-all of pack_rows_for_leaf_owners: returns (world*cols_local, N/world, 8), i.e. ALL
  columns (global column order: source rank major) for this rank's leaf range."""
  world, _ = _world(group)
  send = pack_rows_for_leaf_owners(evals, world)
  recv = torch.empty_like(send)
  dist.all_to_all_single(recv, send, group=group)
  cl, n_local, limbs = send.shape[1], send.shape[2], send.shape[3]
  return recv.view(world * cl, n_local, limbs)


def combine_subtree_roots(roots):
  """Top log2(G) levels of the heap from the G subtree roots (global nodes G .. 2G-1):
  node i = BLAKE2s(node 2i || node 2i+1) (merkle_tree.py:54-55).  Returns {index: digest}."""
  g = len(roots)
  assert g & (g - 1) == 0
  nodes = {g + r: bytes(roots[r]) for r in range(g)}
  for i in range(g - 1, 0, -1):
    nodes[i] = blake2s(nodes[2 * i] + nodes[2 * i + 1]).digest()
  return nodes


def allgather_roots(root: bytes, group=None, device="cpu"):
  world, _ = _world(group)
  mine = torch.tensor(list(root), dtype=torch.uint8, device=device)
  out = [torch.empty_like(mine) for _ in range(world)]
  dist.all_gather(out, mine, group=group)
  return [bytes(t.cpu().tolist()) for t in out]


class ShardedCommit(object):
  """LDE + Merkle commitment with the trace columns sharded over the ranks of `group`."""

  def __init__(self, engine, group=None):
    self.eng, self.group = engine, group
    self.world, self.rank = _world(group)

  def lde_commit(self, trace: torch.Tensor, ext: int, g2: int):
    """trace: (cols_local, steps, 8) int32 CUDA tensor (this rank's columns).
    Returns (root, top_nodes, evals_local, rows_all_columns, local_nodes)."""
    cl, steps, _ = trace.shape
    n = steps * ext
    dev = trace.device
    _adopt_stream(self.eng, trace)
    evals = torch.empty((cl, n, 8), dtype=torch.int32, device=dev)
    self.eng.lde(trace.data_ptr(), steps, steps, ext, cl, g2, evals.data_ptr(), n)
    if self.world == 1:
      nodes = torch.empty((n, 32), dtype=torch.uint8, device=dev)
      root = self.eng.merkle_commit(evals.data_ptr(), n, cl, n, nodes.data_ptr())
      return root, {1: root}, evals, evals, nodes
    rows = exchange_leaf_rows(evals, self.group)            # (cols_total, n/G, 8)
    n_local = n // self.world
    nodes = torch.empty((n_local, 32), dtype=torch.uint8, device=dev)
    sub_root = self.eng.merkle_commit(rows.data_ptr(), n_local, rows.shape[0], n_local, nodes.data_ptr())
    roots = allgather_roots(sub_root, self.group, device=dev)
    top = combine_subtree_roots(roots)
    return top[1], top, evals, rows, nodes


# ---------------------------------------------------------------- four-step NTT

def block_to_cyclic(x: torch.Tensor, group=None) -> torch.Tensor:
  """x: (L, 8) block-distributed (rank s holds indices [s*L, (s+1)*L)) -> cyclic (rank r holds
  indices r + G*m).  One all-to-all."""
  world, _ = _world(group)
  L, limbs = x.shape
  send = x.view(L // world, world, limbs).permute(1, 0, 2).contiguous()   # [r][t] = x[s*L + r + G*t]
  recv = torch.empty_like(send)
  dist.all_to_all_single(recv, send, group=group)
  return recv.view(L, limbs)                                               # [s][t] -> m = s*L/G + t


def cyclic_to_block(x: torch.Tensor, group=None) -> torch.Tensor:
  """Inverse of block_to_cyclic."""
  world, _ = _world(group)
  L, limbs = x.shape
  recv = torch.empty_like(x)
  dist.all_to_all_single(recv, x.contiguous(), group=group)                # [s][t]: from rank s its chunk for me
  return recv.view(world, L // world, limbs).permute(1, 0, 2).contiguous().view(L, limbs)


def transpose_exchange(y: torch.Tensor, group=None) -> torch.Tensor:
  """The four-step transpose: rank r holds y[(m << g) | r] at position m; afterwards rank r'
  holds the contiguous global block [r'*L, (r'+1)*L), low g bits fastest."""
  world, _ = _world(group)
  L, limbs = y.shape
  recv = torch.empty_like(y)
  dist.all_to_all_single(recv, y.contiguous(), group=group)                # [r][m_local]
  return recv.view(world, L // world, limbs).permute(1, 0, 2).contiguous().view(L, limbs)


def output_owner(k: int, world: int):
  """(rank, local position) holding output index k after dist_ntt."""
  g = world.bit_length() - 1
  rho = k % world
  r = int(format(rho, "0%db" % g)[::-1], 2) if g else 0
  return r, k // world


def dist_ntt(engine, x_cyclic: torch.Tensor, root: int, inverse=False, group=None) -> torch.Tensor:
  """x_cyclic: (L, 8) int32 CUDA tensor, rank r holding x[r + G*m].  Returns (L, 8): rank r'
  holds X[K] for K mod G = bitrev(r') at position K // G (see output_owner)."""
  world, rank = _world(group)
  L = x_cyclic.shape[0]
  if world == 1:
    _adopt_stream(engine, x_cyclic)
    out = torch.empty_like(x_cyclic)
    engine.ntt(x_cyclic.data_ptr(), L, L, out.data_ptr(), L, L, 1, root, inverse=inverse)
    return out
  _adopt_stream(engine, x_cyclic)
  y = x_cyclic.clone()
  engine.ntt_dist_phase(0, y.data_ptr(), y.data_ptr(), L, 1, L, root, world, rank, inverse)
  z = transpose_exchange(y, group)
  out = torch.empty_like(z)
  engine.ntt_dist_phase(1, z.data_ptr(), out.data_ptr(), L, 1, L, root, world, rank, inverse)
  return out


class FourStepP2P(object):
  """dist_ntt with the exchange fused into the compute kernel: the last pass of phase 0 stores
  its results directly into the destination ranks' exchange buffers over NVLink (peer
  pointers from torch symmetric memory), so neither an NCCL all-to-all nor a transpose pass
  runs; two device-side barriers order the exchange.  One column of L = N/G elements per rank."""

  def __init__(self, engine, local_n: int, device, group=None):
    import torch.distributed._symmetric_memory as symm_mem
    self.eng, self.L = engine, local_n
    self.group = group if group is not None else dist.group.WORLD
    self.world, self.rank = _world(group)
    self.recv = symm_mem.empty((local_n, 8), dtype=torch.int32, device=device)
    self.hdl = symm_mem.rendezvous(self.recv, self.group)
    self.ptrs = [int(p) for p in self.hdl.buffer_ptrs]

  def ntt(self, x_cyclic: torch.Tensor, root: int, inverse=False) -> torch.Tensor:
    _adopt_stream(self.eng, x_cyclic)
    y = x_cyclic.clone()
    self.hdl.barrier(channel=0)          # every rank is done reading its exchange buffer
    self.eng.ntt_dist_phase0_p2p(y.data_ptr(), self.L, root, self.world, self.rank, self.ptrs, inverse)
    self.hdl.barrier(channel=1)          # every rank's stores have landed
    out = torch.empty_like(x_cyclic)
    self.eng.ntt_dist_phase(2, self.recv.data_ptr(), out.data_ptr(), self.L, 1, self.L, root, self.world, self.rank,
                            inverse)
    return out


class ShardedCommitP2P(object):
  """ShardedCommit with the exchange fused into the LDE: the final pass of every rank's forward
  transform stores each evaluation row directly into the leaf owner's row buffer over NVLink
  (stk_lde_p2p), so the column-sharded evaluations never exist and no all-to-all runs."""

  def __init__(self, engine, cols_total: int, n: int, device, group=None):
    import torch.distributed._symmetric_memory as symm_mem
    self.eng, self.n, self.cols_total = engine, n, cols_total
    self.group = group if group is not None else dist.group.WORLD
    self.world, self.rank = _world(group)
    self.n_local = n // self.world
    self.rows = symm_mem.empty((cols_total, self.n_local, 8), dtype=torch.int32, device=device)
    self.hdl = symm_mem.rendezvous(self.rows, self.group)
    self.ptrs = [int(p) for p in self.hdl.buffer_ptrs]
    self.nodes = torch.empty((self.n_local, 32), dtype=torch.uint8, device=device)

  def lde_commit(self, trace: torch.Tensor, ext: int, g2: int):
    """trace: (cols_local, steps, 8) int32 CUDA tensor.  Returns (root, top_nodes); the rows of
    this rank's leaf range (all columns) stay in self.rows, its subtree nodes in self.nodes."""
    cl, steps, _ = trace.shape
    assert steps * ext == self.n and cl * self.world == self.cols_total
    _adopt_stream(self.eng, trace)
    self.hdl.barrier(channel=0)          # nobody still reads the previous commit's rows
    self.eng.lde_p2p(trace.data_ptr(), steps, steps, ext, cl, g2, self.world, self.rank * cl, self.ptrs)
    self.hdl.barrier(channel=1)          # all rows have landed
    sub_root = self.eng.merkle_commit(self.rows.data_ptr(), self.n_local, self.cols_total, self.n_local,
                                      self.nodes.data_ptr())
    roots = allgather_roots(sub_root, self.group, device=trace.device)
    top = combine_subtree_roots(roots)
    return top[1], top
