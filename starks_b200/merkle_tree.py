"""Drop-in for starks/merkle_tree.py: blake, permute4, get_index_in_permuted, merkelize,
mk_branch, verify_branch, merkelize_polynomial_evaluations, unpack_merkle_leaf -- the trees
are built by libstarks_b200 (stk_merkle_commit / stk_merkle_commit_raw)."""
from hashlib import blake2s
from typing import List

import numpy as np

from .engine import default_engine
from .limbs import ints_to_limbs, limbs_to_be_bytes

blake = lambda x: blake2s(x).digest()  # starks/merkle_tree.py:1-5


def permute4(values: List) -> List:
  """starks/merkle_tree.py:11-23: out[4i + j] = values[i + j*(n//4)] -- the four fold-mates of
  FRI position i become adjacent leaves.  Elements beyond 4*(n//4) are dropped, as upstream."""
  q = len(values) // 4
  out = [None] * (4 * q)
  for j in range(4):
    out[j::4] = values[j * q:(j + 1) * q]
  return out


def get_index_in_permuted(x, L):
  """starks/merkle_tree.py:26-33: where element x of an L-element list sits after permute4."""
  j, i = divmod(x, L // 4)
  return 4 * i + j


def _serialise(L):
  """merkle_tree.py:47-53: int -> 32-byte big-endian, bytes as is, else x.to_bytes()."""
  out = []
  for x in L:
    if isinstance(x, int):
      out.append(x.to_bytes(32, "big"))
    elif isinstance(x, (bytes, bytearray)):
      out.append(bytes(x))
    else:
      out.append(x.to_bytes())
  return out


def merkelize(L, engine=None) -> List[bytes]:
  """Creates a merkle-tree representation of the given list (starks/merkle_tree.py:36-56):
  a list of 2n entries, [0] = b'', [1] = root, [n:] = the permuted (unhashed) leaves."""
  ser = _serialise(L)
  n = len(ser)
  npm = 4 * (n // 4)
  if npm == 0:
    return []
  eng = engine or default_engine()
  width = len(ser[0])
  if any(len(s) != width for s in ser):
    raise ValueError("merkelize: leaves of different widths are not supported on the device path")
  leaves = np.frombuffer(b"".join(ser), dtype=np.uint8)
  d_leaves = eng.alloc(max(leaves.nbytes, 1)).upload(leaves)
  d_nodes = eng.alloc(32 * npm)
  eng.merkle_commit_raw(d_leaves.ptr, n, width, d_nodes.ptr)
  nodes = d_nodes.download((npm, 32), np.uint8)
  d_leaves.free()
  d_nodes.free()
  return [b""] + [nodes[i].tobytes() for i in range(1, npm)] + permute4(ser)


def mk_branch(tree, index: int):
  """starks/merkle_tree.py:59-68: [leaf, sibling, sibling of the parent, ...] up to (excluding)
  the root, for the leaf that held element `index` before permute4."""
  n = len(tree) // 2
  node = n + get_index_in_permuted(index, n)
  path = [tree[node]]
  while node > 1:
    path.append(tree[node ^ 1])
    node >>= 1
  return path


def verify_branch(root, index, proof, output_as_int=False):
  """starks/merkle_tree.py:71-86: re-hashes the branch (hashlib) and returns its leaf; raises
  AssertionError when it does not lead to `root`.  A branch of k entries belongs to a tree of
  2^(k-1) leaves."""
  n = (1 << len(proof)) >> 1
  node = n + get_index_in_permuted(index, n)
  acc = proof[0]
  for sibling in proof[1:]:
    acc = blake(sibling + acc) if node & 1 else blake(acc + sibling)
    node >>= 1
  assert acc == root
  return int.from_bytes(proof[0], "big") if output_as_int else proof[0]


def merkelize_polynomial_evaluations(dims, polynomial_evals, engine=None):
  """Given a list of polynomial evaluations, merkelizes them together
  (starks/merkle_tree.py:94-119): leaf i = concatenation of every column's value i."""
  eng = engine or default_engine()
  cols = []
  for col in polynomial_evals:
    if isinstance(col, np.ndarray):
      cols.append(col)
    else:
      cols.append(ints_to_limbs([v.n if hasattr(v, "n") else int(v) for v in col]))
  n = min(len(c) for c in cols)
  ncols = len(cols)
  npm = 4 * (n // 4)
  if npm == 0:
    return []
  arr = np.stack([c[:n] for c in cols])
  if npm & (npm - 1):
    # non power-of-two row counts go through the raw-leaf kernel
    leaves = np.concatenate([limbs_to_be_bytes(c) for c in arr], axis=1)
    return merkelize([leaves[i].tobytes() for i in range(n)], engine=eng)
  d_cols = eng.alloc(arr.nbytes).upload(arr)
  d_nodes = eng.alloc(32 * npm)
  eng.merkle_commit(d_cols.ptr, n, ncols, n, d_nodes.ptr)
  nodes = d_nodes.download((npm, 32), np.uint8)
  d_cols.free()
  d_nodes.free()
  leaves = np.concatenate([limbs_to_be_bytes(c) for c in arr], axis=1)
  leaf_bytes = [leaves[i].tobytes() for i in range(npm)]
  return [b""] + [nodes[i].tobytes() for i in range(1, npm)] + permute4(leaf_bytes)


def unpack_merkle_leaf(leaf: bytes, dims: int, num_polys: int) -> List[bytes]:
  """starks/merkle_tree.py:121-147: the 32-byte values of a leaf made by
  merkelize_polynomial_evaluations, polynomial-major, dimension-minor."""
  return [leaf[32 * k:32 * (k + 1)] for k in range(num_polys * dims)]
