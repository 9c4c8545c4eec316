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
  """starks/merkle_tree.py:11-23."""
  o = []
  ld4 = len(values) // 4
  for i in range(ld4):
    o.extend([values[i], values[i + ld4], values[i + ld4 * 2], values[i + ld4 * 3]])
  return o


def get_index_in_permuted(x, L):
  """starks/merkle_tree.py:26-33."""
  ld4 = L // 4
  return x // ld4 + 4 * (x % ld4)


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
  """A branch of the merkle tree is a list (starks/merkle_tree.py:59-68)."""
  index = get_index_in_permuted(index, len(tree) // 2)
  index += len(tree) // 2
  o = [tree[index]]
  while index > 1:
    o.append(tree[index ^ 1])
    index //= 2
  return o


def verify_branch(root, index, proof, output_as_int=False):
  """Verifies the proof and returns the leaf on the branch (starks/merkle_tree.py:71-86)."""
  index = get_index_in_permuted(index, 2**len(proof) // 2)
  index += 2**len(proof) // 2
  v = proof[0]
  for p in proof[1:]:
    if index % 2:
      v = blake(p + v)
    else:
      v = blake(v + p)
    index //= 2
  assert v == root
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
  """starks/merkle_tree.py:121-147."""
  vals = []
  for poly_ind in range(num_polys):
    for dim in range(dims):
      start_index = 32 * (poly_ind * dims + dim)
      end_index = 32 * (poly_ind * dims + dim + 1)
      vals.append(leaf[start_index:end_index])
  return vals
