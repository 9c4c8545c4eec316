"""BASELINE configs at full size on one GPU, checked through size-independent properties
(the reference cannot run these sizes: its mk_proof is O(n^2)):
  * config 5 shape: Fibonacci AIR, 2^20 steps, 8x blowup -> proof accepted by verify_proof
    (the mirror of the reference verifier: FRI checks, 80 spot checks of the transition and
    boundary constraints, every Merkle branch re-hashed on the host with hashlib);
  * config 3 shape: 64 columns x 2^18 steps LDE + commit -> ev[i*ext] == trace[i], branches
    verify against the root, the root depends on every column."""
import time

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

P = 2**256 - 351 * 2**32 + 1


def fib_witness(steps):
  from starks_b200.limbs import ints_to_limbs
  a, b = 0, 1
  c0, c1 = [], []
  for _ in range(steps):
    c0.append(a)
    c1.append(b)
    a, b = b, (a + b) % P
  return np.stack([ints_to_limbs(c0), ints_to_limbs(c1)]), (c0[-1], c1[-1])


@pytest.mark.parametrize("logsteps", [14, 20])
def test_fibonacci_proof_verifies(logsteps):
  from starks_b200 import Engine
  from starks_b200.modp import IntegersModP
  from starks_b200.stark import STARK
  steps = 1 << logsteps
  F = IntegersModP(P)
  sp = [{(0, 1): 1}, {(1, 0): 1, (0, 1): 1}]
  witness, _ = fib_witness(steps)
  boundary = [(0, 0, 0), (0, 1, 1)]
  eng = Engine(0)
  S = STARK(F, steps, 8, 2, sp, engine=eng)
  t0 = time.time()
  proof = S.mk_proof(witness, boundary)
  dt = time.time() - t0
  print("mk_proof 2^%d steps: %.3f s" % (logsteps, dt))
  assert len(proof) == 4 and len(proof[2]) == 240
  n_layers = len(proof[3])
  assert n_layers == (logsteps - 4 + 1) // 2 + 1 or n_layers >= 2
  assert S.verify_proof(proof, witness, boundary)
  # tampering with the trace polynomial commitment breaks verification
  bad = [proof[0], proof[1], list(proof[2]), proof[3]]
  leaf = bytearray(bad[2][0][0])
  leaf[5] ^= 1
  bad[2][0] = [bytes(leaf)] + list(bad[2][0][1:])
  with pytest.raises(AssertionError):
    S.verify_proof(bad, witness, boundary)
  eng.close()


def test_config3_lde_commit_64_columns():
  from starks_b200 import Engine
  from starks_b200.limbs import limbs_to_be_bytes
  from starks_b200.merkle_tree import verify_branch
  steps, ext, ncols = 1 << 18, 8, 64
  N = steps * ext
  G2 = pow(7, (P - 1) // N, P)
  rng = np.random.default_rng(3)
  trace = rng.integers(0, 2**32, size=(ncols, steps, 8), dtype=np.uint64).astype(np.uint32)
  trace[:, :, 7] &= 0x7FFFFFFF
  eng = Engine(0)
  d_tr = eng.alloc(trace.nbytes).upload(trace)
  d_ev = eng.alloc(ncols * N * 32)
  nodes = eng.alloc(32 * N)
  root = eng.lde_commit(d_tr.ptr, steps, steps, ext, ncols, G2, d_ev.ptr, N, nodes.ptr)
  # no coset shift: the extension restricted to <G1> is the trace (stark.py:217-224)
  for c in (0, 17, 63):
    col = d_ev.download((N, 8), byte_offset=c * N * 32)
    assert (col[::ext] == trace[c]).all()
  idx = [0, 1, 8, N // 4, N // 2 + 3, N - 1]
  for i, br in zip(idx, eng.merkle_paths(d_ev.ptr, N, ncols, N, nodes.ptr, idx)):
    leaf = verify_branch(root, i, br)
    assert len(leaf) == 32 * ncols and len(br) == 22
    if i % ext == 0:
      assert leaf == b"".join(limbs_to_be_bytes(trace[c, i // ext:i // ext + 1]).tobytes() for c in range(ncols))
  eng.close()
