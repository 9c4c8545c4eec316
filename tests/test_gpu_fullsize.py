"""BASELINE configs at full size on one GPU, pinned to the CPU oracle (the reference itself
cannot run these sizes: fft_1d takes 74 s per 2^20 column and mk_proof is O(n^2)):
  * config 2: forward and inverse NTT of 2 x 2^20 and 1 x 2^24 equal the oracle's fft_1d
    restatement element for element;
  * config 3: 64 columns x 2^18 steps LDE + commit: every evaluation column equals the oracle's
    LDE and the root (and every node) equals the oracle's merkelize over those leaves; plus
    ev[i*ext] == trace[i] and branch checks;
  * config 5: Fibonacci AIR, 2^20 steps, 8x blowup: the proof is accepted by verify_proof WITHOUT
    an engine (every Merkle branch re-hashed on the host with hashlib), the opened values are
    checked against an independent evaluation of the trace polynomials (oracle inverse transform
    + Horner) and the linear combination l(x) = sum_j (1 + k_j c)(D_j + (k1 + k2 c) P_j +
    (k3 + k4 c) B_j) is recomputed from the opened leaves at the 80 positions."""
import hashlib
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
  assert S.verify_proof(proof, witness, boundary)          # branches re-hashed by stk_verify_branches
  # the same proof through the verifier WITHOUT an engine: every branch re-hashed with hashlib
  S_host = STARK(F, steps, 8, 2, sp, engine=None)
  assert S_host._engine is None
  assert S_host.verify_proof(proof, witness, boundary)
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


def _threads(oracle):
  import os
  try:
    n = len(os.sched_getaffinity(0))
  except AttributeError:
    n = os.cpu_count() or 1
  return max(1, min(n, 64))


@pytest.mark.slow
@pytest.mark.parametrize("logn,batch", [(20, 2), (24, 1)])
def test_config2_ntt_equals_oracle_at_full_size(oracle, logn, batch):
  """starks/fft.py:316-331 at BASELINE config 2's sizes, element for element."""
  from starks_b200 import Engine
  n = 1 << logn
  w = pow(7, (P - 1) // n, P)
  rng = np.random.default_rng(100 + logn)
  cols = rng.integers(0, 2**32, size=(batch, n, 8), dtype=np.uint64).astype(np.uint32)
  cols[:, :, 7] &= 0x7FFFFFFF
  eng = Engine(0)
  got = eng.ntt_host(cols, n, w)
  want = oracle.fft_limbs(P, w, cols, n, nthreads=_threads(oracle))
  assert hashlib.blake2s(got.tobytes()).digest() == hashlib.blake2s(want.tobytes()).digest()
  assert (got == want).all()
  back = eng.ntt_host(got, n, w, inverse=True)
  assert (back == cols).all()
  if logn <= 20:
    assert (back == oracle.fft_limbs(P, w, want, n, inv=True, nthreads=_threads(oracle))).all()
  eng.close()


@pytest.mark.slow
def test_config3_root_equals_oracle(oracle):
  """64 x 2^18 -> 2^21: evaluations equal the oracle's LDE (stark.py:27-36, 254-256) column by
  column and the tree equals oracle.merkelize over the packed leaves (merkle_tree.py:94-119)."""
  from starks_b200 import Engine
  steps, ext, ncols = 1 << 18, 8, 64
  N = steps * ext
  G2 = pow(7, (P - 1) // N, P)
  G1 = pow(G2, ext, P)
  rng = np.random.default_rng(33)
  trace = rng.integers(0, 2**32, size=(ncols, steps, 8), dtype=np.uint64).astype(np.uint32)
  trace[:, :, 7] &= 0x7FFFFFFF
  eng = Engine(0)
  d_tr = eng.alloc(trace.nbytes).upload(trace)
  d_ev = eng.alloc(ncols * N * 32)
  nodes = eng.alloc(32 * N)
  root = eng.lde_commit(d_tr.ptr, steps, steps, ext, ncols, G2, d_ev.ptr, N, nodes.ptr)
  thr = _threads(oracle)
  leaves = np.empty((N, 32 * ncols), dtype=np.uint8)
  group = 16
  for c0 in range(0, ncols, group):
    coeffs = oracle.fft_limbs(P, G1, trace[c0:c0 + group], steps, inv=True, nthreads=thr)
    ev = oracle.fft_limbs(P, G2, coeffs, N, nthreads=thr)          # zero-padded (fft.py:323-324)
    for k in range(group):
      got = d_ev.download((N, 8), byte_offset=(c0 + k) * N * 32)
      assert (got == ev[k]).all(), "LDE column %d differs from the oracle" % (c0 + k)
      # 32-byte big-endian serialisation (modp.py:94-95) into the leaf matrix
      leaves[:, 32 * (c0 + k):32 * (c0 + k + 1)] = np.ascontiguousarray(ev[k][:, ::-1]).astype(">u4").view(np.uint8).reshape(N, 32)
  _, onodes = oracle.merkelize_bytes(leaves, nthreads=thr)
  assert onodes[1].tobytes() == root, "Merkle root differs from the oracle"
  gnodes = nodes.download((N, 32), np.uint8)
  assert (gnodes[1:] == onodes[1:]).all()
  eng.close()


def test_config5_opened_values_match_independent_evaluation(oracle):
  """2^20-step proof: at the 80 spot-check positions the opened P, D, B and l values are checked
  on the host against (a) the trace polynomials evaluated independently (oracle inverse
  transform + Horner in C, starks/stark.py:27-36), (b) the constraint and boundary identities
  (stark.py:57-104) and (c) the pseudorandom linear combination with the leaked-index scalar c
  (stark.py:130-177) -- which upstream's verifier does NOT check (:374-381).
  The checker itself is pinned on the CPU against an oracle-made proof (tests/test_host_logic.py)."""
  from proofcheck import check_opened_values
  from starks_b200 import Engine
  from starks_b200.modp import IntegersModP
  from starks_b200.stark import STARK
  steps, ext, w = 1 << 20, 8, 2
  sp = [{(0, 1): 1}, {(1, 0): 1, (0, 1): 1}]
  witness, _ = fib_witness(steps)
  eng = Engine(0)
  proof = STARK(IntegersModP(P), steps, ext, w, sp, engine=eng).mk_proof(witness, [(0, 0, 0), (0, 1, 1)])
  eng.close()
  assert check_opened_values(oracle, proof, witness, [0, 1], steps, ext, sp) == 80
