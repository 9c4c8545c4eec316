"""Host-side check of what a STARK proof OPENS, independent of the prover under test (test
infrastructure; plain Python ints + the CPU oracle).  Used on the CPU against an oracle-made
proof (which pins the checker) and on the GPU against the 2^20-step proof (which pins
stk_lincomb / stk_quotient_eval / stk_boundary_eval at a size the reference cannot reach)."""
from hashlib import blake2s

P = 2**256 - 351 * 2**32 + 1


def _blake(x):
  return blake2s(x).digest()


def _verify_branch(root, index, proof, n):
  """starks/merkle_tree.py:71-86 with hashlib."""
  assert len(proof) == n.bit_length()           # log2(n) + 1 entries
  idx = index // (n // 4) + 4 * (index % (n // 4)) + n
  v = proof[0]
  for sib in proof[1:]:
    v = _blake(sib + v) if idx % 2 else _blake(v + sib)
    idx //= 2
  assert v == root
  return proof[0]


def _indices(entropy, modulus, count, exclude):
  """starks/utils.py:60-90."""
  data = entropy
  while len(data) < 4 * count:
    data += _blake(data[-32:])
  real = modulus * (exclude - 1) // exclude
  o = [int.from_bytes(data[i:i + 4], "big") % real for i in range(0, count * 4, 4)]
  return [x + 1 + x // (exclude - 1) for x in o]


def _ks(m_root, num):
  """starks/stark.py:106-126 (num <= 4: ASCII salts b'0x01'..)."""
  assert num <= 4
  return [int.from_bytes(_blake(m_root + salt), "big") for salt in (b"0x01", b"0x02", b"0x03", b"0x04")[:num]]


def _step(sp, state):
  acc = 0
  for exps, c in sp.items():
    t = c
    for k, e in enumerate(exps):
      t = t * pow(state[k], e, P) % P
    acc = (acc + t) % P
  return acc


def check_opened_values(oracle, proof, witness_limbs, inputs, steps, ext, step_polys, samples=80):
  """witness_limbs: (width, steps, 8) uint32.  Returns the number of positions checked.
    (a) opened P_j(x), P_j(G1 x) == the trace polynomial evaluated independently: coefficients by
        the oracle's inverse transform over <G1> (stark.py:27-36), values by Horner in C;
    (b) D_j(x) Z(x) == P_j(G1 x) - step_j(P(x)),  Z = (x^steps - 1)/(x - last)   (stark.py:57-78)
        B_j(x) (x - 1)(x - last) == P_j(x) - I_j(x)                             (stark.py:80-104)
    (c) l(x) == sum_j (1 + k_j c)(D_j + (k1 + k2 c) P_j + (k3 + k4 c) B_j) with the scalar
        c = (G2^steps)^(N-1) of the leaked loop index (stark.py:130-177)."""
  m_root, l_root, branches, _ = proof
  w = len(step_polys)
  N = steps * ext
  G2 = pow(7, (P - 1) // N, P)
  G1 = pow(G2, ext, P)
  last = pow(G2, (steps - 1) * ext, P)
  positions = _indices(l_root, N, samples, ext)
  assert len(branches) == 3 * samples
  coeffs = oracle.fft_limbs(P, G1, witness_limbs, steps, inv=True, nthreads=min(w, 8))
  xs = [pow(G2, pos, P) for pos in positions]
  pts = xs + [x * G1 % P for x in xs]
  pvals = [oracle.poly_eval(P, oracle.from_limbs(coeffs[j]), pts) for j in range(w)]
  outputs = oracle.from_limbs(witness_limbs[:, -1, :])
  k1, k2, k3, k4 = _ks(m_root, 4)
  l_ks = _ks(m_root, w)
  c = pow(pow(G2, steps, P), N - 1, P)
  for i, pos in enumerate(positions):
    x = xs[i]
    leaf1 = _verify_branch(m_root, pos, branches[3 * i], N)
    leaf2 = _verify_branch(m_root, (pos + ext) % N, branches[3 * i + 1], N)
    l_of_x = int.from_bytes(_verify_branch(l_root, pos, branches[3 * i + 2], N), "big")
    assert len(leaf1) == 96 * w and len(leaf2) == 96 * w
    val = lambda leaf, k: int.from_bytes(leaf[32 * k:32 * k + 32], "big")
    p_x = [val(leaf1, j) for j in range(w)]
    d_x = [val(leaf1, w + j) for j in range(w)]
    b_x = [val(leaf1, 2 * w + j) for j in range(w)]
    p_gx = [val(leaf2, j) for j in range(w)]
    for j in range(w):
      assert p_x[j] == pvals[j][i], "opened P(x) differs from the trace polynomial"
      assert p_gx[j] == pvals[j][samples + i], "opened P(G1 x) differs from the trace polynomial"
    z = (pow(x, steps, P) - 1) * pow((x - last) % P, -1, P) % P
    for j in range(w):
      assert (p_gx[j] - _step(step_polys[j], p_x) - z * d_x[j]) % P == 0, "D is not C / Z"
      slope = (outputs[j] - inputs[j]) * pow((last - 1) % P, -1, P) % P
      interp = (inputs[j] + slope * (x - 1)) % P
      assert (p_x[j] - interp - b_x[j] * ((x - 1) % P) % P * ((x - last) % P)) % P == 0, "B is not (P - I) / Z2"
    want = 0
    for j in range(w):
      want += (1 + l_ks[j] * c) * (d_x[j] + (k1 + k2 * c) * p_x[j] + (k3 + k4 * c) * b_x[j])
    assert want % P == l_of_x, "l(x) is not the pseudorandom linear combination of P, D, B"
  return len(positions)
