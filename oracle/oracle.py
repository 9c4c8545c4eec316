"""Python face of the CPU oracle (TEST INFRASTRUCTURE ONLY -- see starks_oracle.c).

Thin ctypes wrappers over oracle/liboracle.so plus line-by-line restatements of the
reference's protocol glue (Fiat-Shamir helpers, FRI prover, STARK.mk_proof) on plain
Python ints.  Heavy loops (NTT, BLAKE2s tree, the O(n^2) polynomial division the
reference performs) run in the C library; control flow follows the reference and
cites it.  The product package (starks_b200/) never imports this module.

Parity status: PINNED against golden vectors generated from the unmodified Python
reference (oracle/gen_golden.py -> tests/golden/*.json) and the reference tests' own
known answers; see tests/test_oracle_golden.py.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")

P_STARK = 2**256 - 351 * 2**32 + 1

_u32p = ctypes.POINTER(ctypes.c_uint32)
_u8p = ctypes.POINTER(ctypes.c_uint8)
_lib = None


def build(force=False):
  """Compiles liboracle.so with the committed Makefile (gcc only, seconds)."""
  src = os.path.join(_HERE, "starks_oracle.c")
  if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
    subprocess.check_call(["make", "-C", _HERE, "-s", "liboracle.so"])
  return _LIB_PATH


def lib():
  global _lib
  if _lib is None:
    build()
    L = ctypes.CDLL(_LIB_PATH)
    L.orc_fft.restype = ctypes.c_int64
    L.orc_mul_polys.restype = ctypes.c_int64
    L.orc_merkelize.restype = ctypes.c_int64
    L.orc_power_cycle.restype = ctypes.c_int64
    _lib = L
  return _lib


def threads():
  return int(lib().orc_threads())


# ----------------------------------------------------------------- limb helpers

def to_limbs(ints):
  """list of non-negative ints < 2^256 -> (n, 8) uint32, little-endian limbs."""
  buf = b"".join(int(x).to_bytes(32, "little") for x in ints)
  return np.frombuffer(buf, dtype="<u4").reshape(-1, 8).copy()


def from_limbs(arr):
  b = np.ascontiguousarray(arr, dtype="<u4").tobytes()
  return [int.from_bytes(b[i:i + 32], "little") for i in range(0, len(b), 32)]


def _p(arr, typ=_u32p):
  return arr.ctypes.data_as(typ)


def _one(x):
  return to_limbs([x])


# ------------------------------------------------------------------ field / NTT

def field_op(p, op, a, b=None):
  A = to_limbs(a)
  B = to_limbs(b) if b is not None else None
  out = np.empty_like(A)
  code = {"add": 0, "sub": 1, "mul": 2, "inv": 3, "reduce": 4}[op]
  rc = lib().orc_field_op(_p(_one(p)), code, _p(A), _p(B) if B is not None else None, _p(out),
                          ctypes.c_uint64(len(A)))
  assert rc == 0, rc
  return from_limbs(out)


def fpow(p, a, e):
  out = np.empty((1, 8), dtype=np.uint32)
  rc = lib().orc_pow(_p(_one(p)), _p(_one(a % p)), _p(_one(e)), _p(out))
  assert rc == 0
  return from_limbs(out)[0]


def get_power_cycle(p, r, cap=1 << 24):
  """starks/utils.py:30-38."""
  out = np.empty((cap, 8), dtype=np.uint32)
  n = lib().orc_power_cycle(_p(_one(p)), _p(_one(r)), _p(out), ctypes.c_uint64(cap))
  assert n > 0, n
  return from_limbs(out[:n])


def fft_limbs(p, root, cols, order, inv=False, nthreads=1):
  """cols: (batch, n_in, 8) uint32 -> (batch, order, 8).  fft_1d per column
  (starks/fft.py:316-331); `order` must equal the multiplicative order of root."""
  cols = np.ascontiguousarray(cols, dtype=np.uint32)
  batch, n_in, _ = cols.shape
  out = np.empty((batch, order, 8), dtype=np.uint32)
  rc = lib().orc_fft(_p(_one(p)), _p(_one(root)), _p(cols), ctypes.c_uint64(n_in),
                     ctypes.c_uint64(n_in * 8), _p(out), ctypes.c_uint64(order * 8),
                     ctypes.c_uint64(batch), ctypes.c_uint64(order), int(bool(inv)), int(nthreads))
  if rc == -3:
    raise IndexError("input longer than the order of the root / unsupported length")
  assert rc == order, (rc, order)
  return out


def _order(p, root, cap=1 << 26):
  n, x = 1, root % p
  while x != 1:
    x = x * root % p
    n += 1
    assert n <= cap
  return n


def fft_1d(p, vals, root, inv=False, order=None):
  """starks/fft.py:316-331 on Python ints; returns list of ints of length ord(root)."""
  if order is None:
    order = _order(p, root)
  if len(vals) > order:
    raise IndexError("input longer than the order of the root")
  cols = to_limbs(vals).reshape(1, -1, 8) if len(vals) else np.zeros((1, 0, 8), np.uint32)
  return from_limbs(fft_limbs(p, root, cols, order, inv)[0])


def mul_polys(p, a, b, root):
  """starks/fft.py:334-345 (no 1/N scaling)."""
  order = _order(p, root)
  A, B = to_limbs(a), to_limbs(b)
  out = np.empty((order, 8), dtype=np.uint32)
  rc = lib().orc_mul_polys(_p(_one(p)), _p(_one(root)), _p(A), ctypes.c_uint64(len(A)), _p(B),
                           ctypes.c_uint64(len(B)), _p(out), ctypes.c_uint64(order))
  assert rc == order, rc
  return from_limbs(out)


# ----------------------------------------------------------------------- hashing

def blake(data: bytes) -> bytes:
  """starks/merkle_tree.py:5 -- BLAKE2s-256 (C restatement of RFC 7693)."""
  buf = (ctypes.c_uint8 * max(1, len(data))).from_buffer_copy(data or b"\0")
  out = (ctypes.c_uint8 * 32)()
  lib().orc_blake2s(buf, ctypes.c_uint64(len(data)), out)
  return bytes(out)


def merkelize_bytes(leaves: np.ndarray, nthreads=1):
  """leaves: (n, leaf_len) uint8 -> (leaves_perm (n', leaf_len), nodes (n', 32)).
  starks/merkle_tree.py:36-56."""
  leaves = np.ascontiguousarray(leaves, dtype=np.uint8)
  n, ll = leaves.shape
  npm = 4 * (n // 4)
  perm = np.empty((npm, ll), dtype=np.uint8)
  nodes = np.zeros((max(npm, 1), 32), dtype=np.uint8)
  rc = lib().orc_merkelize(_p(leaves, _u8p), ctypes.c_uint64(n), ctypes.c_uint64(ll), _p(perm, _u8p),
                           _p(nodes, _u8p), int(nthreads))
  assert rc == npm
  return perm, nodes[:npm]


def merkelize(L):
  """starks/merkle_tree.py:36-56 on a list of ints / bytes (all the same width)."""
  ser = [x.to_bytes(32, "big") if isinstance(x, int) else bytes(x) for x in L]
  if not ser:
    return []
  ll = len(ser[0])
  assert all(len(s) == ll for s in ser)
  leaves = np.frombuffer(b"".join(ser), dtype=np.uint8).reshape(len(ser), ll)
  perm, nodes = merkelize_bytes(leaves)
  npm = len(perm)
  if npm == 0:  # permute4 drops everything when len(L) < 4
    return []
  tree = [b""] + [nodes[i].tobytes() for i in range(1, npm)] + [perm[i].tobytes() for i in range(npm)]
  return tree


def get_index_in_permuted(x, L):
  """starks/merkle_tree.py:26-33."""
  ld4 = L // 4
  return x // ld4 + 4 * (x % ld4)


def mk_branch(tree, index):
  """starks/merkle_tree.py:59-68."""
  index = get_index_in_permuted(index, len(tree) // 2)
  index += len(tree) // 2
  o = [tree[index]]
  while index > 1:
    o.append(tree[index ^ 1])
    index //= 2
  return o


def verify_branch(root, index, proof, output_as_int=False):
  """starks/merkle_tree.py:71-86."""
  index = get_index_in_permuted(index, 2**len(proof) // 2)
  index += 2**len(proof) // 2
  v = proof[0]
  for pp in proof[1:]:
    v = blake(pp + v) if index % 2 else blake(v + pp)
    index //= 2
  assert v == root
  return int.from_bytes(proof[0], "big") if output_as_int else proof[0]


def pack_leaves(cols):
  """merkelize_polynomial_evaluations' leaf packing (starks/merkle_tree.py:116-118).
  cols: list of equal-length int lists -> (n, 32*ncols) uint8."""
  ncols, n = len(cols), len(cols[0])
  arr = np.stack([to_limbs(c) for c in cols])  # (ncols, n, 8)
  out = np.empty((n, 32 * ncols), dtype=np.uint8)
  lib().orc_pack_leaves(_p(arr), ctypes.c_uint64(n), ctypes.c_uint64(ncols), ctypes.c_uint64(n),
                        _p(out, _u8p))
  return out


def merkelize_polynomial_evaluations(cols):
  """starks/merkle_tree.py:94-119."""
  leaves = pack_leaves(cols)
  perm, nodes = merkelize_bytes(leaves)
  npm = len(perm)
  return [b""] + [nodes[i].tobytes() for i in range(1, npm)] + [perm[i].tobytes() for i in range(npm)]


# ------------------------------------------------------------------ Fiat-Shamir

def get_pseudorandom_indices(entropy, modulus, count, exclude_multiples_of=0):
  """starks/utils.py:60-90."""
  assert modulus < 2**24
  data = entropy
  while len(data) < 4 * count:
    data += blake(data[-32:])
  if exclude_multiples_of == 0:
    return [int.from_bytes(data[i:i + 4], "big") % modulus for i in range(0, count * 4, 4)]
  real_modulus = modulus * (exclude_multiples_of - 1) // exclude_multiples_of
  o = [int.from_bytes(data[i:i + 4], "big") % real_modulus for i in range(0, count * 4, 4)]
  return [x + 1 + x // (exclude_multiples_of - 1) for x in o]


def get_pseudorandom_ks(m_root, num):
  """starks/stark.py:106-126 (salts are the ASCII strings b'0x01'...)."""
  if 0 <= num <= 4:
    byte_list = [b"0x01", b"0x02", b"0x03", b"0x04"]
    return [int.from_bytes(blake(m_root + byte_list[i]), "big") for i in range(num)]
  elif num < 10:
    byte_list = [("0x0%s" % str(i)).encode("UTF-8") for i in range(num)]
    return [int.from_bytes(blake(m_root + byte_list[i]), "big") for i in range(num)]
  return None


# -------------------------------------------------------------------------- FRI

def fri_fold(p, root, values, special_x):
  """The `column` of one FRI layer (starks/fri.py:236-242 via multi_interp_4)."""
  V = to_limbs(values)
  out = np.empty((len(values) // 4, 8), dtype=np.uint32)
  rc = lib().orc_fri_fold(_p(_one(p)), _p(_one(root)), _p(V), ctypes.c_uint64(len(values)),
                          _p(_one(special_x)), _p(out))
  assert rc == 0, rc
  return from_limbs(out)


def _strip(coeffs):
  """Polynomial.__init__ strips trailing zeros (starks/polynomial.py:58)."""
  c = list(coeffs)
  while c and c[-1] == 0:
    c.pop()
  return c


def fri_prove(p, f_coeffs, root, maxdeg_plus_1, exclude_multiples_of=0, security=40):
  """SmoothSubgroupFRI.generate_proximity_proof (starks/fri.py:189-266, commented out
  upstream; restored per SURVEY.md App. B)."""
  order = _order(p, root)
  values = fft_1d(p, f_coeffs, root, order=order)                       # :207-208
  if maxdeg_plus_1 <= 16:                                               # :212-214
    return [[x.to_bytes(32, "big") for x in values]]
  m = merkelize(values)                                                 # :224
  special_x = int.from_bytes(m[1], "big")                               # :229 (unreduced)
  column = fri_fold(p, root, values, special_x)                         # :236-242
  m2 = merkelize(column)                                                # :243
  ys = get_pseudorandom_indices(m2[1], len(column), security,
                                exclude_multiples_of=exclude_multiples_of)  # :246-247
  q = order // 4
  branches = []
  for y in ys:                                                          # :251-254
    branches.append([mk_branch(m2, y)] + [mk_branch(m, y + q * j) for j in range(4)])
  o = [m2[1], branches]
  root4 = pow(root, 4, p)
  column_poly = _strip(fft_1d(p, column, root4, inv=True, order=q))     # :260-261
  return [o] + fri_prove(p, column_poly, root4, maxdeg_plus_1 // 4,
                         exclude_multiples_of=exclude_multiples_of)     # :262-266


# ------------------------------------------------- polynomials (coefficient form)

def poly_mul(p, a, b):
  if not a or not b:
    return []
  out = np.empty((len(a) + len(b) - 1, 8), dtype=np.uint32)
  A, B = to_limbs(a), to_limbs(b)
  rc = lib().orc_poly_mul(_p(_one(p)), _p(A), ctypes.c_uint64(len(a)), _p(B), ctypes.c_uint64(len(b)),
                          _p(out))
  assert rc == 0
  return _strip(from_limbs(out))


def poly_divmod(p, a, b):
  a, b = _strip(a), _strip(b)
  if len(a) < len(b):
    return [], a
  A, B = to_limbs(a), to_limbs(b)
  quo = np.empty((len(a) - len(b) + 1, 8), dtype=np.uint32)
  rem = np.empty((len(a), 8), dtype=np.uint32)
  rc = lib().orc_poly_divmod(_p(_one(p)), _p(A), ctypes.c_uint64(len(a)), _p(B), ctypes.c_uint64(len(b)),
                             _p(quo), _p(rem))
  assert rc == 0, rc
  return _strip(from_limbs(quo)), _strip(from_limbs(rem))


def poly_add(p, a, b):
  n = max(len(a), len(b))
  a = list(a) + [0] * (n - len(a))
  b = list(b) + [0] * (n - len(b))
  return _strip([(x + y) % p for x, y in zip(a, b)])


def poly_sub(p, a, b):
  n = max(len(a), len(b))
  a = list(a) + [0] * (n - len(a))
  b = list(b) + [0] * (n - len(b))
  return _strip([(x - y) % p for x, y in zip(a, b)])


def poly_scale(p, a, k):
  return _strip([x * k % p for x in a])


def poly_eval(p, a, xs):
  if not a:
    return [0] * len(xs)
  A, X = to_limbs(a), to_limbs(xs)
  out = np.empty((len(xs), 8), dtype=np.uint32)
  rc = lib().orc_poly_eval(_p(_one(p)), _p(A), ctypes.c_uint64(len(a)), _p(X), ctypes.c_uint64(len(xs)),
                           _p(out))
  assert rc == 0
  return from_limbs(out)


# ------------------------------------------------------------------------ STARK

def eval_step_poly_on_polys(p, step_poly, polys):
  """MultiVarPoly.__call__ with polynomial arguments
  (starks/multivariate_polynomial.py:329-338): sum_m coeff_m * prod_i polys[i]^e_i.
  step_poly: dict {exponent tuple: int coeff}."""
  acc = []
  for exps, coeff in step_poly.items():
    term = [coeff % p]
    for i, e in enumerate(exps):
      for _ in range(e):
        term = poly_mul(p, term, polys[i])
    acc = poly_add(p, acc, term)
  return acc


def eval_step_poly_on_ints(p, step_poly, state):
  acc = 0
  for exps, coeff in step_poly.items():
    t = coeff
    for i, e in enumerate(exps):
      t = t * pow(state[i], e, p) % p
    acc = (acc + t) % p
  return acc


def step_poly_degree(step_poly):
  return max(sum(e) for e in step_poly.keys())


def computational_trace(p, inp, steps, step_polys):
  """get_computational_trace (starks/air.py:31-52) + AIR.generate_witness (:124):
  returns witness[dim][step]."""
  width = len(step_polys)
  trace = [list(inp)]
  for _ in range(steps - 1):
    trace.append([eval_step_poly_on_ints(p, step_polys[j], trace[-1]) for j in range(width)])
  return [[trace[i][j] for i in range(steps)] for j in range(width)]


class StarkOracle:
  """STARK.__init__ / mk_proof (starks/stark.py:185-279) on ints, coefficient form,
  with the reference's O(n^2) schoolbook division (starks/polynomial.py:128-143)."""

  def __init__(self, steps, extension_factor, width, step_polys, p=P_STARK):
    self.p, self.steps, self.ext, self.width = p, steps, extension_factor, width
    self.step_polys = step_polys
    self.precision = steps * extension_factor
    self.G2 = fpow(p, 7, (p - 1) // self.precision)                     # :217
    self.G1 = pow(self.G2, extension_factor, p)                         # :220
    self.last_step_position = pow(self.G2, (steps - 1) * extension_factor, p)  # :223-224

  def get_degree(self):
    return max(step_poly_degree(sp) for sp in self.step_polys)          # :230-231

  def intermediates(self, witness, boundary):
    p, steps, N = self.p, self.steps, self.precision
    last = self.last_step_position
    # construct_trace_polynomials (:27-36)
    trace_polys = [_strip(fft_1d(p, w, self.G1, inv=True, order=steps)) for w in witness]
    # construct_constraint_polynomials (:38-55): P(G1*X) has coefficients c_i*G1^i.
    next_traces = []
    for tp in trace_polys:
      g, o = 1, []
      for c in tp:
        o.append(c * g % p)
        g = g * self.G1 % p
      next_traces.append(_strip(o))
    constraint_polys = [poly_sub(p, nt, eval_step_poly_on_polys(p, sp, trace_polys))
                        for nt, sp in zip(next_traces, self.step_polys)]
    # construct_remainder_polynomials (:57-78)
    z_num = [p - 1] + [0] * (steps - 1) + [1]
    z_den = [(-last) % p, 1]
    z, r = poly_divmod(p, z_num, z_den)
    assert r == []
    ds = []
    for cp in constraint_polys:
      if not cp:
        ds.append([])
        continue
      d, r = poly_divmod(p, cp, z)
      assert r == [], "constraint polynomial not divisible by Z"
      ds.append(d)
    # construct_boundary_polynomials (:80-104)
    zeropoly2 = poly_mul(p, [p - 1, 1], [(-last) % p, 1])
    b_polys = []
    for dim in range(self.width):
      (_, _, input_value) = boundary[dim]
      output_dim = witness[dim][-1]
      # lagrange_interp_2 (starks/poly_utils.py:397-410)
      xs, ys = [1, last], [input_value % p, output_dim % p]
      eq0, eq1 = [(-xs[1]) % p, 1], [(-xs[0]) % p, 1]
      e0 = (eq0[0] + xs[0]) % p
      e1 = (eq1[0] + xs[1]) % p
      invall = pow(e0 * e1 % p, p - 2, p)
      inv_y0 = ys[0] * invall % p * e1 % p
      inv_y1 = ys[1] * invall % p * e0 % p
      interp = _strip([(eq0[i] * inv_y0 + eq1[i] * inv_y1) % p for i in range(2)])
      num = poly_sub(p, trace_polys[dim], interp)
      if not num:
        b_polys.append([])
        continue
      b, r = poly_divmod(p, num, zeropoly2)
      b_polys.append(b)
    return trace_polys, ds, b_polys

  def linear_combination(self, entropy, trace_polys, ds, b_polys):
    """compute_pseudorandom_linear_combination(_1d) (:130-177), including the leaked
    loop variable: powers[i] with i = precision-1."""
    p = self.p
    k1, k2, k3, k4 = get_pseudorandom_ks(entropy, 4)
    g2s = pow(self.G2, self.steps, p)
    c = pow(g2s, self.precision - 1, p)                                 # powers[i]
    l_polys = []
    for tp, rp, bp in zip(trace_polys, ds, b_polys):
      l = poly_add(p, rp, poly_scale(p, tp, k1 % p))
      l = poly_add(p, l, poly_scale(p, poly_scale(p, tp, k2 % p), c))
      l = poly_add(p, l, poly_scale(p, bp, k3 % p))
      l = poly_add(p, l, poly_scale(p, poly_scale(p, bp, k4 % p), c))
      l_polys.append(l)
    l_ks = get_pseudorandom_ks(entropy, self.width)
    joint = []
    for l_poly, l_k in zip(l_polys, l_ks):
      joint = poly_add(p, joint, poly_add(p, l_poly, poly_scale(p, poly_scale(p, l_poly, l_k % p), c)))
    return joint

  def mk_proof(self, witness, boundary, return_intermediates=False):
    p, N = self.p, self.precision
    trace_polys, ds, b_polys = self.intermediates(witness, boundary)
    polys = trace_polys + ds + b_polys                                  # :247
    evals = [fft_1d(p, poly, self.G2, order=N) for poly in polys]       # :254-256
    mtree = merkelize_polynomial_evaluations(evals)                     # :257
    l_poly = self.linear_combination(mtree[1], trace_polys, ds, b_polys)  # :259-261
    l_evals = fft_1d(p, l_poly, self.G2, order=N)                       # :262
    l_mtree = merkelize(l_evals)                                        # :263
    # compute_merkle_spot_checks (:390-402), samples=80
    branches = []
    positions = get_pseudorandom_indices(l_mtree[1], N, 80, exclude_multiples_of=self.ext)
    for pos in positions:
      branches.append(mk_branch(mtree, pos))
      branches.append(mk_branch(mtree, (pos + self.ext) % N))
      branches.append(mk_branch(l_mtree, pos))
    fri = fri_prove(p, l_poly, self.G2, self.steps * self.get_degree(),
                    exclude_multiples_of=self.ext)                      # :267-276
    proof = [mtree[1], l_mtree[1], branches, fri]
    if return_intermediates:
      return proof, dict(trace_polys=trace_polys, ds=ds, b_polys=b_polys, evals=evals,
                         l_poly=l_poly, l_evals=l_evals)
    return proof


def proof_digest(proof) -> str:
  """Canonical digest of a nested list-of-bytes proof object (for golden files)."""
  import hashlib
  h = hashlib.blake2s()

  def walk(x):
    if isinstance(x, (bytes, bytearray)):
      h.update(b"B" + len(x).to_bytes(4, "big") + bytes(x))
    else:
      h.update(b"L" + len(x).to_bytes(4, "big"))
      for y in x:
        walk(y)

  walk(proof)
  return h.hexdigest()


def synth(col, i, p=P_STARK):
  """Synthetic element generator of SURVEY.md section 8(d)."""
  import hashlib
  d = hashlib.blake2s(col.to_bytes(4, "little") + i.to_bytes(8, "little")).digest()
  return int.from_bytes(d, "big") % p
