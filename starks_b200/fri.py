"""Drop-in for the multiplicative-subgroup FRI of starks/fri.py:176-366 (commented out at
upstream HEAD): SmoothSubgroupFRI / FRI with generate_proximity_proof and
verify_proximity_proof.

Prover: evaluations stay on the device across layers; per layer one fold kernel
(stk_fri_fold4), one Merkle commit and two batched branch gathers.  The reference recomputes
each layer's values by an inverse + forward transform of the previous column and re-builds
its tree (fri.py:207-224, 260-266); both are identities (layer k+1's values == layer k's
column, its tree == layer k's m2), so they are skipped -- the proof object is bit-identical.
Verifier: host-side, O(queries * layers), mirrors fri.py:268-366."""
import numpy as np

from .engine import default_engine
from .limbs import ints_to_limbs, limbs_to_be_bytes, limbs_to_ints
from .merkle_tree import verify_branch
from .modp import element_to_int
from .utils import get_pseudorandom_indices, multiplicative_order


class DeviceLayer(object):
  """One FRI layer's evaluations (+ Merkle nodes) resident in device memory."""

  def __init__(self, eng, d_vals, n, d_nodes=None, root=None, owner=None):
    self.eng, self.d_vals, self.n, self.d_nodes, self.root, self.owner = eng, d_vals, n, d_nodes, root, owner


class SmoothSubgroupFRI(object):
  """Fast Reed-Solomon IOPP for a smooth multiplicative subgroup (starks/fri.py:176-187)."""

  def __init__(self, field, engine=None):
    self.field = field
    self._engine = engine

  # ------------------------------------------------------------------ prover
  def generate_proximity_proof(self, f, root_of_unity, maxdeg_plus_1, exclude_multiples_of=0,
                               fri_spot_check_security_factor=40):
    """fri.py:189-266.  f: polynomial (object with .coefficients) or coefficient list."""
    eng = self._engine or default_engine()
    p = self.field.p
    eng.set_field(p)
    root = element_to_int(root_of_unity) % p
    n = multiplicative_order(root, p)
    coeffs = f.coefficients if hasattr(f, "coefficients") else list(f)
    if len(coeffs) > n:
      raise IndexError("list index out of range")
    c = ints_to_limbs([element_to_int(v) % p for v in coeffs]) if len(coeffs) else np.zeros((0, 8), np.uint32)
    d_in = eng.alloc(max(c.nbytes, 32)).upload(c)
    d_vals = eng.alloc(n * 32)
    eng.ntt(d_in.ptr, len(coeffs), max(len(coeffs), 1), d_vals.ptr, n, n, 1, root)   # :207-208
    layer = DeviceLayer(eng, d_vals.ptr, n, owner=[d_in, d_vals])
    return self.prove_from_device(layer, root, maxdeg_plus_1, exclude_multiples_of, fri_spot_check_security_factor)

  def prove_from_device(self, layer, root, maxdeg_plus_1, exclude_multiples_of=0, security=40, use_driver=True):
    """Same proof, starting from evaluations already on the device (and optionally their
    tree, as STARK.mk_proof has just built it for l_evaluations)."""
    eng = layer.eng
    p = self.field.p
    proof = []
    keep = list(layer.owner or [])
    d_vals, n, d_nodes, m_root = layer.d_vals, layer.n, layer.d_nodes, layer.root
    nn, md = n, maxdeg_plus_1
    fits = use_driver and n & (n - 1) == 0 and maxdeg_plus_1 > 16 and exclude_multiples_of != 1
    while fits and md > 16:
      fits, nn, md = nn >= 16 and nn // 4 < 2**24, nn // 4, md // 4
    if fits:
      # the whole commit phase in one library call (stk_fri_prove): same kernels as the loop
      # below, one host synchronisation per layer and one download of the opened branches
      proof = eng.fri_prove(d_vals, n, d_nodes, m_root, root, maxdeg_plus_1, exclude_multiples_of, security)
      for b in keep:
        b.free()
      return proof
    while True:
      if maxdeg_plus_1 <= 16:                                            # :212-214
        vals = _download(eng, d_vals, n)
        be = limbs_to_be_bytes(vals)
        proof.append([be[i].tobytes() for i in range(n)])
        break
      if d_nodes is None:                                                # m = merkelize(values), :224
        nodes_buf = eng.alloc(32 * n)
        keep.append(nodes_buf)
        d_nodes = nodes_buf.ptr
        m_root = eng.merkle_commit(d_vals, n, 1, n, d_nodes)
      special_x = int.from_bytes(m_root, "big")                         # :229, unreduced
      q = n // 4
      col_buf, col_nodes = eng.alloc(max(q, 1) * 32), eng.alloc(32 * max(q, 1))
      keep += [col_buf, col_nodes]
      eng.fri_fold4(d_vals, n, root, special_x, col_buf.ptr)             # :236-242
      root2 = eng.merkle_commit(col_buf.ptr, q, 1, q, col_nodes.ptr)     # :243
      ys = get_pseudorandom_indices(root2, q, security, exclude_multiples_of=exclude_multiples_of)  # :246-247
      b2 = eng.merkle_paths(col_buf.ptr, q, 1, q, col_nodes.ptr, ys)
      b1 = eng.merkle_paths(d_vals, n, 1, n, d_nodes, [y + q * j for y in ys for j in range(4)])
      branches = [[b2[i]] + b1[4 * i:4 * i + 4] for i in range(len(ys))]  # :251-254
      proof.append([root2, branches])
      # next layer (:256-266): its values are this column, its tree this m2
      d_vals, n, d_nodes, m_root = col_buf.ptr, q, col_nodes.ptr, root2
      root = pow(root, 4, p)
      maxdeg_plus_1 //= 4
      security = 40  # the reference's recursive call does not forward the argument (:262-266)
    for b in keep:
      b.free()
    return proof

  # ---------------------------------------------------------------- verifier
  def verify_proximity_proof(self, proof, merkle_root, root_of_unity, maxdeg_plus_1, exclude_multiples_of=0,
                             fri_spot_check_security_factor=40):
    """fri.py:268-366 on Python ints (host side; 40 queries per layer)."""
    p = self.field.p
    root = element_to_int(root_of_unity) % p
    roudeg = multiplicative_order(root, p)                               # :277-283
    for prf in proof[:-1]:
      root2, branches = prf
      special_x = int.from_bytes(merkle_root, "big") % p                 # :295
      ys = get_pseudorandom_indices(root2, roudeg // 4, fri_spot_check_security_factor,
                                    exclude_multiples_of=exclude_multiples_of)
      quartic = [pow(root, roudeg * j // 4, p) for j in range(4)]        # :285-291
      q = roudeg // 4
      iota_inv, inv4 = quartic[3] if roudeg % 4 == 0 else None, pow(4, -1, p)
      if self._engine is not None and q >= 4 and q & (q - 1) == 0:
        # every branch of the layer re-hashed in two kernels (stk_verify_branches)
        f_ = lambda b: int.from_bytes(b, "big")
        rows_all = [f_(v) for v in self._engine.verify_branches(
            merkle_root, [y + q * j for y in ys for j in range(4)], [br for i in range(len(ys)) for br in branches[i][1:]])]
        cols_all = [f_(v) for v in self._engine.verify_branches(root2, ys, [branches[i][0] for i in range(len(ys))])]
      else:
        rows_all = cols_all = None
      for i, y in enumerate(ys):
        x1 = pow(root, y, p)
        xs = [quartic[j] * x1 % p for j in range(4)]
        if rows_all is not None:
          row, col = rows_all[4 * i:4 * i + 4], cols_all[i]
        else:
          row = [verify_branch(merkle_root, y + q * j, br, output_as_int=True)
                 for j, br in zip(range(4), branches[i][1:])]
          col = verify_branch(root2, y, branches[i][0], output_as_int=True)
        if iota_inv is not None:
          # the degree<4 interpolant through (x1*iota^j, row[j]) at special_x in closed form (what
          # stk_fri_fold4 computes): no modular inverse -- x1^-1 = root^(roudeg - y)
          t = special_x * pow(root, roudeg - y, p) % p
          assert _fold4_eval(row, t, iota_inv, inv4, p) == col % p        # :330-333
        else:
          assert _lagrange_eval(xs, row, special_x, p) == col % p
      merkle_root = root2
      root = pow(root, 4, p)
      maxdeg_plus_1 //= 4
      roudeg //= 4
    data = [int.from_bytes(x, "big") for x in proof[-1]]                 # :342
    assert maxdeg_plus_1 <= 16
    assert _host_merkle_root(data) == merkle_root                        # :346-348
    powers = [pow(root, i, p) for i in range(len(data))]
    pts = [x for x in range(len(data)) if x % exclude_multiples_of] if exclude_multiples_of else list(range(len(data)))
    xs = [powers[x] for x in pts[:maxdeg_plus_1]]
    ys_ = [data[x] % p for x in pts[:maxdeg_plus_1]]
    if len(pts) > maxdeg_plus_1:
      ws = _interp_weights(xs, p)
    for x in pts[maxdeg_plus_1:]:                                        # :357-362
      assert _weighted_eval(xs, ws, ys_, powers[x], p) == data[x] % p
    return True


def _host_merkle_root(data):
  """Root of merkelize(data) for the verifier's final check (a few hundred 32-byte leaves:
  host hashing, like every other verifier step)."""
  from .merkle_tree import blake, permute4
  nodes = [x.to_bytes(32, "big") for x in permute4(list(data))]
  n = len(nodes)
  tree = [b""] * n + nodes
  for i in range(n - 1, 0, -1):
    tree[i] = blake(tree[2 * i] + tree[2 * i + 1])
  return tree[1]


def _fold4_eval(row, t, iota_inv, inv4, p):
  """1/4 * sum_k t^k * sum_j row[j] * iota^(-jk): the value at special_x = t*x1 of the cubic through
  (x1*iota^j, row[j]) -- multi_interp_4 + Polynomial.__call__ (starks/poly_utils.py:412-440) in
  closed form; exact arithmetic, so equal to the interpolation route."""
  v0, v1, v2, v3 = row
  s02, d02, s13, d13 = v0 + v2, v0 - v2, v1 + v3, v1 - v3
  m = d13 * iota_inv % p
  c0, c1, c2, c3 = s02 + s13, d02 + m, s02 - s13, d02 - m
  return (((c3 * t + c2) % p * t + c1) % p * t + c0) % p * inv4 % p


def _interp_weights(xs, p):
  """1 / prod_{j != i} (x_i - x_j) for every i, with one modular inversion (multi_inv,
  starks/poly_utils.py:301-320)."""
  dens = []
  for i, xi in enumerate(xs):
    d = 1
    for j, xj in enumerate(xs):
      if i != j:
        d = d * (xi - xj) % p
    dens.append(d)
  pref = [1]
  for d in dens:
    pref.append(pref[-1] * d % p)
  inv = pow(pref[-1], -1, p)
  out = [0] * len(xs)
  for i in range(len(xs) - 1, -1, -1):
    out[i] = pref[i] * inv % p
    inv = inv * dens[i] % p
  return out


def _weighted_eval(xs, ws, ys, x, p):
  """Value at x of the interpolant through (xs, ys) given the weights above: sum_i y_i w_i
  prod_{j != i} (x - x_j), the products by prefix / suffix (no inversion per point)."""
  n = len(xs)
  pre, suf = [1] * (n + 1), [1] * (n + 1)
  for i in range(n):
    pre[i + 1] = pre[i] * (x - xs[i]) % p
  for i in range(n - 1, -1, -1):
    suf[i] = suf[i + 1] * (x - xs[i]) % p
  return sum(ys[i] * ws[i] % p * pre[i] % p * suf[i + 1] for i in range(n)) % p


def _lagrange_eval(xs, ys, x, p):
  """Value at x of the interpolant through (xs, ys) (what multi_interp_4 /
  lagrange_interp followed by Polynomial.__call__ compute, starks/poly_utils.py:337-440)."""
  total = 0
  for i, (xi, yi) in enumerate(zip(xs, ys)):
    num, den = 1, 1
    for j, xj in enumerate(xs):
      if i != j:
        num = num * (x - xj) % p
        den = den * (xi - xj) % p
    total = (total + yi * num % p * pow(den, -1, p)) % p
  return total


def _download(eng, d_ptr, n):
  out = np.empty((n, 8), dtype=np.uint32)
  eng._check(eng.lib.stk_memcpy_d2h(eng.ctx, out.ctypes.data, d_ptr, out.nbytes))
  return out


FRI = SmoothSubgroupFRI
