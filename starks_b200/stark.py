"""Drop-in for the prover/verifier driver of starks/stark.py: STARK(field, steps,
extension_factor, width, step_polys).mk_proof(witness, boundary) / .verify_proof(proof,
witness, boundary), get_pseudorandom_ks and the construct_* helpers' results.

mk_proof keeps every column on the device from the witness upload to the last FRI layer:
LDE (stk_lde), constraint evaluations (stk_constraint_eval), quotients (stk_quotient_z,
stk_div_linear), Merkle commitments (stk_merkle_commit), linear combination (stk_lincomb),
branch gathers (stk_merkle_paths) and FRI folds (stk_fri_fold4).  Only 32-byte roots, the
Fiat-Shamir scalars and the opened branches cross to the host.  The proof object
[m_root, l_root, branches, fri_proof] (stark.py:270-277) is bit-identical to the reference's.
verify_proof is host-side glue (80 positions + FRI checks) mirroring stark.py:281-372."""
import ctypes
import os
import time
from hashlib import blake2s

import numpy as np

from .engine import P_STARK, default_engine
from .fri import FRI, DeviceLayer
from .limbs import int_to_limbs, ints_to_limbs, limbs_to_ints
from .merkle_tree import unpack_merkle_leaf, verify_branch
from .modp import element_to_int
from .polynomial import monomials_of
from .utils import get_pseudorandom_indices

blake = lambda x: blake2s(x).digest()


def get_pseudorandom_ks(m_root: bytes, num: int):
  """starks/stark.py:106-126: k_i = BLAKE2s(m_root | salt_i) as integers.  The salts are ASCII
  strings: b'0x01'.. for num <= 4, b'0x00'.. (a different numbering) for 5 <= num < 10, and the
  reference returns None for ten or more."""
  if num < 0 or num >= 10:
    return None
  first = 1 if num <= 4 else 0
  return [int.from_bytes(blake(m_root + b"0x0%d" % (first + i)), "big") for i in range(num)]


def _interp2(p, x0, x1, y0, y1):
  """lagrange_interp_2 (starks/poly_utils.py:397-410) -> coefficients [i0, i1]."""
  eq0, eq1 = [(-x1) % p, 1], [(-x0) % p, 1]
  e0, e1 = (eq0[0] + x0) % p, (eq1[0] + x1) % p
  invall = pow(e0 * e1 % p, -1, p)
  inv_y0 = y0 * invall % p * e1 % p
  inv_y1 = y1 * invall % p * e0 % p
  return [(eq0[i] * inv_y0 + eq1[i] * inv_y1) % p for i in range(2)]


class STARK(object):
  """Generates and verifies STARKs (starks/stark.py:179-402)."""

  def __init__(self, field, steps, extension_factor, width, step_polys, spot_check_security_factor=80,
               engine=None):
    self.field = field
    self.width = width
    self.steps = steps
    self.step_polys = step_polys
    self.extension_factor = extension_factor
    self.precision = steps * extension_factor
    self.spot_check_security_factor = spot_check_security_factor
    self._engine = engine
    p = self.field.p
    self.G2 = field(7)**((p - 1) // self.precision)                      # :217
    self.G1 = self.G2**extension_factor                                  # :220
    self.last_step_position = self.G2**((steps - 1) * extension_factor)  # xs[(steps-1)*ext], :223-224
    self._monomials = [monomials_of(sp, width, p) for sp in step_polys]
    self.timings = {}

  @property
  def xs(self):
    from .utils import get_power_cycle
    return get_power_cycle(self.G2, self.field, engine=self._engine)

  def get_degree(self):
    return max(max([sum(k) for k, _ in m] or [0]) for m in self._monomials)  # :230-231

  # ----------------------------------------------------------------- prover
  def _witness_limbs(self, witness):
    """-> (width, steps, 8) uint32 host array, or the device buffer itself when the witness is
    already on the device (DevBuf from air.witness_device / stk_trace_generate_dev)."""
    p = self.field.p
    if hasattr(witness, "ptr") and hasattr(witness, "nbytes"):
      assert witness.nbytes >= self.width * self.steps * 32
      return witness
    if isinstance(witness, np.ndarray):
      assert witness.shape == (self.width, self.steps, 8)
      return np.ascontiguousarray(witness, dtype=np.uint32)
    assert len(witness) == self.width
    return np.stack([ints_to_limbs([element_to_int(v) % p for v in col]) for col in witness])

  def _monomial_arrays(self):
    mono_out, mono_coef, mono_exp = [], [], []
    for j, ms in enumerate(self._monomials):
      for exps, c in ms:
        mono_out.append(j)
        mono_coef.append(c)
        mono_exp.append(list(exps))
    nm, w = len(mono_out), self.width
    h_out = np.asarray(mono_out, dtype=np.uint32)
    h_coef = ints_to_limbs(mono_coef) if nm else np.zeros((0, 8), np.uint32)
    h_exp = np.asarray(mono_exp, dtype=np.uint8).reshape(nm, w) if nm else np.zeros((0, w), np.uint8)
    return nm, h_out, h_coef, h_exp

  def coefficient_rows(self, eng, d_trace_ptr, boundary, out_vals, want=None):
    """The 3w polynomials mk_proof commits to -- P_1..P_w (stark.py:27-36), D_1..D_w (:38-78),
    B_1..B_w (:80-104), in the leaf order of :247 -- as COEFFICIENT rows on the device: returns
    (buffer, row stride cs), row r at byte r*cs*32, zero-padded to cs >= deg + 1.  The sharded
    prover (dist.ShardedProver) evaluates slices of these rows on different GPUs; mk_proof itself
    takes shortcuts (pointwise quotients) that this form does not need.
    want: the rows the caller will read (global row numbers 0..3w-1; default all).  Only those --
    and what they depend on -- are computed: a rank that evaluates one column of a Fibonacci proof
    runs one or two transforms of size `steps` here instead of four."""
    p = self.field.p
    w, steps, ext, N = self.width, self.steps, self.extension_factor, self.precision
    G2, last = int(self.G2), int(self.last_step_position)
    G1 = pow(G2, ext, p)
    E = 32
    want = set(range(3 * w)) if want is None else set(int(c) for c in want)
    nm, h_out, h_coef, h_exp = self._monomial_arrays()
    mult = 1
    while mult < max(self.get_degree(), 1):
      mult *= 2
    M = min(N, steps * mult)
    GM = pow(G2, N // M, p)
    cs = M
    want_d = sorted(c - w for c in want if w <= c < 2 * w)
    want_b = sorted(c - 2 * w for c in want if c >= 2 * w)
    # P's coefficients: for its own rows, for B_j, and -- unless the constraint subgroup is the trace
    # domain itself -- for every column (C_j needs all of P on that subgroup)
    need_p = set(c for c in want if c < w) | set(want_b)
    if want_d and M != steps:
      need_p = set(range(w))
    d_coef = eng.alloc(3 * w * cs * E)
    eng._check(eng.lib.stk_memset(eng.ctx, d_coef.ptr, 0, 3 * w * cs * E))
    if len(need_p) == w:
      eng.ntt(d_trace_ptr, steps, steps, d_coef.ptr, cs, steps, w, G1, inverse=True)
    else:
      for j in sorted(need_p):
        eng.ntt(d_trace_ptr + j * steps * E, steps, steps, d_coef.at(j * cs * E), cs, steps, 1, G1, inverse=True)
    held = []
    bad = ctypes.c_uint32(0)
    last_l, one_l = int_to_limbs(last), int_to_limbs(1)
    u32p = ctypes.POINTER(ctypes.c_uint32)
    d_t1 = eng.alloc(w * M * E)
    held.append(d_t1)
    if want_d:
      if M == steps:
        pev_ptr, pev_stride = d_trace_ptr, steps     # P_j on <G1> is the witness column itself
      else:
        d_t2 = eng.alloc(w * M * E)
        held.append(d_t2)
        eng.ntt(d_coef.ptr, steps, cs, d_t2.ptr, M, M, w, GM)
        pev_ptr, pev_stride = d_t2.ptr, M
      eng._check(eng.lib.stk_constraint_eval(eng.ctx, pev_ptr, M, M // steps, w, pev_stride, h_out.ctypes.data,
                                             h_coef.ctypes.data, h_exp.ctypes.data, nm, d_t1.ptr, M))
      d_t3 = eng.alloc(w * M * E)
      held.append(d_t3)
      if len(want_d) == w:
        eng.ntt(d_t1.ptr, M, M, d_t3.ptr, M, M, w, GM, inverse=True)
      else:
        for j in want_d:
          eng.ntt(d_t1.at(j * M * E), M, M, d_t3.at(j * M * E), M, M, 1, GM, inverse=True)
      for j in want_d:
        eng._check(eng.lib.stk_quotient_z(eng.ctx, d_t3.at(j * M * E), M, steps, last_l.ctypes.data_as(u32p),
                                          d_coef.at((w + j) * cs * E), ctypes.byref(bad)))
        assert bad.value == 0, "constraint polynomial is not divisible by Z (stark.py:74-75)"
    if want_b:
      interps = []
      for j in range(w):
        (_, _, input_value) = boundary[j]
        interps += _interp2(p, 1, last, element_to_int(input_value) % p, out_vals[j])
      d_i = eng.alloc(2 * w * E).upload(ints_to_limbs(interps))
      held.append(d_i)
      for j in want_b:
        a = d_coef.at((2 * w + j) * cs * E)
        eng._check(eng.lib.stk_memcpy_d2d(eng.ctx, a, d_coef.at(j * cs * E), steps * E))
        eng._check(eng.lib.stk_vec_op(eng.ctx, 1, a, d_i.at(2 * j * E), a, 2))
        q1 = d_t1.at(j * steps * E)      # free again: its last reader is already enqueued on the stream
        eng._check(eng.lib.stk_div_linear(eng.ctx, a, steps, one_l.ctypes.data_as(u32p), 1, q1))
        eng._check(eng.lib.stk_div_linear(eng.ctx, q1, steps - 1, last_l.ctypes.data_as(u32p), steps, a))
        eng._check(eng.lib.stk_memset(eng.ctx, a + (steps - 2) * E, 0, 2 * E))
    eng.sync()   # the scratch rows go back to the pool
    for b in held:
      b.free()
    return d_coef, cs

  def mk_proof(self, witness, boundary, keep_device=False):
    """stark.py:233-279."""
    t_start = time.time()
    marks = [("start", t_start)]
    mark = lambda name: marks.append((name, time.time()))
    eng = self._engine or default_engine()
    p = self.field.p
    eng.set_field(p)
    w, steps, ext, N = self.width, self.steps, self.extension_factor, self.precision
    G2, last = int(self.G2), int(self.last_step_position)
    tr = self._witness_limbs(witness)
    E = 32
    on_device = not isinstance(tr, np.ndarray)
    if on_device:
      d_trace = tr
    else:
      assert tr.shape[1] == steps
      d_trace = eng.alloc(w * steps * E).upload(tr, wait=False)   # `tr` lives until the syncs below
    if on_device or isinstance(witness, np.ndarray):
      # caller-supplied limbs are used as they are: the kernels' add/sub/reduce assume canonical
      # residues (the list path reduces mod p like IntegerModP.__init__, modp.py:35-36)
      noncanonical = eng.count_noncanonical(d_trace.ptr, w * steps, wait=False)
    else:
      noncanonical = None
    d_cols = eng.alloc(3 * w * N * E)        # rows: P_1..P_w, D_1..D_w, B_1..B_w (stark.py:247)
    d_t1 = eng.alloc(w * N * E)
    d_t2 = eng.alloc(w * N * E)
    mark("upload")
    # construct_constraint_polynomials (:38-55), evaluation form -- on the SMALLEST subgroup
    # that determines C: deg C <= d*(steps-1) < M = steps*2^ceil(log2 d), so C (and D = C/Z)
    # are recovered exactly from M evaluations; only D's final evaluation runs at size N.
    # On <G_M>, G_M = G2^(N/M), P_j(G1*x_k) is the evaluation M/steps places further on.
    mono_out, mono_coef, mono_exp = [], [], []
    for j, ms in enumerate(self._monomials):
      for exps, c in ms:
        mono_out.append(j)
        mono_coef.append(c)
        mono_exp.append(list(exps))
    nm = len(mono_out)
    h_out = np.asarray(mono_out, dtype=np.uint32)
    h_coef = ints_to_limbs(mono_coef) if nm else np.zeros((0, 8), np.uint32)
    h_exp = np.asarray(mono_exp, dtype=np.uint8).reshape(nm, w) if nm else np.zeros((0, w), np.uint8)
    mult = 1
    while mult < max(self.get_degree(), 1):
      mult *= 2
    M = min(N, steps * mult)
    GM = pow(G2, N // M, p)
    # When every coefficient vector is short enough for the zero-padded transform (M <= N/8) the
    # 3w columns P, D, B are evaluated by ONE batched transform at the end; otherwise P goes
    # first (the constraints may need its size-N evaluations) and D, B follow on their own.
    merged = M * 8 <= N
    cs = M if merged else steps                 # coefficient row stride
    d_coef = eng.alloc((3 * w if merged else w) * cs * E)
    if merged and cs > steps:
      eng._check(eng.lib.stk_memset(eng.ctx, d_coef.ptr, 0, 3 * w * cs * E))
    # D's evaluations can be taken pointwise from P's wherever Z does not vanish (stk_quotient_eval):
    # then only P and B go through the size-N transform
    pointwise_d = M < N and 2 <= ext <= 16 and os.environ.get("STK_PROOF_POINTWISE", "1") != "0"
    if (not merged or pointwise_d) and ext == 8 and p == P_STARK:
      # construct_trace_polynomials (:27-36) and the evaluation of P (:254-256) in one call: with an
      # 8x blowup the evaluations at every 8th point ARE the trace, so stk_lde copies that coset
      # instead of computing it (one eighth of the forward transform)
      eng.lde(d_trace.ptr, steps, steps, ext, w, G2, d_cols.ptr, N, d_coeffs=d_coef.ptr, coeff_stride=cs)
    else:
      # construct_trace_polynomials (:27-36): inverse transform over <G1>
      eng.ntt(d_trace.ptr, steps, steps, d_coef.ptr, cs, steps, w, pow(G2, ext, p), inverse=True)
      if not merged or pointwise_d:
        eng.ntt(d_coef.ptr, steps, cs, d_cols.ptr, N, N, w, G2)          # evaluation of P (:254-256)
    if M == N:
      pev_ptr, pev_stride = d_cols.ptr, N
    elif M == steps:
      pev_ptr, pev_stride = d_trace.ptr, steps   # P_j on <G1> is the witness column itself
    else:
      eng.ntt(d_coef.ptr, steps, cs, d_t2.ptr, M, M, w, GM)
      pev_ptr, pev_stride = d_t2.ptr, M
    eng._check(eng.lib.stk_constraint_eval(eng.ctx, pev_ptr, M, M // steps, w, pev_stride, h_out.ctypes.data,
                                           h_coef.ctypes.data, h_exp.ctypes.data, nm, d_t1.ptr, M))
    # construct_remainder_polynomials (:57-78): D = C / Z in coefficient form
    eng.ntt(d_t1.ptr, M, M, d_t2.ptr, M, M, w, GM, inverse=True)
    bad = ctypes.c_uint32(0)
    last_l = int_to_limbs(last)
    u32p = ctypes.POINTER(ctypes.c_uint32)
    for j in range(w):
      dst = d_coef.at((w + j) * cs * E) if merged else d_t1.at(j * M * E)
      eng._check(eng.lib.stk_quotient_z(eng.ctx, d_t2.at(j * M * E), M, steps, last_l.ctypes.data_as(u32p),
                                        dst, ctypes.byref(bad)))
      if noncanonical is not None:   # stk_quotient_z synchronised the stream: the count is in
        if noncanonical():
          raise ValueError("witness holds elements >= p; reduce them mod p first")
        noncanonical = None
      assert bad.value == 0, "constraint polynomial is not divisible by Z (stark.py:74-75)"
    if pointwise_d:
      # D on <G_M> (contains <G1>, where Z vanishes) from its coefficients, everything else pointwise
      dco, dco_stride, dsub = (d_coef.at(w * cs * E), cs, d_t1) if merged else (d_t1.ptr, M, d_t2)
      eng.ntt(dco, M, dco_stride, dsub.ptr, M, M, w, GM)
      eng._check(eng.lib.stk_quotient_eval(eng.ctx, d_cols.ptr, N, ext, w, N, h_out.ctypes.data, h_coef.ctypes.data,
                                           h_exp.ctypes.data, nm, int_to_limbs(G2).ctypes.data_as(u32p),
                                           last_l.ctypes.data_as(u32p), dsub.ptr, M, M, d_cols.at(w * N * E), N))
    elif not merged:
      eng.ntt(d_t1.ptr, M, M, d_cols.at(w * N * E), N, N, w, G2)
    # construct_boundary_polynomials (:80-104): B = (P - I) / ((X - 1)(X - last))
    if on_device:
      last_rows = np.stack([d_trace.download((1, 8), byte_offset=(j * steps + steps - 1) * E)[0] for j in range(w)])
    else:
      last_rows = tr[:, -1, :]
    out_vals = limbs_to_ints(last_rows)
    one_l = int_to_limbs(1)
    interps = []
    for j in range(w):
      (_, _, input_value) = boundary[j]
      interps += _interp2(p, 1, last, element_to_int(input_value) % p, out_vals[j])
    d_i = eng.alloc(2 * w * E).upload(ints_to_limbs(interps))
    for j in range(w):
      a = d_coef.at((2 * w + j) * cs * E) if merged else d_t2.at(j * steps * E)
      eng._check(eng.lib.stk_memcpy_d2d(eng.ctx, a, d_coef.at(j * cs * E), steps * E))
      eng._check(eng.lib.stk_vec_op(eng.ctx, 1, a, d_i.at(2 * j * E), a, 2))
      q1 = d_t1.at(j * steps * E)   # free again: its last reader is already enqueued on the stream
      eng._check(eng.lib.stk_div_linear(eng.ctx, a, steps, one_l.ctypes.data_as(u32p), 1, q1))
      eng._check(eng.lib.stk_div_linear(eng.ctx, q1, steps - 1, last_l.ctypes.data_as(u32p), steps, a))
      if merged:   # the quotient has steps-2 coefficients; clear the two stale ones above it
        eng._check(eng.lib.stk_memset(eng.ctx, a + (steps - 2) * E, 0, 2 * E))
    if pointwise_d and p == P_STARK and steps >= 4:
      # B on <G1> (holds x = 1 and x = last, where the denominator vanishes) from its coefficients,
      # everything else pointwise from P's evaluations (stk_boundary_eval)
      bco, bco_stride, bsub = (d_coef.at(2 * w * cs * E), cs, d_t2) if merged else (d_t2.ptr, steps, d_t1)
      eng.ntt(bco, steps - 2, bco_stride, bsub.ptr, steps, steps, w, pow(G2, ext, p))
      h_interp = ints_to_limbs(interps)
      eng._check(eng.lib.stk_boundary_eval(eng.ctx, d_cols.ptr, N, ext, w, N, int_to_limbs(G2).ctypes.data_as(u32p),
                                           (steps - 1) * ext, h_interp.ctypes.data, bsub.ptr, steps,
                                           d_cols.at(2 * w * N * E), N))
    elif pointwise_d and merged:
      eng.ntt(d_coef.at(2 * w * cs * E), cs, cs, d_cols.at(2 * w * N * E), N, N, w, G2)   # evaluation of B
    elif merged:
      eng.ntt(d_coef.ptr, cs, cs, d_cols.ptr, N, N, 3 * w, G2)           # evaluation of P, D, B (:254-256)
    else:
      eng.ntt(d_t2.ptr, steps - 2, steps, d_cols.at(2 * w * N * E), N, N, w, G2)
    mark("enqueue_polys")
    # merkelize_polynomial_evaluations (:257)
    d_mnodes = eng.alloc(32 * N)
    m_root = eng.merkle_commit(d_cols.ptr, N, 3 * w, N, d_mnodes.ptr)
    mark("m_root")
    # compute_pseudorandom_linear_combination (:130-177), evaluation form
    k1, k2, k3, k4 = get_pseudorandom_ks(m_root, 4)
    l_ks = get_pseudorandom_ks(m_root, w)
    c = pow(pow(G2, steps, p), N - 1, p)     # powers[i] with the leaked i = precision-1 (:153-160)
    wP, wD, wB = [], [], []
    for j in range(w):
      aj = (1 + l_ks[j] * c) % p
      wD.append(aj)
      wP.append(aj * ((k1 + k2 * c) % p) % p)
      wB.append(aj * ((k3 + k4 * c) % p) % p)
    weights = ints_to_limbs(wP + wD + wB)
    d_l = eng.alloc(N * E)
    eng._check(eng.lib.stk_lincomb(eng.ctx, d_cols.ptr, N, 3 * w, N, weights.ctypes.data, d_l.ptr))
    d_lnodes = eng.alloc(32 * N)
    l_root = eng.merkle_commit(d_l.ptr, N, 1, N, d_lnodes.ptr)           # :262-263
    mark("l_root")
    # compute_merkle_spot_checks (:390-402), samples = 80
    positions = get_pseudorandom_indices(l_root, N, 80, exclude_multiples_of=ext)
    mb = eng.merkle_paths(d_cols.ptr, N, 3 * w, N, d_mnodes.ptr,
                          [x for pos in positions for x in (pos, (pos + ext) % N)])
    lb = eng.merkle_paths(d_l.ptr, N, 1, N, d_lnodes.ptr, positions)
    branches = []
    for i in range(len(positions)):
      branches += [mb[2 * i], mb[2 * i + 1], lb[i]]
    mark("spot_checks")
    # FRI on l (:267-276); its first layer's tree is l_mtree
    fri = FRI(self.field, engine=eng)
    fri_proof = fri.prove_from_device(DeviceLayer(eng, d_l.ptr, N, d_lnodes.ptr, l_root), G2,
                                      steps * self.get_degree(), exclude_multiples_of=ext)
    proof = [m_root, l_root, branches, fri_proof]
    mark("fri")
    self.timings = {"mk_proof_s": time.time() - t_start}
    for (a, ta), (b, tb) in zip(marks[:-1], marks[1:]):
      self.timings[b + "_ms"] = (tb - ta) * 1e3
    if keep_device:
      self.device = dict(cols=d_cols, pcoef=d_coef, l=d_l, mnodes=d_mnodes, lnodes=d_lnodes)
    else:
      for b in (d_coef, d_i, d_cols, d_t1, d_t2, d_mnodes, d_l, d_lnodes) + (() if on_device else (d_trace,)):
        b.free()
    return proof

  # --------------------------------------------------------------- verifier
  def verify_proof(self, proof, witness, boundary):
    """stark.py:281-317 (host-side; the reference hands the verifier the witness, :370)."""
    m_root, l_root, branches, fri_proof = proof
    fri = FRI(self.field, engine=self._engine)
    assert fri.verify_proximity_proof(fri_proof, l_root, self.G2, self.steps * self.get_degree(),
                                      exclude_multiples_of=self.extension_factor)
    samples = self.spot_check_security_factor
    positions = get_pseudorandom_indices(l_root, self.precision, samples,
                                         exclude_multiples_of=self.extension_factor)
    ks = get_pseudorandom_ks(m_root, 4)
    N, ext = self.precision, self.extension_factor
    leaves = None
    if self._engine is not None and N >= 4 and N & (N - 1) == 0:
      # all 3 * samples branches re-hashed in two kernels (stk_verify_branches)
      ml = self._engine.verify_branches(m_root, [x for pos in positions for x in (pos, (pos + ext) % N)],
                                        [branches[3 * i + d] for i in range(len(positions)) for d in (0, 1)])
      self._engine.verify_branches(l_root, positions, [branches[3 * i + 2] for i in range(len(positions))])
      leaves = ml
    for i, pos in enumerate(positions):
      self.verify_proof_at_position(witness, boundary, ks, proof, i, pos,
                                    checked_leaves=None if leaves is None else leaves[2 * i:2 * i + 2])
    return True

  def _step(self, j, state, p):
    acc = 0
    for exps, c in self._monomials[j]:
      t = c
      for k, e in enumerate(exps):
        t = t * pow(state[k], e, p) % p
      acc = (acc + t) % p
    return acc

  def verify_proof_at_position(self, witness, boundary, ks, proof, i, pos, checked_leaves=None):
    """stark.py:319-372.  checked_leaves: the two m-tree leaves of this position when their
    branches (and the l-tree branch) have already been verified in a batch."""
    p, width = self.field.p, self.width
    m_root, l_root, branches, fri_proof = proof
    G2, last = int(self.G2), int(self.last_step_position)
    x = pow(G2, pos, p)
    if checked_leaves is not None:
      leaf1, leaf2 = (unpack_merkle_leaf(v, width, 3) for v in checked_leaves)
    else:
      leaf1 = unpack_merkle_leaf(verify_branch(m_root, pos, branches[i * 3]), width, 3)
      leaf2 = unpack_merkle_leaf(verify_branch(m_root, (pos + self.extension_factor) % self.precision,
                                               branches[i * 3 + 1]), width, 3)
      verify_branch(l_root, pos, branches[i * 3 + 2], output_as_int=True)
    f_ = lambda b: int.from_bytes(b, "big")   # field(bytes) does not reduce; values are canonical
    p_of_x = [f_(v) for v in leaf1[:width]]
    p_of_g1x = [f_(v) for v in leaf2[:width]]
    d_of_x = [f_(v) for v in leaf1[width:2 * width]]
    b_of_x = [f_(v) for v in leaf1[2 * width:]]
    zvalue = (pow(x, self.steps, p) - 1) * pow((x - last) % p, -1, p) % p
    for dim in range(width):                                             # transition constraints
      assert (p_of_g1x[dim] - self._step(dim, p_of_x, p) - zvalue * d_of_x[dim]) % p == 0
    zeropoly2_x = (x - 1) * (x - last) % p
    for dim in range(width):                                             # boundary constraints
      (_, _, input_value) = boundary[dim]
      if hasattr(witness, "ptr") and hasattr(witness, "download"):
        output_dim = limbs_to_ints(witness.download((1, 8), byte_offset=(dim * self.steps + self.steps - 1) * 32))[0]
      elif isinstance(witness, np.ndarray):
        output_dim = limbs_to_ints(witness[dim, -1:, :])[0]
      else:
        output_dim = element_to_int(witness[dim][-1]) % p
      i0, i1 = _interp2(p, 1, last, element_to_int(input_value) % p, output_dim)
      assert (p_of_x[dim] - b_of_x[dim] * zeropoly2_x - (i0 + i1 * x)) % p == 0
