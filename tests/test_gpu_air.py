"""Trace generation (starks/air.py:31-52, :124) through stk_trace_generate vs the oracle's
restatement, and the witness -> proof -> verify round trip on it.  Bit-exact."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

P = 2**256 - 351 * 2**32 + 1


@pytest.fixture(scope="module")
def eng():
  from starks_b200 import Engine
  e = Engine(0)
  yield e
  e.close()


CASES = [
    ("fibonacci", [0, 1], [{(0, 1): 1}, {(1, 0): 1, (0, 1): 1}]),
    ("quadratic", [2, 3], [{(0, 1): 1}, {(1, 0): 1, (0, 2): 1}]),
    ("cubic+const", [5], [{(3,): 1, (0,): 42}]),
    ("3-wide mixed", [1, 2, 3], [{(0, 1, 0): 1}, {(0, 0, 1): 1}, {(1, 1, 0): 7, (0, 0, 2): P - 1, (0, 0, 0): 9}]),
]


@pytest.mark.parametrize("name,inp,sp", CASES, ids=[c[0] for c in CASES])
def test_trace_matches_oracle(eng, oracle, name, inp, sp):
  from starks_b200.air import witness_limbs, get_computational_trace, generate_witness
  from starks_b200.modp import IntegersModP
  F = IntegersModP(P)
  width = len(inp)
  for steps in (1, 2, 64, 1000):
    want = oracle.computational_trace(P, inp, steps, sp)
    got = witness_limbs(F, inp, steps, width, sp, engine=eng)
    assert got.shape == (width, steps, 8)
    for j in range(width):
      assert oracle.from_limbs(got[j]) == want[j], (name, steps, j)
  trace, output = get_computational_trace([F(v) for v in inp], 16, width, sp, field=F, engine=eng)
  want = oracle.computational_trace(P, inp, 16, sp)
  assert [[int(x) for x in col] for col in generate_witness(trace)] == want
  assert [int(x) for x in output] == [want[j][-1] for j in range(width)]
  assert isinstance(trace[3][0], F)


def test_generated_witness_proves_and_verifies(eng, oracle):
  from starks_b200.air import witness_limbs
  from starks_b200.modp import IntegersModP
  from starks_b200.stark import STARK
  F = IntegersModP(P)
  sp = [{(0, 1): 1}, {(1, 0): 1, (0, 1): 1}]
  steps = 1 << 10
  pin = eng.pinned((2, steps, 8))
  w = witness_limbs(F, [0, 1], steps, 2, sp, engine=eng, out=pin.array)
  boundary = [(0, 0, 0), (0, 1, 1)]
  S = STARK(F, steps, 8, 2, sp, engine=eng)
  proof = S.mk_proof(w, boundary)
  assert S.verify_proof(proof, w, boundary)
  want = oracle.computational_trace(P, [0, 1], steps, sp)
  assert proof == S.mk_proof([[F(v) for v in col] for col in want], boundary)
  pin.free()


def test_small_prime_field(eng, oracle):
  from starks_b200.air import witness_limbs
  from starks_b200.modp import IntegersModP
  F = IntegersModP(31)
  sp = [{(2,): 1, (0,): 3}]
  want = oracle.computational_trace(31, [4], 40, sp)
  got = witness_limbs(F, [4], 40, 1, sp, engine=eng)
  assert oracle.from_limbs(got[0]) == want[0]
  eng.set_field(P)


def test_pointwise_quotients_equal_transform_route(eng, oracle, monkeypatch):
  """Whenever the constraint subgroup is smaller than the domain mk_proof takes D and B pointwise from P's evaluations (stk_quotient_eval,
  stk_boundary_eval) instead of transforming their coefficients; both routes must give the same
  columns and the same proof, and the oracle's proof where it is cheap enough."""
  from starks_b200.air import witness_limbs
  from starks_b200.modp import IntegersModP
  from starks_b200.stark import STARK
  F = IntegersModP(P)
  cases = [(64, 8, [3, 5], [{(0, 1): 1}, {(1, 0): 1, (0, 1): 1}]),
           (256, 16, [1, 2, 3], [{(0, 1, 0): 1}, {(0, 0, 1): 1, (0, 0, 0): 5}, {(1, 0, 0): 7, (0, 1, 0): P - 2, (0, 0, 0): 11}]),
           (1 << 12, 8, [9], [{(1,): 3, (0,): 1}]),
           (128, 8, [2, 3], [{(0, 1): 1}, {(1, 0): 1, (0, 2): 1}]),           # degree 2: M = 2 * steps
           (64, 8, [3], [{(3,): 1, (1,): 2, (0,): 5}])]                        # degree 3: M = 4 * steps
  for steps, ext, inp, sp in cases:
    width = len(inp)
    w = witness_limbs(F, inp, steps, width, sp, engine=eng)
    boundary = [(0, j, inp[j]) for j in range(width)]
    S = STARK(F, steps, ext, width, sp, engine=eng)
    got = {}
    for mode in ("1", "0"):
      monkeypatch.setenv("STK_PROOF_POINTWISE", mode)
      proof = S.mk_proof(w, boundary, keep_device=True)
      cols = S.device["cols"].download((3 * width, steps * ext, 8))
      for b in S.device.values():
        b.free()
      got[mode] = (proof, cols)
    assert (got["1"][1] == got["0"][1]).all(), (steps, ext)
    assert got["1"][0] == got["0"][0]
    assert S.verify_proof(got["1"][0], w, boundary)
    if steps <= 256:
      want = oracle.StarkOracle(steps, ext, width, sp).mk_proof(oracle.computational_trace(P, inp, steps, sp), boundary)
      assert got["1"][0] == want


AFFINE = [
    ("fibonacci", [0, 1], [{(0, 1): 1}, {(1, 0): 1, (0, 1): 1}]),
    ("affine+const", [2, 5], [{(1, 0): 1, (0, 0): 7}, {(1, 0): 1, (0, 1): 3, (0, 0): P - 2}]),
    ("3-wide rotate", [1, 2, 3], [{(0, 1, 0): 1}, {(0, 0, 1): 1}, {(1, 0, 0): 5, (0, 1, 0): 1}]),
]


@pytest.mark.parametrize("name,inp,sp", AFFINE, ids=[c[0] for c in AFFINE])
@pytest.mark.parametrize("steps", [1, 2, 63, 4096, 5000, 1 << 16])
def test_device_trace_affine_chunked(eng, oracle, name, inp, sp, steps):
  """stk_trace_generate_dev on AIRs of degree <= 1: from 4096 steps on, one trace runs as parallel
  chunks whose start states are powers of the companion matrix -- equal to the sequential
  recurrence (starks/air.py:31-52) element for element, ragged last chunk included."""
  from starks_b200.air import witness_device, witness_limbs
  from starks_b200.modp import IntegersModP
  F = IntegersModP(P)
  width = len(inp)
  d_w = witness_device(F, inp, steps, width, sp, engine=eng)
  got = d_w.download((width, steps, 8))
  d_w.free()
  want = witness_limbs(F, inp, steps, width, sp, engine=eng)
  assert (got == want).all()
  if steps <= 4096:
    ref = oracle.computational_trace(P, inp, steps, sp)
    for j in range(width):
      assert oracle.from_limbs(got[j]) == ref[j]


def test_device_traces_batch_and_nonlinear(eng, oracle):
  """Many independent traces of a NON-linear AIR, one thread each; and a single non-linear trace
  through the host recurrence with the upload overlapped (stk_trace_generate_upload)."""
  from starks_b200.air import witness_device, witness_limbs
  from starks_b200.modp import IntegersModP
  F = IntegersModP(P)
  sp = [{(0, 1): 1}, {(1, 0): 1, (0, 3): 1, (0, 0): 42}]
  steps, nt = 300, 37
  inputs = [[i + 1, 2 * i + 5] for i in range(nt)]
  d_w = witness_device(F, inputs, steps, 2, sp, engine=eng, ntraces=nt)
  got = d_w.download((nt, 2, steps, 8))
  d_w.free()
  for t in (0, 1, 17, nt - 1):
    ref = oracle.computational_trace(P, inputs[t], steps, sp)
    for j in range(2):
      assert oracle.from_limbs(got[t, j]) == ref[j], (t, j)
  steps = (1 << 16) + 77            # three upload blocks, the last one ragged
  d_w = witness_device(F, [2, 3], steps, 2, sp, engine=eng)
  eng.sync()
  got = d_w.download((2, steps, 8))
  d_w.free()
  assert (got == witness_limbs(F, [2, 3], steps, 2, sp, engine=eng)).all()


def test_prover_takes_a_device_witness_and_rejects_noncanonical_limbs(eng, oracle):
  from starks_b200.air import witness_device, witness_limbs
  from starks_b200.modp import IntegersModP
  from starks_b200.stark import STARK
  F = IntegersModP(P)
  sp = [{(0, 1): 1}, {(1, 0): 1, (0, 1): 1}]
  steps = 1 << 12
  bnd = [(0, 0, 0), (0, 1, 1)]
  S = STARK(F, steps, 8, 2, sp, engine=eng)
  host = witness_limbs(F, [0, 1], steps, 2, sp, engine=eng)
  d_w = witness_device(F, [0, 1], steps, 2, sp, engine=eng)
  proof = S.mk_proof(d_w, bnd)
  assert proof == S.mk_proof(host, bnd)
  assert S.verify_proof(proof, d_w, bnd)
  d_w.free()
  # limbs >= p are not field elements: the array path must refuse them (the list path reduces)
  bad = host.copy()
  bad[1, 5] = np.array([0xFFFFFFFF] * 8, dtype=np.uint32)
  with pytest.raises(ValueError):
    S.mk_proof(bad, bnd)
  assert eng.count_noncanonical(eng.alloc(bad.nbytes).upload(bad).ptr, 2 * steps) == 1


def test_device_traces_small_prime_field(eng, oracle):
  """The run-time-modulus (Montgomery) instantiation of the device trace kernel: chunked affine
  trace and a batch of non-linear traces over p = 31, against the host recurrence / the oracle."""
  from starks_b200.air import witness_device, witness_limbs
  from starks_b200.modp import IntegersModP
  F = IntegersModP(31)
  try:
    sp = [{(0, 1): 1}, {(1, 0): 3, (0, 1): 1, (0, 0): 7}]
    steps = 5000
    d_w = witness_device(F, [4, 9], steps, 2, sp, engine=eng)
    got = d_w.download((2, steps, 8))
    d_w.free()
    assert (got == witness_limbs(F, [4, 9], steps, 2, sp, engine=eng)).all()
    ref = oracle.computational_trace(31, [4, 9], 64, sp)
    assert oracle.from_limbs(got[0, :64]) == ref[0] and oracle.from_limbs(got[1, :64]) == ref[1]
    sq = [{(2,): 1, (0,): 3}]
    d_w = witness_device(F, [[i] for i in range(20)], 40, 1, sq, engine=eng, ntraces=20)
    got = d_w.download((20, 1, 40, 8))
    d_w.free()
    for t in (0, 7, 19):
      assert oracle.from_limbs(got[t, 0]) == oracle.computational_trace(31, [t], 40, sq)[0]
  finally:
    eng.set_field(P)
