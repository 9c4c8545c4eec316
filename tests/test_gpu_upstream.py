"""The drop-in actually dropping in (north star: "stark.py and the existing tests call it as a
drop-in"): the UNMODIFIED upstream package -- staged under baseline/_ref by
baseline/stage_ref.py, with the one in-memory restoration of SURVEY.md App. B -- runs on the
GPU through starks_b200.install:

  (i)   the upstream unit tests of the hot-path modules (starks/test/test_fft.py,
        test_merkle_tree.py, test_utils.py, test_compression.py) run unmodified after install();
  (ii)  the upstream STARK.mk_proof BODY (starks/stark.py:233-279) runs with only its callees
        rebound (install(replace_prover=False)) and produces the golden proofs; so does the
        whole-prover replacement (install());
  (iii) the upstream, pure-Python STARK.verify_proof (stark.py:281-317, restored FRI verifier,
        hashlib) -- nothing of this repository on its path -- accepts a GPU-made proof of a
        2^12-step trace, and rejects a tampered one.
"""
import os
import sys
import unittest

import pytest

from conftest import ROOT, load_golden

pytestmark = pytest.mark.gpu

P = 2**256 - 351 * 2**32 + 1
sys.path.insert(0, os.path.join(ROOT, "oracle"))


@pytest.fixture(scope="module")
def upstream():
  import pyref
  if not pyref.available():
    pytest.skip("no upstream tree (run baseline/stage_ref.py in the authoring container)")
  with pyref.quiet():
    starks = pyref.load()
    import starks.stark  # noqa: F401  (binds the from-import aliases install() must also rebind)
    import starks.compression  # noqa: F401
  return starks


@pytest.fixture()
def shim():
  import starks_b200.install as s
  yield s
  s.uninstall()


def _proof_digest(proof):
  import oracle as orc
  return orc.proof_digest(proof)


@pytest.mark.parametrize("module", ["test_fft", "test_merkle_tree", "test_utils", "test_compression"])
def test_upstream_unit_tests_run_through_the_shim(upstream, shim, module):
  import pyref
  shim.install()
  import starks.fft
  import starks_b200.fft as bfft
  assert starks.fft.fft_1d is bfft.fft_1d
  path = os.path.join(pyref.REF_ROOT, "starks", "test", module + ".py")
  import importlib.util
  spec = importlib.util.spec_from_file_location("upstream_" + module, path)
  mod = importlib.util.module_from_spec(spec)
  with pyref.quiet():
    spec.loader.exec_module(mod)      # `from starks.x import y` now binds the GPU functions
    suite = unittest.defaultTestLoader.loadTestsFromModule(mod)
    assert suite.countTestCases() > 0
    res = unittest.TestResult()
    suite.run(res)
  assert res.testsRun > 0
  assert not res.errors and not res.failures, (res.errors + res.failures)[0][1]


def _golden(tag):
  for g in load_golden("stark_big.json")["proofs"]:
    if g["tag"] == tag:
      return g
  raise KeyError(tag)


def _upstream_case(starks, steps):
  from starks.modp import IntegersModP
  from starks.polynomial import polynomials_over  # noqa: F401
  from starks.multivariate_polynomial import multivariates_over  # noqa: F401
  from starks.utils import generate_Xi_s
  from starks.air import get_computational_trace
  F = IntegersModP(P)
  Xs = generate_Xi_s(F, 2)
  step_polys = [Xs[1], Xs[0] + Xs[1]]
  trace, _ = get_computational_trace([F(0), F(1)], steps, 2, step_polys)
  witness = [[trace[i][j] for i in range(steps)] for j in range(2)]
  boundary = [(0, 0, F(0)), (0, 1, F(1))]
  return F, step_polys, witness, boundary


@pytest.mark.parametrize("steps,tag,replace", [(32, "fib32", False), (32, "fib32", True), (1024, "fib1024", True),
                                               (1024, "fib1024", False)])
def test_upstream_prover_through_the_shim_matches_golden(upstream, shim, steps, tag, replace):
  import pyref
  g = _golden(tag)
  with pyref.quiet():
    F, step_polys, witness, boundary = _upstream_case(upstream, steps)
    shim.install(replace_prover=replace)
    import starks.stark as us
    S = us.STARK(F, steps, 8, 2, step_polys)
    proof = S.mk_proof(witness, boundary)
  assert proof[0].hex() == g["m_root"] and proof[1].hex() == g["l_root"]
  assert len(proof[2]) == g["n_branches"] and len(proof[3]) == g["fri_layers"]
  assert _proof_digest(proof) == g["digest"]
  with pyref.quiet():
    assert S.verify_proof(proof, witness, boundary)     # upstream verifier, rebound callees
    shim.uninstall()
    S2 = us.STARK(F, steps, 8, 2, step_polys)            # upstream verifier, nothing rebound
    assert S2.verify_proof(proof, witness, boundary)


def test_upstream_verifier_accepts_gpu_proof_2p12(upstream):
  import pyref
  import starks_b200.install as shim
  assert not shim.installed()
  from starks_b200 import Engine
  from starks_b200.modp import IntegersModP as BF
  from starks_b200.stark import STARK as BSTARK
  steps = 1 << 12
  eng = Engine(0)
  sp = [{(0, 1): 1}, {(1, 0): 1, (0, 1): 1}]
  a, b, c0, c1 = 0, 1, [], []
  for _ in range(steps):
    c0.append(a)
    c1.append(b)
    a, b = b, (a + b) % P
  proof = BSTARK(BF(P), steps, 8, 2, sp, engine=eng).mk_proof([c0, c1], [(0, 0, 0), (0, 1, 1)])
  eng.close()
  with pyref.quiet():
    from starks.modp import IntegersModP
    from starks.utils import generate_Xi_s
    import starks.stark as us
    import starks.merkle_tree as umt
    import hashlib
    assert umt.blake(b"abc") == hashlib.blake2s(b"abc").digest()
    assert us.merkelize.__module__ == "starks.merkle_tree"      # really the upstream code path
    F = IntegersModP(P)
    Xs = generate_Xi_s(F, 2)
    S = us.STARK(F, steps, 8, 2, [Xs[1], Xs[0] + Xs[1]])
    witness = [[F(v) for v in c0], [F(v) for v in c1]]
    boundary = [(0, 0, F(0)), (0, 1, F(1))]
    assert S.verify_proof(proof, witness, boundary)
    bad = [proof[0], proof[1], list(proof[2]), proof[3]]
    leaf = bytearray(bad[2][3][0])
    leaf[7] ^= 1
    bad[2][3] = [bytes(leaf)] + list(bad[2][3][1:])
    with pytest.raises(AssertionError):
      S.verify_proof(bad, witness, boundary)
