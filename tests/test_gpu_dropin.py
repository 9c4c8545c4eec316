"""GPU parity of the reference-facing Python modules (starks_b200.fft / merkle_tree / fri /
stark), written after the reference's own tests (starks/test/test_fft.py, test_merkle_tree.py,
test_utils.py, the commented test_fri.py / test_stark.py) and checked against the golden
proofs generated from the reference and against the CPU oracle.  Bit-exact."""
import pytest

from conftest import load_golden

pytestmark = pytest.mark.gpu

P = 2**256 - 351 * 2**32 + 1


@pytest.fixture(scope="module")
def mods():
  import starks_b200.fft as fft
  import starks_b200.merkle_tree as mt
  import starks_b200.fri as fri
  import starks_b200.stark as stark
  import starks_b200.utils as utils
  from starks_b200.modp import IntegersModP
  from starks_b200.polynomial import polynomials_over, generate_Xi_s
  return dict(fft=fft, mt=mt, fri=fri, stark=stark, utils=utils, IntegersModP=IntegersModP,
              polynomials_over=polynomials_over, generate_Xi_s=generate_Xi_s)


# ---- starks/test/test_fft.py ---------------------------------------------------------
def test_fft_basic_inv_type(mods):
  F = mods["IntegersModP"](31)
  polysOver = mods["polynomials_over"](F).factory
  poly = polysOver([val for val in range(4)])
  root = F(3)**((31 - 1) // 6)
  solver = mods["fft"].NonBinaryFFT(F, root)
  ev = solver.fft(poly)
  assert len(ev) == 6                                   # test_basic (:98-113)
  assert all(isinstance(v, F) for v in ev)              # test_fft_output_type (:151-168)
  assert solver.inv_fft(ev) == poly                     # test_fft_inv (:132-149)
  g = load_golden("fft.json")["p31_n6"]
  assert [int(v) for v in ev] == g["out"]


def test_fft_large_modulus_and_mul_polys(mods):
  F = mods["IntegersModP"](P)
  poly = mods["polynomials_over"](F).factory([val for val in range(4)])
  root = F(7)**((P - 1) // 8)
  ev = mods["fft"].NonBinaryFFT(F, root).fft(poly)
  assert len(ev) == 8                                   # test_large_modulus (:115-130)
  g = load_golden("fft.json")
  assert [("%064x" % int(v)) for v in ev] == g["stark_n8"]["out"]
  r512 = F(7)**((P - 1) // 512)
  a = [F(v) for v in range(4)]
  prod = mods["fft"].mul_polys(a, a, r512)              # test_mul_polys (:185-194)
  assert [("%064x" % int(v)) for v in prod[:8]] == g["mul_polys_512"]["first"] and len(prod) == 512
  with pytest.raises(IndexError):
    mods["fft"].fft_1d(F, [1] * 9, P, root)


def test_utils_power_cycle(mods):
  F = mods["IntegersModP"](31)                          # test_utils.py:20-30
  assert [int(x) for x in mods["utils"].get_power_cycle(F(3)**5, F)] == [1, 26, 25, 30, 5, 6]
  g = load_golden("field_utils.json")
  for e in g["indices"]:
    assert mods["utils"].get_pseudorandom_indices(bytes.fromhex(e["seed"]), e["modulus"], e["count"], e["exclude"]) == e["out"]
  for e in g["ks"]:
    assert [("%064x" % k) for k in mods["stark"].get_pseudorandom_ks(bytes.fromhex(e["root"]), e["num"])] == e["out"]


# ---- starks/test/test_merkle_tree.py -------------------------------------------------
def test_merkle_module(mods, oracle):
  mt = mods["mt"]
  t = mt.merkelize([x.to_bytes(32, "big") for x in range(128)])
  b = mt.mk_branch(t, 59)
  assert mt.verify_branch(t[1], 59, b, output_as_int=True) == 59      # :16-22
  assert len(t) == 256 and len(b) == 8                                # :41-51
  F7 = mods["IntegersModP"](7)
  t = mt.merkelize([F7(i) for i in range(144)])                       # :32-38, n = 144
  assert len(t) == 288
  assert t == oracle.merkelize([i % 7 for i in range(144)])
  t = mt.merkelize(list(range(1000, 1064)))                           # ints
  assert t == oracle.merkelize(list(range(1000, 1064)))
  assert mt.merkelize([1, 2, 3]) == []
  # merkelize_polynomial_evaluations + unpack (:62-79)
  F = mods["IntegersModP"](P)
  cols = [[F(oracle.synth(c, i)) for i in range(64)] for c in range(3)]
  mtree = mt.merkelize_polynomial_evaluations(1, cols)
  assert mtree == oracle.merkelize_polynomial_evaluations([[int(v) for v in c] for c in cols])
  leaf = mt.verify_branch(mtree[1], 5, mt.mk_branch(mtree, 5))
  assert mt.unpack_merkle_leaf(leaf, 1, 3) == [cols[c][5].to_bytes() for c in range(3)]


# ---- FRI (commented starks/test/test_fri.py:34-52, 105-258) ---------------------------
def test_fri_proofs_golden(mods, oracle):
  F = mods["IntegersModP"](P)
  fri = mods["fri"].SmoothSubgroupFRI(F)
  for pr in load_golden("fri.json")["proofs"]:
    n = 1 << pr["logn"]
    w = F(7)**((P - 1) // n)
    f = mods["polynomials_over"](F).factory([oracle.synth(pr["col"], i) for i in range(pr["deg"])])
    proof = fri.generate_proximity_proof(f, w, pr["deg"], exclude_multiples_of=pr["exclude"])
    assert len(proof) == pr["layers"]
    assert [layer[0].hex() for layer in proof[:-1]] == pr["roots"]
    assert oracle.proof_digest(proof) == pr["digest"]
    assert fri.verify_proximity_proof(proof, bytes.fromhex(pr["merkle_root"]), w, pr["deg"],
                                      exclude_multiples_of=pr["exclude"])
    for layer in proof[:-1]:
      assert len(layer[1]) == 40 and all(len(b) == 5 for b in layer[1])
    # a corrupted proof is rejected
    bad = [list(x) if isinstance(x, list) else x for x in proof]
    bad[-1] = list(bad[-1])
    bad[-1][3] = (int.from_bytes(bad[-1][3], "big") ^ 1).to_bytes(32, "big")
    with pytest.raises(AssertionError):
      fri.verify_proximity_proof(bad, bytes.fromhex(pr["merkle_root"]), w, pr["deg"], exclude_multiples_of=pr["exclude"])


def test_fri_vs_oracle_larger(mods, oracle):
  F = mods["IntegersModP"](P)
  n, deg = 1 << 14, 1 << 11
  w = pow(7, (P - 1) // n, P)
  f = [oracle.synth(4, i) for i in range(deg)]
  got = mods["fri"].FRI(F).generate_proximity_proof(f, F(w), deg, exclude_multiples_of=8)
  want = oracle.fri_prove(P, f, w, deg, exclude_multiples_of=8)
  assert got == want and len(got) == 5


# ---- STARK (commented starks/test/test_stark.py:215-350) -------------------------------
def _golden_proofs():
  out = {}
  for name in ("stark.json", "stark_big.json", "stark_degrees.json"):
    try:
      for e in load_golden(name)["proofs"]:
        out[e["tag"]] = e
    except FileNotFoundError:
      pass
  return out


def _step_polys(mods, F, e):
  Xs = mods["generate_Xi_s"](F, e["width"])
  polys = []
  for sp in e["step_polys"]:
    acc = None
    for k, v in sp.items():
      term = F(v)
      exps = [int(t) for t in k.split(",")]
      mono = None
      for X, ex in zip(Xs, exps):
        if ex:
          mono = X**ex if mono is None else mono * X**ex
      term = term * mono if mono is not None else (Xs[0] * 0 + term)
      acc = term if acc is None else acc + term
    polys.append(acc)
  return polys


@pytest.mark.parametrize("tag", ["fib8", "fib32", "cubic8", "affine32", "w3_8", "w6_8", "quad128", "fib256", "fib1024",
                                 "deg4_8", "deg5_16", "deg7_8", "deg8_8", "deg8_32", "w4_mixed_16"])
def test_stark_proofs_golden(mods, oracle, tag):
  g = _golden_proofs()
  if tag not in g:
    pytest.skip("golden proof %s not generated" % tag)
  e = g[tag]
  F = mods["IntegersModP"](P)
  step_polys = _step_polys(mods, F, e)
  # get_computational_trace (starks/air.py:31-52) + generate_witness (:124)
  trace = [[F(v) for v in e["inp"]]]
  for _ in range(e["steps"] - 1):
    trace.append([sp(trace[-1]) for sp in step_polys])
  witness = [[trace[i][j] for i in range(e["steps"])] for j in range(e["width"])]
  assert [("%064x" % int(wc[-1])) for wc in witness] == e["output"]
  boundary = [(0, j, F(e["inp"][j])) for j in range(e["width"])]
  S = mods["stark"].STARK(F, e["steps"], e["ext"], e["width"], step_polys)
  proof = S.mk_proof(witness, boundary)
  assert isinstance(proof, list) and len(proof) == 4
  assert proof[0].hex() == e["m_root"]
  assert proof[1].hex() == e["l_root"]
  assert len(proof[2]) == e["n_branches"] and len(proof[3]) == e["fri_layers"]
  assert [b.hex() for b in proof[2][0]] == e["branch0"]
  assert oracle.proof_digest(proof) == e["digest"]
  assert S.verify_proof(proof, witness, boundary)


def test_stark_rejects_bad_witness(mods):
  F = mods["IntegersModP"](P)
  X = mods["generate_Xi_s"](F, 2)
  step_polys = [X[1], X[0] + X[1]]
  steps = 32
  trace = [[F(0), F(1)]]
  for _ in range(steps - 1):
    trace.append([sp(trace[-1]) for sp in step_polys])
  witness = [[trace[i][j] for i in range(steps)] for j in range(2)]
  witness[1][7] = witness[1][7] + 1
  S = mods["stark"].STARK(F, steps, 8, 2, step_polys)
  with pytest.raises(AssertionError):   # the reference asserts `cp % z == 0` (stark.py:74-75)
    S.mk_proof(witness, [(0, 0, F(0)), (0, 1, F(1))])


@pytest.mark.parametrize("logsteps", [10, 11])
def test_stark_vs_oracle_intermediates(mods, oracle, logsteps):
  """Sizes beyond the golden files: the C oracle restates mk_proof with the reference's
  O(n^2) coefficient-form division; proofs must be equal object for object."""
  steps = 1 << logsteps
  F = mods["IntegersModP"](P)
  sp = [{(1, 0): 1, (0, 2): 1}, {(1, 1): 3, (0, 0): 5}]  # X1 + X2^2 ; 3*X1*X2 + 5
  witness = oracle.computational_trace(P, [2, 3], steps, sp)
  boundary = [(0, 0, 2), (0, 1, 3)]
  want = oracle.StarkOracle(steps, 8, 2, sp).mk_proof(witness, boundary)
  S = mods["stark"].STARK(F, steps, 8, 2, sp)
  got = S.mk_proof(witness, boundary)
  assert got[0] == want[0] and got[1] == want[1]
  assert got == want
  assert S.verify_proof(got, witness, boundary)
