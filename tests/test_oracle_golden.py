"""Pins the CPU oracle (oracle/) against golden vectors generated from the unmodified
Python reference (oracle/gen_golden.py) and against the reference tests' own known
answers.  CPU only."""
import hashlib

import pytest

from conftest import load_golden

P = 2**256 - 351 * 2**32 + 1


def synth(col, i, p=P):
  d = hashlib.blake2s(col.to_bytes(4, "little") + i.to_bytes(8, "little")).digest()
  return int.from_bytes(d, "big") % p


def H(ints):
  return hashlib.blake2s(b"".join(x.to_bytes(32, "big") for x in ints)).hexdigest()


def test_power_cycle_kat(oracle):
  # starks/test/test_utils.py:20-30 -- the reference's only numeric KAT near the path
  assert oracle.get_power_cycle(31, pow(3, 5, 31)) == [1, 26, 25, 30, 5, 6]
  g = load_golden("field_utils.json")["power_cycle_p31"]
  assert oracle.get_power_cycle(g["p"], g["r"]) == g["cycle"]


def test_field_identities(oracle):
  # starks/test/test_modpy.py:28-35, 54-61
  g = load_golden("field_utils.json")
  assert oracle.fpow(P, 2, 256) == 351 * 2**32 - 1 == int(g["two_pow_256"], 16)
  for e in g["pow7"]:
    assert oracle.fpow(P, 7, int(e["e"], 16)) == int(e["r"], 16) == pow(7, int(e["e"], 16), P)
  a = [int(x["a"], 16) for x in g["field_ops"]]
  b = [int(x["b"], 16) for x in g["field_ops"]]
  assert oracle.field_op(P, "mul", a, b) == [int(x["mul"], 16) for x in g["field_ops"]]
  assert oracle.field_op(P, "add", a, b) == [int(x["add"], 16) for x in g["field_ops"]]
  assert oracle.field_op(P, "sub", a, b) == [int(x["sub"], 16) for x in g["field_ops"]]
  nz = [x for x in g["field_ops"] if x["inv"]]
  assert oracle.field_op(P, "inv", [int(x["a"], 16) for x in nz]) == [int(x["inv"], 16) for x in nz]
  # generic small moduli
  for p in (7, 31, 2**61 - 1, 2**255 - 19):
    xs = [synth(1, i, p) for i in range(16)]
    ys = [synth(2, i, p) for i in range(16)]
    assert oracle.field_op(p, "mul", xs, ys) == [x * y % p for x, y in zip(xs, ys)]
    assert oracle.field_op(p, "sub", xs, ys) == [(x - y) % p for x, y in zip(xs, ys)]


def test_blake2s(oracle):
  g = load_golden("field_utils.json")
  # RFC 7693 appendix B
  assert oracle.blake(b"abc").hex() == "508c5e8c327c14e2e1a72ba34eeb452f37458b209ed63a294d999b4c86675982"
  assert g["blake_abc"] == oracle.blake(b"abc").hex()
  for e in g["blake_lens"]:
    data = bytes((i * 7 + 3) & 255 for i in range(e["len"]))
    assert oracle.blake(data).hex() == e["digest"] == hashlib.blake2s(data).hexdigest()


def test_fiat_shamir(oracle):
  g = load_golden("field_utils.json")
  for e in g["indices"]:
    assert oracle.get_pseudorandom_indices(bytes.fromhex(e["seed"]), e["modulus"], e["count"], e["exclude"]) == e["out"]
  for e in g["ks"]:
    assert [("%064x" % k) for k in oracle.get_pseudorandom_ks(bytes.fromhex(e["root"]), e["num"])] == e["out"]
  assert oracle.get_pseudorandom_ks(b"\0" * 32, 10) is None


def test_fft_small_fields(oracle):
  g = load_golden("fft.json")
  e = g["p31_n6"]
  out = oracle.fft_1d(e["p"], e["in"], e["root"])
  assert out == e["out"] and len(out) == 6
  assert oracle._strip(oracle.fft_1d(e["p"], out, e["root"], inv=True)) == e["inv_of_out"] == [0, 1, 2, 3][:4]
  for c in g["cases"]:
    assert oracle.fft_1d(c["p"], c["in"], c["root"]) == c["out"], c
    assert oracle.fft_1d(c["p"], c["in"], c["root"], inv=True) == c["inv"], c
  with pytest.raises(IndexError):
    oracle.fft_1d(31, [1, 2, 3, 4, 5, 6, 7], pow(3, 5, 31))
  with pytest.raises(IndexError):  # N = 5: the reference's recursion indexes out of range
    oracle.fft_1d(31, [1, 2, 3, 4, 5], pow(3, 6, 31))


def test_fft_stark_prime(oracle):
  g = load_golden("fft.json")
  e = g["stark_n8"]
  assert [("%064x" % x) for x in oracle.fft_1d(P, e["in"], int(e["root"], 16))] == e["out"]
  for s in g["synth"]:
    n = 1 << s["logn"]
    w = pow(7, (P - 1) // n, P)
    assert "%064x" % w == s["w"]
    vals = [synth(0, i) for i in range(n)]
    ev = oracle.fft_1d(P, vals, w, order=n)
    assert H(ev) == s["H_ev"] and "%064x" % ev[1] == s["ev1"]
    assert [("%064x" % x) for x in ev[:4]] == s["first"] and "%064x" % ev[-1] == s["last"]
    iv = oracle.fft_1d(P, vals, w, inv=True, order=n)
    assert H(iv) == s["H_inv"]
    assert oracle.merkelize(ev)[1].hex() == s["root"]
    if n <= 64:  # definition check: out[k] = sum_j in[j] w^(jk)
      assert ev == [sum(v * pow(w, j * k, P) for j, v in enumerate(vals)) % P for k in range(n)]
  e = g["padded"]
  w = pow(7, (P - 1) // e["n"], P)
  assert H(oracle.fft_1d(P, [synth(e["col"], i) for i in range(e["n_in"])], w)) == e["H_ev"]
  e = g["mul_polys_512"]
  prod = oracle.mul_polys(P, e["a"], e["a"], int(e["root"], 16))
  assert H(prod) == e["H"] and [("%064x" % x) for x in prod[:8]] == e["first"]


def test_survey_appendix_b_kats(oracle):
  # SURVEY.md App. B table (generated during the survey from the unmodified reference)
  kat = {3: ("26aa9a062f137e335cdb0604c2bab9129760c0672d6658951ef1770bf7ad398e",
             "a6b6fba700d1a51349925119ab07a565ab22b192db4c0b036d8a121cae56f845"),
         10: ("b3dfdea464ff8bd51884f283ed1c8f6229cfefe2944590fe590c0dfbd26d3f36",
              "6b163761b5ca1175a76d5ada3af2d2d343d92330738a09af49b21b46adf011d8")}
  for logn, (h, root) in kat.items():
    n = 1 << logn
    w = pow(7, (P - 1) // n, P)
    ev = oracle.fft_1d(P, [synth(0, i) for i in range(n)], w, order=n)
    assert H(ev) == h and oracle.merkelize(ev)[1].hex() == root


def _leaves(entry):
  n, ll = entry["n"], entry["leaf_len"]
  if entry["leaves_rule"] == "range":
    return [x.to_bytes(32, "big") for x in range(n)]
  return [bytes(hashlib.blake2s(b"%d-%d-%d" % (n, i, k)).digest()[0] for k in range(ll)) for i in range(n)]


def test_merkle(oracle):
  g = load_golden("merkle.json")
  assert [0, 2, 4, 6, 1, 3, 5, 7] == g["permute4_8"]
  for x, L, y in g["index_in_permuted"]:
    assert oracle.get_index_in_permuted(x, L) == y
  for e in g["trees"]:
    L = _leaves(e)
    t = oracle.merkelize(L)
    assert len(t) == e["tree_len"], e["tag"]
    if e["root"] is not None:
      assert t[1].hex() == e["root"], e["tag"]
    assert hashlib.blake2s(b"".join(t)).hexdigest() == e["tree_digest"], e["tag"]
    for b in e.get("branches", []):
      br = oracle.mk_branch(t, b["index"])
      assert [x.hex() for x in br] == b["branch"]
      if e["n"] & (e["n"] - 1) == 0:
        assert oracle.verify_branch(t[1], b["index"], br) == L[b["index"]]
  # starks/test/test_merkle_tree.py:16-22
  t = oracle.merkelize([x.to_bytes(32, "big") for x in range(128)])
  assert oracle.verify_branch(t[1], 59, oracle.mk_branch(t, 59), output_as_int=True) == 59
  assert len(oracle.mk_branch(t, 59)) == 8
  # multithreaded level-order variant produces the same nodes
  import numpy as np
  leaves = np.frombuffer(b"".join(_leaves(g["trees"][1])), dtype=np.uint8).reshape(144, 32)
  p1, n1 = oracle.merkelize_bytes(leaves, 1)
  p4, n4 = oracle.merkelize_bytes(leaves, 4)
  assert (p1 == p4).all() and (n1 == n4).all()


def test_lde_commit(oracle):
  e = load_golden("merkle.json")["lde_commit"]
  steps, ext = e["steps"], e["ext"]
  N = steps * ext
  G2 = pow(7, (P - 1) // N, P)
  G1 = pow(G2, ext, P)
  evs = []
  for c in range(e["cols"]):
    coeffs = oracle._strip(oracle.fft_1d(P, [synth(c, i) for i in range(steps)], G1, inv=True, order=steps))
    evs.append(oracle.fft_1d(P, coeffs, G2, order=N))
    assert H(evs[-1]) == e["H_cols"][c]
    assert evs[-1][::ext] == [synth(c, i) for i in range(steps)]  # no coset shift
  mt = oracle.merkelize_polynomial_evaluations(evs)
  assert mt[1].hex() == e["root"] == "a4e01ec6f96c6f2c34a92bf4998512919b0385ce83c57d6445c0f608ae155cf5"
  assert [b.hex() for b in oracle.mk_branch(mt, 5)] == e["branch5"] and len(e["branch5"]) == 12
  assert hashlib.blake2s(b"".join(mt)).hexdigest() == e["tree_digest"]


def test_fri(oracle):
  g = load_golden("fri.json")
  e = g["fold64"]
  vals = [synth(e["col"], i) for i in range(e["n"])]
  col = oracle.fri_fold(P, int(e["root"], 16), vals, int(e["special_x"], 16))
  assert [("%064x" % c) for c in col] == e["column"]
  for pr in g["proofs"]:
    n = 1 << pr["logn"]
    w = pow(7, (P - 1) // n, P)
    f = [synth(pr["col"], i) for i in range(pr["deg"])]
    prf = oracle.fri_prove(P, f, w, pr["deg"], exclude_multiples_of=pr["exclude"])
    assert len(prf) == pr["layers"]
    assert [layer[0].hex() for layer in prf[:-1]] == pr["roots"]
    assert oracle.proof_digest(prf) == pr["digest"]


def _step_polys(desc):
  return [{tuple(int(t) for t in k.split(",")): v for k, v in sp.items()} for sp in desc]


@pytest.mark.parametrize("tag", ["fib8", "fib32", "cubic8", "affine32", "w3_8", "w6_8", "quad128", "fib256",
                                 "deg4_8", "deg5_16", "deg7_8", "deg8_8", "deg8_32", "w4_mixed_16"])
def test_stark_proofs(oracle, tag):
  g = {e["tag"]: e for name in ("stark.json", "stark_degrees.json") for e in load_golden(name)["proofs"]}
  e = g[tag]
  sp = _step_polys(e["step_polys"])
  S = oracle.StarkOracle(e["steps"], e["ext"], e["width"], sp)
  witness = oracle.computational_trace(P, e["inp"], e["steps"], sp)
  assert [("%064x" % w[-1]) for w in witness] == e["output"]
  boundary = [(0, j, e["inp"][j]) for j in range(e["width"])]
  proof = S.mk_proof(witness, boundary)
  assert proof[0].hex() == e["m_root"]
  assert proof[1].hex() == e["l_root"]
  assert len(proof[2]) == e["n_branches"] and len(proof[3]) == e["fri_layers"]
  assert [b.hex() for b in proof[2][0]] == e["branch0"]
  assert oracle.proof_digest(proof) == e["digest"]
