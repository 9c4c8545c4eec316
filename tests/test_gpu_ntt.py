"""GPU parity: stk_ntt / stk_ntt_host / stk_mul_polys / stk_power_cycle vs the CPU oracle
(fft_1d restatement) and the committed golden vectors.  Bit-exact (integer work)."""
import hashlib

import numpy as np
import pytest

from conftest import load_golden

pytestmark = pytest.mark.gpu

P = 2**256 - 351 * 2**32 + 1


def H(ints):
  return hashlib.blake2s(b"".join(x.to_bytes(32, "big") for x in ints)).hexdigest()


@pytest.fixture(scope="module")
def eng():
  from starks_b200 import Engine
  e = Engine(0)
  yield e
  e.close()


def rand_cols(rng, batch, n, p=P):
  # uniform-ish residues below p
  a = rng.integers(0, 2**32, size=(batch, n, 8), dtype=np.uint64).astype(np.uint32)
  if p == P:
    a[:, :, 7] &= 0x7FFFFFFF
  return a


@pytest.mark.parametrize("logn", list(range(3, 17)))
def test_ntt_matches_oracle(eng, oracle, logn):
  n = 1 << logn
  batch = 5 if logn <= 12 else 2
  w = pow(7, (P - 1) // n, P)
  rng = np.random.default_rng(logn)
  cols = rand_cols(rng, batch, n)
  # edge values in the first column
  cols[0, 0] = oracle.to_limbs([P - 1])[0]
  cols[0, 1] = 0
  cols[0, 2] = oracle.to_limbs([1])[0]
  eng.set_field(P)
  got = eng.ntt_host(cols, n, w)
  want = oracle.fft_limbs(P, w, cols, n, nthreads=4)
  assert (got == want).all()
  goti = eng.ntt_host(cols, n, w, inverse=True)
  wanti = oracle.fft_limbs(P, w, cols, n, inv=True, nthreads=4)
  assert (goti == wanti).all()


def test_golden_synth(eng, oracle):
  g = load_golden("fft.json")
  eng.set_field(P)
  for s in g["synth"]:
    n = 1 << s["logn"]
    if n < 8:
      continue
    w = int(s["w"], 16)
    cols = oracle.to_limbs([oracle.synth(0, i) for i in range(n)]).reshape(1, n, 8)
    ev = oracle.from_limbs(eng.ntt_host(cols, n, w)[0])
    assert H(ev) == s["H_ev"] and "%064x" % ev[1] == s["ev1"]
    iv = oracle.from_limbs(eng.ntt_host(cols, n, w, inverse=True)[0])
    assert H(iv) == s["H_inv"]


def test_zero_padding_and_index_error(eng, oracle):
  g = load_golden("fft.json")["padded"]
  n = g["n"]
  w = pow(7, (P - 1) // n, P)
  eng.set_field(P)
  cols = oracle.to_limbs([oracle.synth(g["col"], i) for i in range(g["n_in"])]).reshape(1, -1, 8)
  assert H(oracle.from_limbs(eng.ntt_host(cols, n, w)[0])) == g["H_ev"]
  with pytest.raises(IndexError):
    eng.ntt_host(np.zeros((1, n + 1, 8), np.uint32), n, w)
  # empty input -> all zeros
  assert not eng.ntt_host(np.zeros((2, 0, 8), np.uint32), n, w).any()


def test_small_and_generic_moduli(eng, oracle):
  g = load_golden("fft.json")
  for c in g["cases"] + [dict(p=31, root=g["p31_n6"]["root"], n=6, **{"in": g["p31_n6"]["in"]},
                               out=g["p31_n6"]["out"], inv=None)]:
    eng.set_field(c["p"])
    cols = oracle.to_limbs(c["in"]).reshape(1, -1, 8)
    got = oracle.from_limbs(eng.ntt_host(cols, c["n"], c["root"])[0])
    assert got == c["out"], c
    if c["inv"] is not None:
      assert oracle.from_limbs(eng.ntt_host(cols, c["n"], c["root"], inverse=True)[0]) == c["inv"]
  # power-of-two transform over generic (Montgomery) moduli, incl. STARK-prime N=8 golden
  for p, gen in ((2**64 - 2**32 + 1, 7), (3 * 2**30 + 1, 5), (2**255 - 19, None)):
    if gen is None:
      continue
    eng.set_field(p)
    for logn in (3, 6, 10, 13):
      n = 1 << logn
      w = pow(gen, (p - 1) // n, p)
      assert pow(w, n // 2, p) == p - 1
      rng = np.random.default_rng(logn)
      vals = [int(x) % p for x in rng.integers(0, 2**62, size=n)]
      cols = oracle.to_limbs(vals).reshape(1, n, 8)
      got = eng.ntt_host(cols, n, w)
      want = oracle.fft_limbs(p, w, cols, n)
      assert (got == want).all(), (p, logn)
      assert (eng.ntt_host(got, n, w, inverse=True) == cols).all()
  eng.set_field(P)
  e = g["stark_n8"]
  cols = oracle.to_limbs(e["in"]).reshape(1, -1, 8)
  assert [("%064x" % x) for x in oracle.from_limbs(eng.ntt_host(cols, 8, int(e["root"], 16))[0])] == e["out"]


def test_mul_polys_and_power_cycle(eng, oracle):
  g = load_golden("fft.json")["mul_polys_512"]
  eng.set_field(P)
  a = oracle.to_limbs(g["a"])
  prod = oracle.from_limbs(eng.mul_polys(a, a, 512, int(g["root"], 16)))
  assert H(prod) == g["H"] and [("%064x" % x) for x in prod[:8]] == g["first"]
  eng.set_field(31)
  assert oracle.from_limbs(eng.power_cycle(pow(3, 5, 31), 6)) == [1, 26, 25, 30, 5, 6]
  eng.set_field(P)
  w = pow(7, (P - 1) // 4096, P)
  assert oracle.from_limbs(eng.power_cycle(w, 4096)) == oracle.get_power_cycle(P, w)


@pytest.mark.parametrize("logn,batch", [(18, 3), (20, 2), (21, 1), (22, 1), (24, 1)])
def test_large_sizes_by_properties(eng, oracle, logn, batch):
  """Sizes the oracle cannot finish in seconds: inverse(forward(x)) == x, out[k] equals a
  direct Horner evaluation at w^k for a few k (plain Python ints), and linearity."""
  n = 1 << logn
  w = pow(7, (P - 1) // n, P)
  rng = np.random.default_rng(logn)
  eng.set_field(P)
  cols = rand_cols(rng, batch, n)
  ev = eng.ntt_host(cols, n, w)
  back = eng.ntt_host(ev, n, w, inverse=True)
  assert (back == cols).all()
  # sparse polynomial check: a column with few non-zero coefficients evaluates cheaply
  sp = np.zeros((1, n, 8), np.uint32)
  idx = [0, 1, 5, n // 3, n // 2 + 1, n - 1]
  coef = [int(x) for x in rng.integers(1, 2**62, size=len(idx))]
  for i, cf in zip(idx, coef):
    sp[0, i] = oracle.to_limbs([cf])[0]
  evs = eng.ntt_host(sp, n, w)[0]
  for k in [0, 1, 2, 3, n // 2, n // 2 + 1, n - 1, 12345 % n, (n // 7) | 1]:
    x = pow(w, k, P)
    want = sum(cf * pow(x, i, P) for i, cf in zip(idx, coef)) % P
    assert oracle.from_limbs(evs[k:k + 1])[0] == want, k


def test_zero_padded_routes_agree(eng, oracle, monkeypatch):
  """A zero-padded forward transform (n_in <= N/8) has two routes: the expansion round of the long
  transform and eight coset transforms over <w^8> (ntt.cuh, cshift).  Both must give the same
  evaluations, and the oracle's where it is cheap."""
  import numpy as np
  rng = np.random.default_rng(9)
  for logn, cols, n_in in ((6, 3, 8), (9, 5, 33), (12, 4, 512), (15, 2, 4096), (16, 70, 8192), (21, 2, 1 << 18)):
    n = 1 << logn
    w = pow(7, (P - 1) // n, P)
    a = rng.integers(0, 2**32, size=(cols, n_in, 8), dtype=np.uint64).astype(np.uint32)
    a[:, :, 7] &= 0x7FFFFFFF
    d_in = eng.alloc(a.nbytes).upload(a)
    d_out = eng.alloc(cols * n * 32)
    got = {}
    for mode in ("0", "2"):
      monkeypatch.setenv("STK_LDE_COSET", mode)
      eng._check(eng.lib.stk_memset(eng.ctx, d_out.ptr, 0x5A, cols * n * 32))
      eng.ntt(d_in.ptr, n_in, n_in, d_out.ptr, n, n, cols, w)
      got[mode] = d_out.download((cols, n, 8))
    assert (got["0"] == got["2"]).all(), (logn, cols, n_in)
    if logn <= 12:
      assert (got["2"] == oracle.fft_limbs(P, w, a, n)).all()
    d_in.free(); d_out.free()
