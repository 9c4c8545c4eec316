"""Generates tests/golden/*.json by running the UNMODIFIED upstream Python reference
(/root/reference, with the App. B FRI restoration applied in memory by pyref.py).

Run in the authoring container only:   python oracle/gen_golden.py [--big | --degrees]
The reference tree does not travel to the GPU box; the JSON fixtures do.
Inputs use the synthetic generator of SURVEY.md 8(d):
  synth(col, i) = int(blake2s(le32(col) || le64(i))) mod p ;  w_N = 7^((p-1)/N).
`--big` adds the 1024-step Fibonacci proof (about a minute of reference time).
"""
import hashlib
import json
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import pyref  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
P = 2**256 - 351 * 2**32 + 1


def synth(col, i, p=P):
  d = hashlib.blake2s(col.to_bytes(4, "little") + i.to_bytes(8, "little")).digest()
  return int.from_bytes(d, "big") % p


def hx(x):
  return "%064x" % x


def H(ints):
  return hashlib.blake2s(b"".join(x.to_bytes(32, "big") for x in ints)).hexdigest()


def proof_digest(proof):
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


def shape(x):
  if isinstance(x, (bytes, bytearray)):
    return len(x)
  return [shape(y) for y in x]


def main():
  big = "--big" in sys.argv
  degrees = "--degrees" in sys.argv   # only tests/golden/stark_degrees.json (everything else untouched)
  import tempfile
  OUTX = tempfile.mkdtemp() if degrees else OUT
  st = pyref.load()
  from starks.modp import IntegersModP
  from starks.polynomial import polynomials_over
  from starks.fft import NonBinaryFFT, fft_1d, mul_polys
  from starks.merkle_tree import (merkelize, mk_branch, verify_branch, merkelize_polynomial_evaluations,
                                  permute4, get_index_in_permuted, blake)
  from starks.utils import get_power_cycle, get_pseudorandom_indices, generate_Xi_s
  from starks.poly_utils import multi_interp_4
  from starks.air import get_computational_trace
  import starks.stark as stark_mod
  from starks.fri import SmoothSubgroupFRI

  os.makedirs(OUT, exist_ok=True)
  F = IntegersModP(P)
  F31 = IntegersModP(31)

  # ------------------------------------------------------------------ field / utils
  g = {}
  g["power_cycle_p31"] = {"p": 31, "r": int(F31(3)**5), "cycle": [int(x) for x in get_power_cycle(F31(3)**5, F31)]}
  g["two_pow_256"] = hx(int(F(2)**256))
  g["pow7"] = [{"e": hx(e), "r": hx(int(F(7)**e))} for e in (0, 1, 2, 65537, (P - 1) // 2, (P - 1) // 2**20, P - 2)]
  muls = []
  for k in range(24):
    a, b = synth(100, k), synth(101, k)
    if k == 0:
      a, b = P - 1, P - 1
    if k == 1:
      a, b = 0, synth(101, 1)
    muls.append({"a": hx(a), "b": hx(b), "mul": hx(int(F(a) * F(b))), "add": hx(int(F(a) + F(b))),
                 "sub": hx(int(F(a) - F(b))), "inv": hx(int(F(a).inverse())) if a else None})
  g["field_ops"] = muls
  g["indices"] = []
  for (seed, mod, cnt, ex) in [(b"\x01" * 32, 8192, 80, 8), (b"\x02" * 32, 512, 40, 8), (b"\xfe" * 32, 2**23, 80, 8),
                               (b"\x03" * 32, 100, 40, 0), (b"abc", 64, 3, 4)]:
    g["indices"].append({"seed": seed.hex(), "modulus": mod, "count": cnt, "exclude": ex,
                         "out": get_pseudorandom_indices(seed, mod, cnt, exclude_multiples_of=ex)})
  g["ks"] = []
  for num in (0, 2, 4, 5, 6, 9):
    ks = stark_mod.get_pseudorandom_ks(b"\x07" * 32, num)
    g["ks"].append({"root": (b"\x07" * 32).hex(), "num": num, "out": [hx(k) for k in ks]})
  g["blake_abc"] = blake(b"abc").hex()
  g["blake_lens"] = [{"len": n, "digest": blake(bytes((i * 7 + 3) & 255 for i in range(n))).hex()}
                     for n in (0, 1, 32, 63, 64, 65, 128, 192, 384, 2048, 4096)]
  json.dump(g, open(os.path.join(OUTX, "field_utils.json"), "w"), indent=1)

  # ---------------------------------------------------------------------------- fft
  g = {"cases": []}
  polys31 = polynomials_over(F31).factory
  r6 = F31(3)**5
  ev = NonBinaryFFT(F31, r6).fft(polys31([0, 1, 2, 3]))
  g["p31_n6"] = {"p": 31, "root": int(r6), "in": [0, 1, 2, 3], "out": [int(x) for x in ev],
                 "inv_of_out": [int(x) for x in NonBinaryFFT(F31, r6).inv_fft(ev).coefficients]}
  for (pp, gen, n, vals) in [(31, 3, 2, [5, 9]), (31, 3, 3, [1, 2, 3]), (31, 3, 6, [7, 0, 30, 4, 11, 2]),
                             (7, 3, 6, [1, 2, 3, 4, 5, 6]), (7, 3, 3, [6, 6, 1]), (31, 3, 1, [9])]:
    Fp = IntegersModP(pp)
    r = Fp(gen)**((pp - 1) // n)
    out = fft_1d(Fp, [Fp(v) for v in vals], pp, r)
    inv = fft_1d(Fp, [Fp(v) for v in vals], pp, r, inv=True)
    g["cases"].append({"p": pp, "root": int(r), "n": n, "in": vals, "out": [int(x) for x in out],
                       "inv": [int(x) for x in inv]})
  # STARK prime, N=8 with 4 coefficients (test_large_modulus)
  r8 = F(7)**((P - 1) // 8)
  ev = NonBinaryFFT(F, r8).fft(polynomials_over(F).factory([0, 1, 2, 3]))
  g["stark_n8"] = {"root": hx(int(r8)), "in": [0, 1, 2, 3], "out": [hx(int(x)) for x in ev]}
  # synthetic sweeps
  g["synth"] = []
  for logn in (2, 3, 4, 5, 6, 7, 8, 10, 11, 12) + ((14, 16) if big else ()):
    n = 1 << logn
    w = F(7)**((P - 1) // n)
    vals = [synth(0, i) for i in range(n)]
    t0 = time.time()
    ev = [int(x) for x in fft_1d(F, vals, P, w)]
    iv = [int(x) for x in fft_1d(F, vals, P, w, inv=True)]
    t = merkelize(ev)
    g["synth"].append({"logn": logn, "w": hx(int(w)), "ev1": hx(ev[1]), "H_ev": H(ev), "H_inv": H(iv),
                       "root": t[1].hex(), "first": [hx(x) for x in ev[:4]], "last": hx(ev[-1])})
    print("fft 2^%d: %.2fs" % (logn, time.time() - t0))
  # short input is zero padded
  n = 64
  w = F(7)**((P - 1) // n)
  vals = [synth(3, i) for i in range(23)]
  g["padded"] = {"n": n, "n_in": 23, "col": 3, "H_ev": H([int(x) for x in fft_1d(F, vals, P, w)])}
  # mul_polys (test_mul_polys)
  r512 = F(7)**((P - 1) // 512)
  a = [F(v) for v in range(4)]
  prod = mul_polys(a, a, r512)
  g["mul_polys_512"] = {"root": hx(int(r512)), "a": [0, 1, 2, 3], "H": H([int(x) for x in prod]),
                        "first": [hx(int(x)) for x in prod[:8]]}
  json.dump(g, open(os.path.join(OUTX, "fft.json"), "w"), indent=1)

  # ------------------------------------------------------------------------- merkle
  g = {"trees": []}
  for (n, ll, tag) in [(128, 32, "test_merkle_tree"), (144, 32, "n144"), (8, 32, "n8"), (4, 32, "n4"),
                       (16, 64, "wide64"), (64, 192, "w6cols"), (6, 32, "n6_truncates"), (32, 2048, "w64cols"),
                       (2, 32, "n2_empty")]:
    if tag == "test_merkle_tree":
      L = [x.to_bytes(32, "big") for x in range(128)]
    else:
      L = [bytes(hashlib.blake2s(b"%d-%d-%d" % (n, i, k)).digest()[0] for k in range(ll)) for i in range(n)]
    t = merkelize(L)
    entry = {"tag": tag, "n": n, "leaf_len": ll, "leaves_rule": "range" if tag == "test_merkle_tree" else "blake_bytes",
             "tree_len": len(t), "root": t[1].hex() if len(t) > 1 else None,
             "tree_digest": hashlib.blake2s(b"".join(t)).hexdigest()}
    if len(t) >= 8:
      idxs = [0, 1, (len(t) // 2) - 1, min(59, len(t) // 2 - 1)]
      entry["branches"] = [{"index": i, "branch": [b.hex() for b in mk_branch(t, i)]} for i in idxs]
      if n & (n - 1) == 0:  # verify_branch assumes a power-of-two leaf count
        for i in idxs:
          assert verify_branch(t[1], i, mk_branch(t, i)) == L[i]
    g["trees"].append(entry)
  g["permute4_8"] = permute4(list(range(8)))
  g["index_in_permuted"] = [[x, 16, get_index_in_permuted(x, 16)] for x in range(16)]
  # LDE + multi-column commit: 4 columns, steps=256, ext=8
  steps, ext = 256, 8
  N = steps * ext
  G2 = F(7)**((P - 1) // N)
  G1 = G2**ext
  evs = []
  for c in range(4):
    tr = [F(synth(c, i)) for i in range(steps)]
    poly = NonBinaryFFT(F, G1).inv_fft(tr)
    evs.append(NonBinaryFFT(F, G2).fft(poly))
  mt = merkelize_polynomial_evaluations(4, evs)
  g["lde_commit"] = {"steps": steps, "ext": ext, "cols": 4, "root": mt[1].hex(),
                     "H_cols": [H([int(x) for x in e]) for e in evs],
                     "branch5": [b.hex() for b in mk_branch(mt, 5)],
                     "tree_digest": hashlib.blake2s(b"".join(mt)).hexdigest()}
  json.dump(g, open(os.path.join(OUTX, "merkle.json"), "w"), indent=1)

  # ---------------------------------------------------------------------------- fri
  g = {}
  n = 64
  w = F(7)**((P - 1) // n)
  xs = get_power_cycle(w, F)
  vals = [F(synth(5, i)) for i in range(n)]
  sx = F(bytes([0xAB]) * 32)  # unreduced bytes constructor, as fri.py:229
  q = n // 4
  with pyref.quiet():
    xp = multi_interp_4(F, [[xs[i + q * j] for j in range(4)] for i in range(q)],
                        [[vals[i + q * j] for j in range(4)] for i in range(q)])
  col = [p_(sx) for p_ in xp]
  g["fold64"] = {"n": n, "col": 5, "root": hx(int(w)), "special_x": hx(sx.n), "column": [hx(int(c)) for c in col]}
  g["proofs"] = []
  polysF = polynomials_over(F).factory
  for (logn, deg, ex) in [(8, 32, 0), (10, 128, 8), (12, 512, 8)]:
    n = 1 << logn
    w = F(7)**((P - 1) // n)
    f = polysF([F(synth(9, i)) for i in range(deg)])
    t0 = time.time()
    with pyref.quiet():
      prf = SmoothSubgroupFRI(F).generate_proximity_proof(f, w, deg, exclude_multiples_of=ex)
      evs_ = NonBinaryFFT(F, w).fft(f)
      mroot = merkelize(evs_)[1]
      assert SmoothSubgroupFRI(F).verify_proximity_proof(prf, mroot, w, deg, exclude_multiples_of=ex)
    g["proofs"].append({"logn": logn, "deg": deg, "exclude": ex, "col": 9, "merkle_root": mroot.hex(),
                        "layers": len(prf), "roots": [layer[0].hex() for layer in prf[:-1]],
                        "final_len": len(prf[-1][0]) if False else len(prf[-1]),
                        "digest": proof_digest(prf)})
    print("fri 2^%d: %.2fs" % (logn, time.time() - t0))
  json.dump(g, open(os.path.join(OUTX, "fri.json"), "w"), indent=1)

  # -------------------------------------------------------------------------- stark
  g = {"proofs": []}

  def run(tag, width, steps, inp, mk_step_polys, sp_desc):
    Xs = generate_Xi_s(F, width)
    step_polys = mk_step_polys(Xs)
    with pyref.quiet():
      trace, output = get_computational_trace([F(v) for v in inp], steps, width, step_polys)
      witness = [[trace[i][j] for i in range(steps)] for j in range(width)]
      boundary = [(0, j, F(inp[j])) for j in range(width)]
      S = stark_mod.STARK(F, steps, 8, width, step_polys)
      t0 = time.time()
      proof = S.mk_proof(witness, boundary)
      dt = time.time() - t0
      ok = S.verify_proof(proof, witness, boundary)
    assert ok
    m_root, l_root, branches, fri = proof
    g["proofs"].append({"tag": tag, "width": width, "steps": steps, "ext": 8, "inp": inp, "step_polys": sp_desc,
                        "output": [hx(int(x)) for x in output], "m_root": m_root.hex(), "l_root": l_root.hex(),
                        "n_branches": len(branches), "fri_layers": len(fri),
                        "fri_roots": [layer[0].hex() for layer in fri[:-1]],
                        "branch0": [b.hex() for b in branches[0]],
                        "digest": proof_digest(proof), "ref_seconds": round(dt, 3)})
    print("stark %s: %.2fs" % (tag, dt))

  # step_polys description: list of {exponent-tuple-as-list-string: coeff}
  fib = lambda X: [X[1], X[0] + X[1]]
  fib_desc = [{"0,1": 1}, {"1,0": 1, "0,1": 1}]
  run("fib32", 2, 32, [0, 1], fib, fib_desc)
  run("fib8", 2, 8, [0, 1], fib, fib_desc)
  run("cubic8", 2, 8, [2, 5], lambda X: [X[0], X[0] + X[1]**3], [{"1,0": 1}, {"1,0": 1, "0,3": 1}])
  run("quad128", 2, 128, [2, 5], lambda X: [X[1], X[0] + 2 * X[1]**2], [{"0,1": 1}, {"1,0": 1, "0,2": 2}])
  run("affine32", 2, 32, [2, 5], lambda X: [X[0], X[0] + 3 * X[1]], [{"1,0": 1}, {"1,0": 1, "0,1": 3}])
  run("w3_8", 3, 8, [2, 2, 5], lambda X: [X[0], X[1], X[0] + X[1] * X[2]**2],
      [{"1,0,0": 1}, {"0,1,0": 1}, {"1,0,0": 1, "0,1,2": 1}])
  run("w6_8", 6, 8, [1, 2, 3, 4, 5, 6],
      lambda X: [X[0], X[1], X[2], X[3], X[4], X[0] * X[1] * X[2] * X[3] * X[4] * X[5]],
      [{"1,0,0,0,0,0": 1}, {"0,1,0,0,0,0": 1}, {"0,0,1,0,0,0": 1}, {"0,0,0,1,0,0": 1}, {"0,0,0,0,1,0": 1},
       {"1,1,1,1,1,1": 1}])
  run("fib256", 2, 256, [0, 1], fib, fib_desc)
  if degrees:
    # constraint degrees the other files do not reach (they hold 1, 2, 3, 6): 4, 5, 7 and 8 = the
    # extension factor, the largest the reference's domain supports (deg C < N), in the style of
    # the commented starks/test/test_stark.py:268-350 (cubic "MiMC-like", varying quintic)
    g = {"proofs": []}
    run("deg4_8", 2, 8, [2, 5], lambda X: [X[0], X[0] + X[1]**4], [{"1,0": 1}, {"1,0": 1, "0,4": 1}])
    run("deg5_16", 2, 16, [3, 2], lambda X: [X[1], X[0] + X[1]**5], [{"0,1": 1}, {"1,0": 1, "0,5": 1}])
    run("deg7_8", 2, 8, [2, 3], lambda X: [X[0], X[0] + 2 * X[1]**7], [{"1,0": 1}, {"1,0": 1, "0,7": 2}])
    run("deg8_8", 2, 8, [5, 2], lambda X: [X[0], 3 * X[0] + X[1]**8], [{"1,0": 1}, {"1,0": 3, "0,8": 1}])
    run("deg8_32", 2, 32, [5, 2], lambda X: [X[1], 3 * X[0] + X[1]**8], [{"0,1": 1}, {"1,0": 3, "0,8": 1}])
    run("w4_mixed_16", 4, 16, [1, 2, 3, 4],
        lambda X: [X[1], X[2] * X[3], X[0]**2 * X[1]**2 + X[3], X[0] + X[1] * X[2] * X[3]**3],
        [{"0,1,0,0": 1}, {"0,0,1,1": 1}, {"2,2,0,0": 1, "0,0,0,1": 1}, {"1,0,0,0": 1, "0,1,1,3": 1}])
    json.dump(g, open(os.path.join(OUT, "stark_degrees.json"), "w"), indent=1)
    print("golden vectors written to", OUT)
    return
  if big:
    run("fib1024", 2, 1024, [0, 1], fib, fib_desc)
  json.dump(g, open(os.path.join(OUT, "stark.json" if not big else "stark_big.json"), "w"), indent=1)
  print("golden vectors written to", OUT)


if __name__ == "__main__":
  main()
