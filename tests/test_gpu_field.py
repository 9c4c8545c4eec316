"""Field arithmetic corner cases (starks/modp.py:31-53: `(a op b) % p`) through the pointwise
entry point stk_vec_op: operands that stress every carry / fold / fix-up path of the 256-bit
multiply and of the modular add / sub -- all-ones limbs, values next to p, to 2^256 - p and to
limb boundaries -- in all pairs, plus random operands biased towards saturated limbs.
Bit-exact against Python integers, for the STARK prime and for run-time moduli."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

P = 2**256 - 351 * 2**32 + 1
SECP = 2**256 - 2**32 - 977


@pytest.fixture(scope="module")
def eng():
  from starks_b200 import Engine
  e = Engine(0)
  yield e
  e.set_field(P)
  e.close()


def specials(p):
  c = (1 << 256) - p if p.bit_length() == 256 else 1
  vals = {0, 1, 2, 3, p - 1, p - 2, p - 3, p // 2, p // 2 + 1, c % p, (c - 1) % p, (c + 1) % p, (p - c) % p}
  for k in (16, 31, 32, 33, 63, 64, 65, 96, 127, 128, 129, 160, 192, 223, 224, 255):
    for d in (-1, 0, 1):
      vals.add(((1 << k) + d) % p)
  for mask in (0x00000000FFFFFFFF, 0xFFFFFFFF00000000):
    v = 0
    for i in range(4):
      v |= mask << (64 * i)
    vals.add(v % p)
  vals.add(int("ff" * 31 + "00", 16) % p)
  vals.add(int("0f" * 32, 16) % p)
  vals.add((351 << 32) % p)
  vals.add(((351 << 32) * ((1 << 200) - 1)) % p)
  return sorted(vals)


def run_ops(eng, p, a, b):
  from starks_b200.limbs import ints_to_limbs, limbs_to_ints
  la, lb = ints_to_limbs(a), ints_to_limbs(b)
  n = len(a)
  da, db, do = eng.alloc(la.nbytes).upload(la), eng.alloc(lb.nbytes).upload(lb), eng.alloc(la.nbytes)
  out = []
  for op in (0, 1, 2):
    eng._check(eng.lib.stk_vec_op(eng.ctx, op, da.ptr, db.ptr, do.ptr, n))
    out.append(limbs_to_ints(do.download((n, 8))))
  for buf in (da, db, do):
    buf.free()
  return out


@pytest.mark.parametrize("p", [P, SECP, 2**255 - 19, 31, 7, 2**61 - 1], ids=["stark", "secp256k1", "25519", "31", "7", "m61"])
def test_special_operand_pairs(eng, p):
  eng.set_field(p)
  sp = specials(p)
  a = [x for x in sp for _ in sp]
  b = [y for _ in sp for y in sp]
  add, sub, mul = run_ops(eng, p, a, b)
  for i, (x, y) in enumerate(zip(a, b)):
    assert add[i] == (x + y) % p, ("add", hex(x), hex(y))
    assert sub[i] == (x - y) % p, ("sub", hex(x), hex(y))
    assert mul[i] == (x * y) % p, ("mul", hex(x), hex(y))


@pytest.mark.parametrize("p", [P, SECP], ids=["stark", "secp256k1"])
def test_random_saturated_limbs(eng, p):
  eng.set_field(p)
  rng = np.random.default_rng(77)
  n = 1 << 16

  def draw():
    limbs = rng.integers(0, 2**32, size=(n, 8), dtype=np.uint64)
    kind = rng.integers(0, 4, size=(n, 8))
    limbs[kind == 0] = 0xFFFFFFFF
    limbs[kind == 1] = 0
    vals = [sum(int(l) << (32 * i) for i, l in enumerate(row)) % p for row in limbs]
    return vals

  a, b = draw(), draw()
  add, sub, mul = run_ops(eng, p, a, b)
  assert add == [(x + y) % p for x, y in zip(a, b)]
  assert sub == [(x - y) % p for x, y in zip(a, b)]
  assert mul == [(x * y) % p for x, y in zip(a, b)]
