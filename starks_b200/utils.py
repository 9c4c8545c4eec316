"""Mirror of the hot-path helpers of starks/utils.py: get_power_cycle (:30-38) on the device,
get_pseudorandom_indices (:60-90) and is_a_power_of_2 on the host (Fiat-Shamir glue, a few
hundred bytes of hashing)."""
from hashlib import blake2s

from .engine import default_engine
from .limbs import limbs_to_ints
from .modp import element_to_int

blake = lambda x: blake2s(x).digest()  # starks/merkle_tree.py:5


def multiplicative_order(root: int, p: int, cap: int = 1 << 16) -> int:
  """Order of root in (Z/p)*: the reference finds it by walking the power cycle
  (starks/fft.py:319-321).  Power-of-two orders are found by repeated squaring; anything else
  falls back to the walk (tiny orders only, e.g. 6 in starks/test/test_fft.py:98-113)."""
  root %= p
  x, k = root, 0
  while x != 1 and k <= 40:
    x = x * x % p
    k += 1
  if x == 1:
    # order divides 2^k; it is exactly 2^k' for the first k' reaching 1
    return 1 << k
  n, x = 1, root
  while x != 1:
    x = x * root % p
    n += 1
    if n > cap:
      raise ValueError("the root's order is neither a power of two nor <= %d" % cap)
  return n


def get_power_cycle(r, field, engine=None):
  """starks/utils.py:30-38: [1, r, r^2, ...] up to (excluding) the return to 1."""
  p = field.p
  rn = element_to_int(r) % p
  n = multiplicative_order(rn, p)
  eng = engine or default_engine()
  eng.set_field(p)
  return [field(v) for v in limbs_to_ints(eng.power_cycle(rn, n))]


def get_pseudorandom_indices(entropy, modulus, count, exclude_multiples_of=0):
  """starks/utils.py:60-90: `count` indices below `modulus` from the 4-byte big-endian words of
  entropy | H(last 32 bytes) | H(...) ...; with exclude_multiples_of = e the draw is taken over the
  positions that are NOT multiples of e (x -> x + 1 + x // (e - 1))."""
  assert modulus < 2**24
  stream = bytes(entropy)
  while len(stream) < 4 * count:            # entropy expansion: chain BLAKE2s over the last digest
    stream += blake(stream[-32:])
  words = [int.from_bytes(stream[4 * k:4 * k + 4], "big") for k in range(count)]
  e = exclude_multiples_of
  if not e:
    return [w % modulus for w in words]
  allowed = modulus * (e - 1) // e          # how many positions are not multiples of e
  return [x + 1 + x // (e - 1) for x in (w % allowed for w in words)]


def is_a_power_of_2(x):
  """starks/utils.py:93-94 (x >= 1)."""
  return x & (x - 1) == 0
