// hostmath.h -- tiny host-side 256-bit modular arithmetic (set-up only: roots, inverse
// of N, Montgomery constants).  Never on a hot path; a handful of calls per transform.
#pragma once
#include <stdint.h>
#include <string.h>
#include "field.cuh"

namespace stk {
namespace host {

inline fe from_u64(uint64_t x) {
  fe r = fe_zero();
  r.v[0] = (uint32_t)x; r.v[1] = (uint32_t)(x >> 32);
  return r;
}
inline fe addmod(const fe& a, const fe& b, const fe& p) {
  fe s;
  uint32_t k = add8(s.v, a.v, b.v);
  if (k || geq8(s.v, p.v)) sub8(s.v, s.v, p.v);
  return s;
}
inline fe submod(const fe& a, const fe& b, const fe& p) {
  fe d;
  if (sub8(d.v, a.v, b.v)) add8(d.v, d.v, p.v);
  return d;
}
// x mod p for arbitrary 256-bit x (binary long reduction).
inline fe reduce(const fe& x, const fe& p) {
  fe r = fe_zero();
  for (int i = 255; i >= 0; --i) {
    r = addmod(r, r, p);
    if ((x.v[i >> 5] >> (i & 31)) & 1u) r = addmod(r, from_u64(1), p);
  }
  return r;
}
// a*b mod p, double-and-add (256 iterations), a,b < p.
inline fe mulmod(const fe& a, const fe& b, const fe& p) {
  fe r = fe_zero();
  for (int i = 255; i >= 0; --i) {
    r = addmod(r, r, p);
    if ((b.v[i >> 5] >> (i & 31)) & 1u) r = addmod(r, a, p);
  }
  return r;
}
inline fe powmod(const fe& a, const fe& e, const fe& p) {
  fe r = reduce(from_u64(1), p);
  fe base = a;
  for (int i = 0; i < 256; ++i) {
    if ((e.v[i >> 5] >> (i & 31)) & 1u) r = mulmod(r, base, p);
    base = mulmod(base, base, p);
  }
  return r;
}
inline fe pow_u64(const fe& a, uint64_t e, const fe& p) { return powmod(a, from_u64(e), p); }
// a^{-1} = a^{p-2} (p prime)
inline fe invmod(const fe& a, const fe& p) {
  fe e;
  fe two = from_u64(2);
  sub8(e.v, p.v, two.v);
  return powmod(a, e, p);
}
inline bool is_stark_prime(const fe& p) {
  fe q = StarkField::modulus();
  return fe_eq(p, q);
}
// Montgomery constants for an odd modulus.
inline bool mont_setup(MontField* F, const fe& p) {
  if (!(p.v[0] & 1u)) return false;
  fe one = from_u64(1);
  if (geq8(one.v, p.v)) return false;
  F->p = p;
  uint32_t inv = 1;
  for (int i = 0; i < 5; ++i) inv *= 2u - p.v[0] * inv;
  F->ninv = 0u - inv;
  fe x = one;
  for (int i = 0; i < 512; ++i) {
    x = addmod(x, x, p);
    if (i == 255) F->rone = x;
  }
  F->r2 = x;
  return true;
}

}  // namespace host
}  // namespace stk
