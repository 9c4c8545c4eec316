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
// 4x64-limb Montgomery context for the host (per modulus, cached thread-locally).
struct HostMont {
  fe p;
  uint64_t m[4];    // modulus
  uint64_t r2[4];   // 2^512 mod p
  uint64_t ninv;    // -p^-1 mod 2^64
  bool valid = false;
};
inline void to64(const fe& a, uint64_t* o) {
  for (int i = 0; i < 4; ++i) o[i] = (uint64_t)a.v[2 * i] | ((uint64_t)a.v[2 * i + 1] << 32);
}
inline fe from64(const uint64_t* a) {
  fe r;
  for (int i = 0; i < 4; ++i) { r.v[2 * i] = (uint32_t)a[i]; r.v[2 * i + 1] = (uint32_t)(a[i] >> 32); }
  return r;
}
inline void mont64(const HostMont& H, const uint64_t* a, const uint64_t* b, uint64_t* out) {
  typedef unsigned __int128 u128;
  uint64_t t[6] = {0, 0, 0, 0, 0, 0};
  for (int i = 0; i < 4; ++i) {
    u128 c = 0;
    for (int j = 0; j < 4; ++j) { c += (u128)a[j] * b[i] + t[j]; t[j] = (uint64_t)c; c >>= 64; }
    c += t[4]; t[4] = (uint64_t)c; t[5] = (uint64_t)(c >> 64);
    uint64_t mm = t[0] * H.ninv;
    c = (u128)mm * H.m[0] + t[0]; c >>= 64;
    for (int j = 1; j < 4; ++j) { c += (u128)mm * H.m[j] + t[j]; t[j - 1] = (uint64_t)c; c >>= 64; }
    c += t[4]; t[3] = (uint64_t)c; t[4] = t[5] + (uint64_t)(c >> 64);
  }
  bool ge = t[4] != 0;
  if (!ge) {
    ge = true;
    for (int i = 3; i >= 0; --i) { if (t[i] > H.m[i]) break; if (t[i] < H.m[i]) { ge = false; break; } }
  }
  if (ge) {
    uint64_t bw = 0;
    for (int i = 0; i < 4; ++i) { u128 d = (u128)t[i] - H.m[i] - bw; t[i] = (uint64_t)d; bw = (uint64_t)(d >> 64) & 1; }
  }
  for (int i = 0; i < 4; ++i) out[i] = t[i];
}
inline const HostMont& host_mont(const fe& p) {
  static thread_local HostMont H;
  if (H.valid && fe_eq(H.p, p)) return H;
  H.p = p;
  to64(p, H.m);
  uint64_t inv = 1;
  for (int i = 0; i < 6; ++i) inv *= 2 - H.m[0] * inv;
  H.ninv = (uint64_t)0 - inv;
  fe x = from_u64(1);
  for (int i = 0; i < 512; ++i) x = addmod(x, x, p);
  to64(x, H.r2);
  H.valid = true;
  return H;
}
// a*b mod p (p odd), a, b < p: mont(mont(a, b), R^2) = a*b
inline fe mulmod(const fe& a, const fe& b, const fe& p) {
  const HostMont& H = host_mont(p);
  uint64_t x[4], y[4], t[4];
  to64(a, x); to64(b, y);
  mont64(H, x, y, t);
  mont64(H, t, H.r2, t);
  return from64(t);
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
// small exponents (strides, orders): stop at the top bit instead of walking all 256
inline fe pow_u64(const fe& a, uint64_t e, const fe& p) {
  fe r = reduce(from_u64(1), p);
  fe base = a;
  for (; e; e >>= 1) {
    if (e & 1) r = mulmod(r, base, p);
    if (e > 1) base = mulmod(base, base, p);
  }
  return r;
}
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
