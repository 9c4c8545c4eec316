// field.cuh -- 256-bit prime-field arithmetic on the B200 integer pipes (8 x u32 limbs).
//
// Replaces the reference's IntegerModP object arithmetic (starks/modp.py:31-53: every
// + - * is a Python bigint op followed by `% p`).  Two field policies share one kernel
// code base (kernels are templated on the policy):
//
//   StarkField  p = 2^256 - 351*2^32 + 1 (starks/utils.py:22, starks/stark.py:217).
//               Residues stay canonical and in the plain domain.  The 512-bit product is
//               built from 64 IMAD.WIDE.U32(.X) (even/odd column chains, carries in
//               predicates) and reduced with the pseudo-Mersenne identity
//               2^256 = 351*2^32 - 1 (mod p): two folds (8 + 2 wide multiplies by 351) and
//               a 3-limb fix-up, instead of 8 Montgomery rounds.
//   MontField   any odd modulus < 2^256 given at run time (the reference tests use
//               p = 31 and p = 7).  Classic CIOS Montgomery, R = 2^256; data stays in the
//               plain domain and only twiddles / constants are kept as x*R, so that
//               mont(x, wR) = x*w needs no domain conversion of the data.
//
// Both expose:  add, sub, mul_tw(a, t) (t in "twiddle form": plain for StarkField,
// Montgomery for MontField), to_tw / from_tw, and are bit-exact with `(a op b) % p`.
#pragma once
#include <stdint.h>
#include <cuda_runtime.h>

namespace stk {

struct __align__(16) fe {
  uint32_t v[8];
};

__host__ __device__ __forceinline__ fe fe_zero() {
  fe r;
#pragma unroll
  for (int i = 0; i < 8; ++i) r.v[i] = 0;
  return r;
}
__host__ __device__ __forceinline__ bool fe_is_zero(const fe& a) {
  uint32_t o = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) o |= a.v[i];
  return o == 0;
}
__host__ __device__ __forceinline__ bool fe_eq(const fe& a, const fe& b) {
  uint32_t o = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) o |= a.v[i] ^ b.v[i];
  return o == 0;
}

// 128-bit vector load/store of one element (two LDG.128 / STG.128).
__device__ __forceinline__ fe fe_load(const fe* p) {
  const uint4* q = reinterpret_cast<const uint4*>(p);
  uint4 a = q[0], b = q[1];
  fe r;
  r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
  r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
  return r;
}
__device__ __forceinline__ fe fe_load_ro(const fe* p) {  // read-only path (twiddle tables)
  const uint4* q = reinterpret_cast<const uint4*>(p);
  uint4 a = __ldg(q), b = __ldg(q + 1);
  fe r;
  r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
  r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
  return r;
}
__device__ __forceinline__ void fe_store(fe* p, const fe& a) {
  uint4* q = reinterpret_cast<uint4*>(p);
  q[0] = make_uint4(a.v[0], a.v[1], a.v[2], a.v[3]);
  q[1] = make_uint4(a.v[4], a.v[5], a.v[6], a.v[7]);
}

// ----------------------------------------------------------------------------------
// 256 x 256 -> 512-bit product.  Products a_i*b_j with i+j even land on even limb
// boundaries and are accumulated in E, the others in O (one limb up); inside one row
// the four 64-bit products do not overlap, so a row is ONE carry chain of
// mad.lo.cc / madc.hi.cc pairs, each of which ptxas fuses into a single
// IMAD.WIDE.U32.X with the carry in a predicate register.
// ----------------------------------------------------------------------------------
#ifdef __CUDA_ARCH__
__device__ __forceinline__ void mul_row_first(uint32_t* acc, const uint32_t* a, uint32_t b) {
#pragma unroll
  for (int j = 0; j < 8; j += 2)
    asm volatile("mul.lo.u32 %0, %2, %3; mul.hi.u32 %1, %2, %3;"
        : "=&r"(acc[j]), "=r"(acc[j + 1]) : "r"(a[j]), "r"(b));  // %0 is written before %2 is last read
}
__device__ __forceinline__ void mad_row(uint32_t* acc, const uint32_t* a, uint32_t b) {
  asm volatile("mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;"
      : "+&r"(acc[0]), "+r"(acc[1]) : "r"(a[0]), "r"(b));
#pragma unroll
  for (int j = 2; j < 8; j += 2)
    asm volatile("madc.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;"
        : "+&r"(acc[j]), "+r"(acc[j + 1]) : "r"(a[j]), "r"(b));
  asm volatile("addc.u32 %0, 0, 0;" : "=r"(acc[8]));
}
__device__ __forceinline__ void mul512(uint32_t* T, const uint32_t* a, const uint32_t* b) {
  uint32_t E[16], O[16];  // O[k] holds limb k+1
#pragma unroll
  for (int i = 0; i < 16; ++i) { E[i] = 0; O[i] = 0; }
  mul_row_first(E, a, b[0]);
  mul_row_first(O, a + 1, b[0]);
#pragma unroll
  for (int j = 1; j < 8; ++j) {
    if (j & 1) { mad_row(E + j + 1, a + 1, b[j]); mad_row(O + j - 1, a, b[j]); }
    else       { mad_row(E + j, a, b[j]);         mad_row(O + j, a + 1, b[j]); }
  }
  T[0] = E[0];
  asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(T[1]) : "r"(E[1]), "r"(O[0]));
#pragma unroll
  for (int i = 2; i < 15; ++i) asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(T[i]) : "r"(E[i]), "r"(O[i - 1]));
  asm volatile("addc.u32 %0, %1, %2;" : "=r"(T[15]) : "r"(E[15]), "r"(O[14]));
}
#else
// Host twin (same result), used by host-side table set-up and CPU unit tests of the
// reduction logic.
inline void mul512(uint32_t* T, const uint32_t* a, const uint32_t* b) {
  for (int i = 0; i < 16; ++i) T[i] = 0;
  for (int j = 0; j < 8; ++j) {
    uint32_t carry = 0;
    for (int i = 0; i < 8; ++i) {
      uint64_t t = (uint64_t)a[i] * b[j] + T[i + j] + carry;
      T[i + j] = (uint32_t)t;
      carry = (uint32_t)(t >> 32);
    }
    T[j + 8] = carry;
  }
}
#endif

// Portable add/sub with carry used by both host and device generic paths.
__host__ __device__ __forceinline__ uint32_t add8(uint32_t* r, const uint32_t* a, const uint32_t* b) {
  uint64_t c = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) { c += (uint64_t)a[i] + b[i]; r[i] = (uint32_t)c; c >>= 32; }
  return (uint32_t)c;
}
__host__ __device__ __forceinline__ uint32_t sub8(uint32_t* r, const uint32_t* a, const uint32_t* b) {
  uint32_t bw = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    uint64_t d = (uint64_t)a[i] - b[i] - bw;
    r[i] = (uint32_t)d;
    bw = (uint32_t)(d >> 32) & 1u;
  }
  return bw;
}
__host__ __device__ __forceinline__ bool geq8(const uint32_t* a, const uint32_t* b) {
#pragma unroll
  for (int i = 7; i >= 0; --i) {
    if (a[i] > b[i]) return true;
    if (a[i] < b[i]) return false;
  }
  return true;
}

// ----------------------------------------------------------------------------------
// StarkField: p = 2^256 - c,  c = 351*2^32 - 1 = {0xFFFFFFFF, 350, 0, ...}
// limbs of p: {1, 0xFFFFFEA1, 0xFFFFFFFF x 6}
// ----------------------------------------------------------------------------------
struct StarkField {
  static constexpr bool kMontgomery = false;

  __host__ __device__ static __forceinline__ fe modulus() {
    fe p;
    p.v[0] = 1u; p.v[1] = 0xFFFFFEA1u;
#pragma unroll
    for (int i = 2; i < 8; ++i) p.v[i] = 0xFFFFFFFFu;
    return p;
  }
  __host__ __device__ __forceinline__ fe one_tw() const { fe r = fe_zero(); r.v[0] = 1; return r; }

  // x in [0, 2^256) -> canonical.  x >= p iff limbs 2..7 are all ones and the low 64
  // bits are >= 0xFFFFFEA1_00000001; then x - p has only the low 64 bits set.
  __host__ __device__ static __forceinline__ void canon(uint32_t* B) {
    if ((B[2] & B[3] & B[4] & B[5] & B[6] & B[7]) == 0xFFFFFFFFu) {
      uint64_t lo = ((uint64_t)B[1] << 32) | B[0];
      if (lo >= 0xFFFFFEA100000001ull) {
        lo -= 0xFFFFFEA100000001ull;
        B[0] = (uint32_t)lo; B[1] = (uint32_t)(lo >> 32);
        B[2] = B[3] = B[4] = B[5] = B[6] = B[7] = 0;
      }
    }
  }

#ifdef __CUDA_ARCH__
  // T (16 limbs) mod p, canonical.
  __device__ static __forceinline__ fe reduce512(const uint32_t* T) {
    const uint32_t* L = T;
    const uint32_t* H = T + 8;
    // fold 1:  L + H*2^256 = L + (351*H << 32) - H
    // U = 351*H (9 limbs): even limbs of H give four non-overlapping 41-bit products,
    // the odd ones are accumulated one limb up in a single carry chain.
#ifdef STK_REDUCE_SPLIT
    // Measured alternative (tools/gpu_altlib.py): 351*H as two sets of four independent 41-bit
    // products (even limbs, odd limbs one up) merged by two add chains -- no accumulating wide
    // multiply, no register-pair realignment moves.  +2 % on the in-register butterfly
    // (87.9 -> 89.7 G/s), nothing on the transform (8.82 vs 8.84 ms): not the default.
    uint32_t UE[8], UO[8];
    mul_row_first(UE, H, 351u);
    mul_row_first(UO, H + 1, 351u);
    uint32_t B[10];
    B[0] = L[0];
    asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(B[1]) : "r"(L[1]), "r"(UE[0]));
#pragma unroll
    for (int i = 2; i < 8; ++i) asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(B[i]) : "r"(L[i]), "r"(UE[i - 1]));
    asm volatile("addc.cc.u32 %0, %1, 0;" : "=r"(B[8]) : "r"(UE[7]));
    asm volatile("addc.u32 %0, 0, 0;" : "=r"(B[9]));
    asm volatile("add.cc.u32 %0, %0, %1;" : "+r"(B[2]) : "r"(UO[0]));
#pragma unroll
    for (int i = 3; i < 9; ++i) asm volatile("addc.cc.u32 %0, %0, %1;" : "+r"(B[i]) : "r"(UO[i - 2]));
    asm volatile("addc.u32 %0, %0, %1;" : "+r"(B[9]) : "r"(UO[7]));
#else
    uint32_t U[10];
    U[8] = 0;
    mul_row_first(U, H, 351u);
    mad_row(U + 1, H + 1, 351u);
    uint32_t B[10];
    B[0] = L[0];
    asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(B[1]) : "r"(L[1]), "r"(U[0]));
#pragma unroll
    for (int i = 2; i < 8; ++i) asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(B[i]) : "r"(L[i]), "r"(U[i - 1]));
    asm volatile("addc.cc.u32 %0, %1, 0;" : "=r"(B[8]) : "r"(U[7]));
    asm volatile("addc.u32 %0, %1, 0;" : "=r"(B[9]) : "r"(U[8]));
#endif
    asm volatile("sub.cc.u32 %0, %0, %1;" : "+r"(B[0]) : "r"(H[0]));
#pragma unroll
    for (int i = 1; i < 8; ++i) asm volatile("subc.cc.u32 %0, %0, %1;" : "+r"(B[i]) : "r"(H[i]));
    asm volatile("subc.cc.u32 %0, %0, 0;" : "+r"(B[8]));
    asm volatile("subc.u32 %0, %0, 0;" : "+r"(B[9]));
    // fold 2: Vhi = B8 + B9*2^32 (< 2^42);  B[0..8) += (351*Vhi << 32) - Vhi
    uint64_t w = (uint64_t)B[8] * 351u;
    uint32_t W0 = (uint32_t)w, W1 = (uint32_t)(w >> 32) + B[9] * 351u;
    uint32_t k1, k2;
    asm volatile("add.cc.u32 %0, %0, %1;" : "+r"(B[1]) : "r"(W0));
    asm volatile("addc.cc.u32 %0, %0, %1;" : "+r"(B[2]) : "r"(W1));
#pragma unroll
    for (int i = 3; i < 8; ++i) asm volatile("addc.cc.u32 %0, %0, 0;" : "+r"(B[i]));
    asm volatile("addc.u32 %0, 0, 0;" : "=r"(k1));
    asm volatile("sub.cc.u32 %0, %0, %1;" : "+r"(B[0]) : "r"(B[8]));
    asm volatile("subc.cc.u32 %0, %0, %1;" : "+r"(B[1]) : "r"(B[9]));
#pragma unroll
    for (int i = 2; i < 8; ++i) asm volatile("subc.cc.u32 %0, %0, 0;" : "+r"(B[i]));
    asm volatile("subc.u32 %0, 0, 0;" : "=r"(k2));  // 0 or 0xFFFFFFFF
    // net overflow k = k1 - borrow in {0,1}; if 1 the wrapped value is < 2^84 and
    // 2^256 = c (mod p) is added to its three low limbs.
    uint32_t m = 0u - (k1 + k2);
    asm volatile("add.cc.u32 %0, %0, %1;" : "+r"(B[0]) : "r"(m));
    asm volatile("addc.cc.u32 %0, %0, %1;" : "+r"(B[1]) : "r"(m & 350u));
    asm volatile("addc.u32 %0, %0, 0;" : "+r"(B[2]));
    canon(B);
    fe r;
#pragma unroll
    for (int i = 0; i < 8; ++i) r.v[i] = B[i];
    return r;
  }
  __device__ __forceinline__ fe add(const fe& a, const fe& b) const {
    fe s;
    uint32_t k;
    asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(s.v[0]) : "r"(a.v[0]), "r"(b.v[0]));
#pragma unroll
    for (int i = 1; i < 8; ++i) asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(s.v[i]) : "r"(a.v[i]), "r"(b.v[i]));
    asm volatile("addc.u32 %0, 0, 0;" : "=r"(k));
    // carry out: a + b - 2^256 + c = a + b - p, and a + b < 2p so the result is canonical
    uint32_t m = 0u - k;
    asm volatile("add.cc.u32 %0, %0, %1;" : "+r"(s.v[0]) : "r"(m));
    asm volatile("addc.cc.u32 %0, %0, %1;" : "+r"(s.v[1]) : "r"(m & 350u));
#pragma unroll
    for (int i = 2; i < 7; ++i) asm volatile("addc.cc.u32 %0, %0, 0;" : "+r"(s.v[i]));
    asm volatile("addc.u32 %0, %0, 0;" : "+r"(s.v[7]));
    canon(s.v);  // no carry but a + b in [p, 2^256)
    return s;
  }
  __device__ __forceinline__ fe sub(const fe& a, const fe& b) const {
    fe d;
    uint32_t m;
    asm volatile("sub.cc.u32 %0, %1, %2;" : "=r"(d.v[0]) : "r"(a.v[0]), "r"(b.v[0]));
#pragma unroll
    for (int i = 1; i < 8; ++i) asm volatile("subc.cc.u32 %0, %1, %2;" : "=r"(d.v[i]) : "r"(a.v[i]), "r"(b.v[i]));
    asm volatile("subc.u32 %0, 0, 0;" : "=r"(m));  // borrow -> 0xFFFFFFFF
    // borrow: a - b + 2^256 - c = a - b + p
    asm volatile("sub.cc.u32 %0, %0, %1;" : "+r"(d.v[0]) : "r"(m));
    asm volatile("subc.cc.u32 %0, %0, %1;" : "+r"(d.v[1]) : "r"(m & 350u));
#pragma unroll
    for (int i = 2; i < 7; ++i) asm volatile("subc.cc.u32 %0, %0, 0;" : "+r"(d.v[i]));
    asm volatile("subc.u32 %0, %0, 0;" : "+r"(d.v[7]));
    return d;
  }
#else
  // Host twins.
  static inline fe reduce512(const uint32_t* T) {
    // same two folds, written with 64-bit temporaries
    uint32_t U[9];
    uint32_t carry = 0;
    for (int i = 0; i < 8; ++i) {
      uint64_t t = (uint64_t)T[8 + i] * 351u + carry;
      U[i] = (uint32_t)t; carry = (uint32_t)(t >> 32);
    }
    U[8] = carry;
    uint32_t B[10];
    uint64_t c = 0;
    for (int i = 0; i < 10; ++i) {
      c += (uint64_t)(i < 8 ? T[i] : 0) + (i >= 1 ? U[i - 1] : 0);
      B[i] = (uint32_t)c; c >>= 32;
    }
    uint32_t bw = 0;
    for (int i = 0; i < 10; ++i) {
      uint64_t d = (uint64_t)B[i] - (i < 8 ? T[8 + i] : 0) - bw;
      B[i] = (uint32_t)d; bw = (uint32_t)(d >> 32) & 1u;
    }
    uint64_t w = (uint64_t)B[8] * 351u;
    uint32_t W[8] = {0, (uint32_t)w, (uint32_t)(w >> 32) + B[9] * 351u, 0, 0, 0, 0, 0};
    uint32_t Vh[8] = {B[8], B[9], 0, 0, 0, 0, 0, 0};
    uint32_t k1 = add8(B, B, W);
    uint32_t k2 = sub8(B, B, Vh);
    if (k1 - k2) {
      uint32_t C[8] = {0xFFFFFFFFu, 350u, 0, 0, 0, 0, 0, 0};
      add8(B, B, C);
    }
    canon(B);
    fe r;
    for (int i = 0; i < 8; ++i) r.v[i] = B[i];
    return r;
  }
  inline fe add(const fe& a, const fe& b) const {
    fe s; fe p = modulus();
    uint32_t k = add8(s.v, a.v, b.v);
    if (k || geq8(s.v, p.v)) sub8(s.v, s.v, p.v);
    return s;
  }
  inline fe sub(const fe& a, const fe& b) const {
    fe d; fe p = modulus();
    if (sub8(d.v, a.v, b.v)) add8(d.v, d.v, p.v);
    return d;
  }
#endif
  __host__ __device__ __forceinline__ fe mul_tw(const fe& a, const fe& t) const {
    uint32_t T[16];
    mul512(T, a.v, t.v);
    return reduce512(T);
  }
  __host__ __device__ __forceinline__ fe to_tw(const fe& a) const { return a; }
  __host__ __device__ __forceinline__ fe from_tw(const fe& a) const { return a; }
  // arbitrary 256-bit value -> canonical residue (IntegerModP.__init__, modp.py:35-36)
  __host__ __device__ __forceinline__ fe reduce(const fe& a) const {
    fe r = a;
    canon(r.v);
    return r;
  }
};

// ----------------------------------------------------------------------------------
// MontField: run-time odd modulus, CIOS Montgomery with R = 2^256.
// ----------------------------------------------------------------------------------
struct MontField {
  static constexpr bool kMontgomery = true;
  fe p;        // modulus
  fe r2;       // R^2 mod p
  fe rone;     // R mod p
  uint32_t ninv;  // -p^{-1} mod 2^32

  __host__ __device__ __forceinline__ fe one_tw() const { return rone; }

  __host__ __device__ __forceinline__ fe add(const fe& a, const fe& b) const {
    fe s;
    uint32_t k = add8(s.v, a.v, b.v);
    if (k || geq8(s.v, p.v)) sub8(s.v, s.v, p.v);
    return s;
  }
  __host__ __device__ __forceinline__ fe sub(const fe& a, const fe& b) const {
    fe d;
    if (sub8(d.v, a.v, b.v)) add8(d.v, d.v, p.v);
    return d;
  }
  // a*b/R mod p.  Deliberately NOT inlined on the device: the run-time-modulus path is for
  // small test fields, and one out-of-line copy keeps the kernels' compile time in seconds.
  __host__ __device__ __noinline__ fe mmul(const fe& a, const fe& b) const {
    uint32_t t[10];
#pragma unroll
    for (int i = 0; i < 10; ++i) t[i] = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      uint64_t c = 0;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        c += (uint64_t)a.v[j] * b.v[i] + t[j];
        t[j] = (uint32_t)c; c >>= 32;
      }
      c += t[8]; t[8] = (uint32_t)c; t[9] = (uint32_t)(c >> 32);
      uint32_t m = t[0] * ninv;
      c = (uint64_t)m * p.v[0] + t[0]; c >>= 32;
#pragma unroll
      for (int j = 1; j < 8; ++j) {
        c += (uint64_t)m * p.v[j] + t[j];
        t[j - 1] = (uint32_t)c; c >>= 32;
      }
      c += t[8]; t[7] = (uint32_t)c; t[8] = t[9] + (uint32_t)(c >> 32);
    }
    fe s;
#pragma unroll
    for (int i = 0; i < 8; ++i) s.v[i] = t[i];
    if (t[8] || geq8(s.v, p.v)) sub8(s.v, s.v, p.v);
    return s;
  }
  __host__ __device__ __forceinline__ fe mul_tw(const fe& a, const fe& t) const { return mmul(a, t); }
  __host__ __device__ __forceinline__ fe to_tw(const fe& a) const { return mmul(a, r2); }
  __host__ __device__ __forceinline__ fe from_tw(const fe& a) const {
    fe one = fe_zero(); one.v[0] = 1;
    return mmul(a, one);
  }
  __host__ __device__ __forceinline__ fe reduce(const fe& a) const { return from_tw(mmul(a, r2)); }
};

// plain * plain -> plain for either field
template <class F>
__host__ __device__ __forceinline__ fe f_mul(const F& f, const fe& a, const fe& b) {
  return f.mul_tw(a, f.to_tw(b));
}

}  // namespace stk
