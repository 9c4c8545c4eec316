/*
 * starks_oracle.c -- CPU restatement of the reference's hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product (starks_b200/) may link, load or
 * call this file; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference leg do, and there only as the checker or as the timed CPU arm.
 *
 * Parity status: PINNED.  The restatement is checked (tests/test_oracle_*.py) against
 *   - the reference's own known answers (starks/test/test_utils.py:20-30,
 *     starks/test/test_modpy.py:28-35,54-61, RFC 7693 App. B for BLAKE2s), and
 *   - golden vectors produced by importing the unmodified Python reference in the
 *     authoring container (oracle/gen_golden.py -> tests/golden/).
 *
 * Every function cites the reference file:line (relative to the upstream repo root)
 * whose algorithm it follows.  Field elements cross the ABI as canonical residues in
 * eight little-endian uint32 limbs (the same layout the CUDA library uses), so one
 * numpy buffer can be handed to both sides.  Internally arithmetic is 4x64-bit
 * Montgomery (R = 2^256) for an arbitrary odd modulus p < 2^256: the reference
 * computes (a*b) % p on Python bigints (starks/modp.py:43-53); Montgomery is only a
 * faster way to the same canonical residue.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef unsigned __int128 u128;
typedef struct { uint64_t v[4]; } fe;

typedef struct {
  fe p;          /* modulus */
  fe r2;         /* 2^512 mod p */
  fe one;        /* 2^256 mod p  (Montgomery 1) */
  uint64_t ninv; /* -p^{-1} mod 2^64 */
} field_t;

/* ------------------------------------------------------------------ limbs */

static int fe_is_zero(const fe *a) { return (a->v[0] | a->v[1] | a->v[2] | a->v[3]) == 0; }
static int fe_eq(const fe *a, const fe *b) {
  return a->v[0] == b->v[0] && a->v[1] == b->v[1] && a->v[2] == b->v[2] && a->v[3] == b->v[3];
}
static int fe_geq(const fe *a, const fe *b) {
  for (int i = 3; i >= 0; --i) {
    if (a->v[i] > b->v[i]) return 1;
    if (a->v[i] < b->v[i]) return 0;
  }
  return 1;
}
static uint64_t fe_add_raw(fe *r, const fe *a, const fe *b) {
  u128 c = 0;
  for (int i = 0; i < 4; ++i) { c += (u128)a->v[i] + b->v[i]; r->v[i] = (uint64_t)c; c >>= 64; }
  return (uint64_t)c;
}
static uint64_t fe_sub_raw(fe *r, const fe *a, const fe *b) {
  uint64_t borrow = 0;
  for (int i = 0; i < 4; ++i) {
    u128 d = (u128)a->v[i] - b->v[i] - borrow;
    r->v[i] = (uint64_t)d;
    borrow = (uint64_t)(d >> 64) & 1;
  }
  return borrow;
}
static void fe_load(fe *r, const uint32_t *w) {
  for (int i = 0; i < 4; ++i) r->v[i] = (uint64_t)w[2 * i] | ((uint64_t)w[2 * i + 1] << 32);
}
static void fe_store(uint32_t *w, const fe *a) {
  for (int i = 0; i < 4; ++i) { w[2 * i] = (uint32_t)a->v[i]; w[2 * i + 1] = (uint32_t)(a->v[i] >> 32); }
}

/* ---------------------------------------------------- field (modp.py:31-53) */

/* IntegerModP.__add__ (starks/modp.py:43-45): (a + b) % p */
static void f_add(const field_t *F, fe *r, const fe *a, const fe *b) {
  fe s; uint64_t c = fe_add_raw(&s, a, b);
  if (c || fe_geq(&s, &F->p)) fe_sub_raw(&s, &s, &F->p);
  *r = s;
}
/* IntegerModP.__sub__ (starks/modp.py:47-49): (a - b) % p */
static void f_sub(const field_t *F, fe *r, const fe *a, const fe *b) {
  fe d; uint64_t bw = fe_sub_raw(&d, a, b);
  if (bw) fe_add_raw(&d, &d, &F->p);
  *r = d;
}
/* Montgomery product a*b/2^256 mod p (CIOS); canonical output. */
static void f_mmul(const field_t *F, fe *r, const fe *a, const fe *b) {
  uint64_t t[6] = {0, 0, 0, 0, 0, 0};
  for (int i = 0; i < 4; ++i) {
    u128 c = 0;
    for (int j = 0; j < 4; ++j) {
      c += (u128)a->v[j] * b->v[i] + t[j];
      t[j] = (uint64_t)c; c >>= 64;
    }
    c += t[4]; t[4] = (uint64_t)c; t[5] = (uint64_t)(c >> 64);
    uint64_t m = t[0] * F->ninv;
    c = (u128)m * F->p.v[0] + t[0]; c >>= 64;
    for (int j = 1; j < 4; ++j) {
      c += (u128)m * F->p.v[j] + t[j];
      t[j - 1] = (uint64_t)c; c >>= 64;
    }
    c += t[4]; t[3] = (uint64_t)c; t[4] = t[5] + (uint64_t)(c >> 64);
  }
  fe s = {{t[0], t[1], t[2], t[3]}};
  if (t[4] || fe_geq(&s, &F->p)) fe_sub_raw(&s, &s, &F->p);
  *r = s;
}
static void f_to_mont(const field_t *F, fe *r, const fe *a) { f_mmul(F, r, a, &F->r2); }
static void f_from_mont(const field_t *F, fe *r, const fe *a) {
  fe one = {{1, 0, 0, 0}};
  f_mmul(F, r, a, &one);
}
/* IntegerModP.__mul__ (starks/modp.py:51-53): (a * b) % p on canonical residues. */
static void f_mul(const field_t *F, fe *r, const fe *a, const fe *b) {
  fe am; f_to_mont(F, &am, a);
  f_mmul(F, r, &am, b);
}
static int field_init(field_t *F, const uint32_t p32[8]) {
  fe_load(&F->p, p32);
  if (!(F->p.v[0] & 1)) return -1; /* odd moduli only */
  uint64_t p0 = F->p.v[0], inv = 1;
  for (int i = 0; i < 6; ++i) inv *= 2 - p0 * inv; /* Newton: p0^{-1} mod 2^64 */
  F->ninv = (uint64_t)0 - inv;
  /* 2^512 mod p by 512 modular doublings of 1 (1 < p since p is odd and > 1). */
  fe x = {{1, 0, 0, 0}};
  fe one_c = {{1, 0, 0, 0}};
  if (fe_geq(&one_c, &F->p)) return -1;
  for (int i = 0; i < 512; ++i) {
    f_add(F, &x, &x, &x);
    if (i == 255) F->one = x;
  }
  F->r2 = x;
  return 0;
}
/* Reduce an arbitrary 256-bit value mod p (IntegerModP.__init__, modp.py:35-36). */
static void f_reduce(const field_t *F, fe *r, const fe *a) {
  /* a*R2/R = a*R, then /R again = a mod p */
  fe t; f_mmul(F, &t, a, &F->r2);
  f_from_mont(F, r, &t);
}
/* DomainElement.__pow__ (starks/numbertype.py:68-84): square-and-multiply.
 * Input and output in Montgomery form, exponent a 256-bit integer. */
static void f_mpow(const field_t *F, fe *r, const fe *a_m, const fe *e) {
  fe acc = F->one, base = *a_m;
  for (int i = 0; i < 256; ++i) {
    if ((e->v[i / 64] >> (i % 64)) & 1) f_mmul(F, &acc, &acc, &base);
    f_mmul(F, &base, &base, &base);
  }
  *r = acc;
}
/* IntegerModP.inverse (starks/modp.py:71-79) computes x with a*x = 1 (mod p) by the
 * extended Euclidean algorithm; for prime p the unique such x is a^(p-2). */
static void f_minv(const field_t *F, fe *r, const fe *a_m) {
  fe e, two = {{2, 0, 0, 0}};
  fe_sub_raw(&e, &F->p, &two);
  f_mpow(F, r, a_m, &e);
}

/* --------------------------------------------------------------- exports */

#define EXPORT __attribute__((visibility("default")))

EXPORT int orc_version(void) { return 1; }

EXPORT int orc_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* op: 0 add, 1 sub, 2 mul, 3 inverse(a), 4 reduce(a); n independent pairs. */
EXPORT int orc_field_op(const uint32_t *p, int op, const uint32_t *a, const uint32_t *b,
                        uint32_t *out, uint64_t n) {
  field_t F; if (field_init(&F, p)) return -1;
  for (uint64_t i = 0; i < n; ++i) {
    fe x, y, r; fe_load(&x, a + 8 * i);
    if (b) fe_load(&y, b + 8 * i);
    switch (op) {
      case 0: f_add(&F, &r, &x, &y); break;
      case 1: f_sub(&F, &r, &x, &y); break;
      case 2: f_mul(&F, &r, &x, &y); break;
      case 3: { fe xm; f_to_mont(&F, &xm, &x); f_minv(&F, &xm, &xm); f_from_mont(&F, &r, &xm); } break;
      case 4: f_reduce(&F, &r, &x); break;
      default: return -2;
    }
    fe_store(out + 8 * i, &r);
  }
  return 0;
}

/* a ** e (DomainElement.__pow__, starks/numbertype.py:68-84), e a 256-bit integer. */
EXPORT int orc_pow(const uint32_t *p, const uint32_t *a, const uint32_t *e, uint32_t *out) {
  field_t F; if (field_init(&F, p)) return -1;
  fe x, ee, r; fe_load(&x, a); fe_load(&ee, e);
  f_to_mont(&F, &x, &x); f_mpow(&F, &r, &x, &ee); f_from_mont(&F, &r, &r);
  fe_store(out, &r);
  return 0;
}

/* get_power_cycle (starks/utils.py:30-38): [1, r, r^2, ...] up to but excluding the
 * return to 1.  Writes at most cap elements; returns the cycle length (or -1 if > cap). */
EXPORT int64_t orc_power_cycle(const uint32_t *p, const uint32_t *r32, uint32_t *out, uint64_t cap) {
  field_t F; if (field_init(&F, p)) return -1;
  fe r, rm, cur; fe_load(&r, r32); f_to_mont(&F, &rm, &r);
  cur = F.one;
  uint64_t n = 0;
  for (;;) {
    if (n >= cap) return -1;
    fe c; f_from_mont(&F, &c, &cur); fe_store(out + 8 * n, &c); ++n;
    f_mmul(&F, &cur, &cur, &rm);
    if (fe_eq(&cur, &F.one)) break;
  }
  return (int64_t)n;
}

/* ------------------------------------------------ NTT (starks/fft.py:287-331) */

/* _simple_ft (starks/fft.py:287-300): naive DFT, L = len(roots). */
static void simple_ft(const field_t *F, const fe *vals, size_t vs, const fe *roots, size_t rs,
                      size_t L, fe *out) {
  for (size_t i = 0; i < L; ++i) {
    fe last = {{0, 0, 0, 0}};
    for (size_t j = 0; j < L; ++j) {
      fe t; f_mmul(F, &t, &vals[j * vs], &roots[((i * j) % L) * rs]);
      f_add(F, &last, &last, &t);
    }
    out[i] = last;
  }
}
/* _fft (starks/fft.py:303-314): recursive radix-2 decimation in time.  vals[::2] /
 * roots[::2] become strided views; out and tmp each hold n elements.  vals are plain
 * residues, roots are in Montgomery form, so every product is a plain residue. */
static int fft_rec(const field_t *F, const fe *vals, size_t vs, size_t n, const fe *roots,
                   size_t rs, size_t nroots, fe *out, fe *tmp) {
  if (n <= 4) {
    if (nroots != n) return -3; /* the reference raises IndexError here */
    simple_ft(F, vals, vs, roots, rs, n, out);
    return 0;
  }
  size_t nl = (n + 1) / 2, nr = n / 2, hr = (nroots + 1) / 2;
  int e = fft_rec(F, vals, 2 * vs, nl, roots, 2 * rs, hr, tmp, out);
  if (e) return e;
  e = fft_rec(F, vals + vs, 2 * vs, nr, roots, 2 * rs, hr, tmp + nl, out + nl);
  if (e) return e;
  if (nl != nr) return -3;
  for (size_t i = 0; i < nl; ++i) {
    fe yr; f_mmul(F, &yr, &tmp[nl + i], &roots[i * rs]);
    f_add(F, &out[i], &tmp[i], &yr);
    f_sub(F, &out[i + nl], &tmp[i], &yr);
  }
  return 0;
}

/* Root table of fft_1d (starks/fft.py:319-321): rootz = [1, w, w^2, ... , w^N = 1].
 * Returns N (order of w) or 0 if it exceeds cap.  Entries in Montgomery form. */
static size_t root_table(const field_t *F, const fe *w_m, fe **table, size_t cap) {
  size_t alloc = 1024, n = 2;
  fe *t = (fe *)malloc(alloc * sizeof(fe));
  t[0] = F->one; t[1] = *w_m;
  while (!fe_eq(&t[n - 1], &F->one)) {
    if (n == alloc) { alloc *= 2; t = (fe *)realloc(t, alloc * sizeof(fe)); }
    if (n > cap + 1) { free(t); return 0; }
    f_mmul(F, &t[n], &t[n - 1], w_m);
    ++n;
  }
  *table = t;
  return n - 1;
}

/* fft_1d on one column given a prebuilt root table (starks/fft.py:316-331). */
static int fft_1d_col(const field_t *F, const fe *rootz, size_t N, const uint32_t *in, size_t n_in,
                      uint32_t *out, int inv) {
  if (n_in > N) return -3; /* reference: IndexError */
  fe *vals = (fe *)calloc(N, sizeof(fe)); /* zero padding, fft.py:323-324 */
  fe *o = (fe *)malloc(N * sizeof(fe));
  fe *tmp = (fe *)malloc(N * sizeof(fe));
  fe *roots = (fe *)malloc(N * sizeof(fe));
  for (size_t i = 0; i < n_in; ++i) { fe_load(&vals[i], in + 8 * i); f_reduce(F, &vals[i], &vals[i]); }
  if (inv) for (size_t i = 0; i < N; ++i) roots[i] = rootz[N - i]; /* rootz[:0:-1] */
  else memcpy(roots, rootz, N * sizeof(fe));                       /* rootz[:-1]  */
  int e = fft_rec(F, vals, 1, N, roots, 1, N, o, tmp);
  if (!e) {
    if (inv) {
      /* invlen = pow(len(vals), p-2, p)  (fft.py:327) */
      fe nn = {{(uint64_t)N, 0, 0, 0}}, nm, ninv;
      f_reduce(F, &nn, &nn); f_to_mont(F, &nm, &nn); f_minv(F, &ninv, &nm);
      for (size_t i = 0; i < N; ++i) { fe r; f_mmul(F, &r, &o[i], &ninv); fe_store(out + 8 * i, &r); }
    } else {
      for (size_t i = 0; i < N; ++i) fe_store(out + 8 * i, &o[i]);
    }
  }
  free(vals); free(o); free(tmp); free(roots);
  return e;
}

/* Batched fft_1d: `batch` columns, column c reads in + c*in_stride (n_in elements of
 * 8 limbs) and writes out + c*out_stride (N elements).  N = multiplicative order of
 * root (must be <= cap_n).  Columns run in parallel under OpenMP when threads > 1; the
 * reference itself is single-threaded (one column after another, stark.py:254-256). */
EXPORT int64_t orc_fft(const uint32_t *p, const uint32_t *root, const uint32_t *in, uint64_t n_in,
                       uint64_t in_stride, uint32_t *out, uint64_t out_stride, uint64_t batch,
                       uint64_t cap_n, int inv, int threads) {
  field_t F; if (field_init(&F, p)) return -1;
  fe w, wm; fe_load(&w, root); f_reduce(&F, &w, &w); f_to_mont(&F, &wm, &w);
  fe *rootz = NULL;
  size_t N = root_table(&F, &wm, &rootz, cap_n);
  if (!N) return -2;
  int err = 0;
  (void)threads;
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads > 0 ? threads : 1)
#endif
  for (int64_t c = 0; c < (int64_t)batch; ++c) {
    int e = fft_1d_col(&F, rootz, N, in + c * in_stride, n_in, out + c * out_stride, inv);
    if (e) err = e;
  }
  free(rootz);
  return err ? err : (int64_t)N;
}

/* mul_polys (starks/fft.py:334-345): three _fft calls, NO 1/N scaling. */
EXPORT int64_t orc_mul_polys(const uint32_t *p, const uint32_t *root, const uint32_t *a, uint64_t na,
                             const uint32_t *b, uint64_t nb, uint32_t *out, uint64_t cap_n) {
  field_t F; if (field_init(&F, p)) return -1;
  fe w, wm; fe_load(&w, root); f_reduce(&F, &w, &w); f_to_mont(&F, &wm, &w);
  fe *rootz = NULL;
  size_t N = root_table(&F, &wm, &rootz, cap_n);
  if (!N) return -2;
  if (na > N || nb > N) { free(rootz); return -3; }
  fe *va = (fe *)calloc(N, sizeof(fe)), *vb = (fe *)calloc(N, sizeof(fe));
  fe *x1 = (fe *)malloc(N * sizeof(fe)), *x2 = (fe *)malloc(N * sizeof(fe));
  fe *tmp = (fe *)malloc(N * sizeof(fe)), *rr = (fe *)malloc(N * sizeof(fe));
  for (size_t i = 0; i < na; ++i) { fe_load(&va[i], a + 8 * i); f_reduce(&F, &va[i], &va[i]); }
  for (size_t i = 0; i < nb; ++i) { fe_load(&vb[i], b + 8 * i); f_reduce(&F, &vb[i], &vb[i]); }
  int e = fft_rec(&F, va, 1, N, rootz, 1, N, x1, tmp);
  if (!e) e = fft_rec(&F, vb, 1, N, rootz, 1, N, x2, tmp);
  if (!e) {
    for (size_t i = 0; i < N; ++i) f_mul(&F, &va[i], &x1[i], &x2[i]);
    for (size_t i = 0; i < N; ++i) rr[i] = rootz[N - i];
    e = fft_rec(&F, va, 1, N, rr, 1, N, x1, tmp);
    if (!e) for (size_t i = 0; i < N; ++i) fe_store(out + 8 * i, &x1[i]);
  }
  free(va); free(vb); free(x1); free(x2); free(tmp); free(rr); free(rootz);
  return e ? e : (int64_t)N;
}

/* ---------------------------------------------- BLAKE2s-256 (RFC 7693) */
/* The reference hashes with hashlib.blake2s(x).digest() (starks/merkle_tree.py:1-5):
 * unkeyed, digest length 32, no salt/personalisation. */

static const uint32_t B2S_IV[8] = {0x6A09E667u, 0xBB67AE85u, 0x3C6EF372u, 0xA54FF53Au,
                                   0x510E527Fu, 0x9B05688Cu, 0x1F83D9ABu, 0x5BE0CD19u};
static const uint8_t B2S_SIGMA[10][16] = {
    {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15},
    {14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3},
    {11, 8, 12, 0, 5, 2, 15, 13, 10, 14, 3, 6, 7, 1, 9, 4},
    {7, 9, 3, 1, 13, 12, 11, 14, 2, 6, 5, 10, 4, 0, 15, 8},
    {9, 0, 5, 7, 2, 4, 10, 15, 14, 1, 11, 12, 6, 8, 3, 13},
    {2, 12, 6, 10, 0, 11, 8, 3, 4, 13, 7, 5, 15, 14, 1, 9},
    {12, 5, 1, 15, 14, 13, 4, 10, 0, 7, 6, 3, 9, 2, 8, 11},
    {13, 11, 7, 14, 12, 1, 3, 9, 5, 0, 15, 4, 8, 6, 2, 10},
    {6, 15, 14, 9, 11, 3, 0, 8, 12, 2, 13, 7, 1, 4, 10, 5},
    {10, 2, 8, 4, 7, 6, 1, 5, 15, 11, 9, 14, 3, 12, 13, 0}};

static inline uint32_t rotr32(uint32_t x, int n) { return (x >> n) | (x << (32 - n)); }

static void b2s_compress(uint32_t h[8], const uint8_t block[64], uint64_t t, int last) {
  uint32_t m[16], v[16];
  for (int i = 0; i < 16; ++i)
    m[i] = (uint32_t)block[4 * i] | ((uint32_t)block[4 * i + 1] << 8) |
           ((uint32_t)block[4 * i + 2] << 16) | ((uint32_t)block[4 * i + 3] << 24);
  for (int i = 0; i < 8; ++i) { v[i] = h[i]; v[i + 8] = B2S_IV[i]; }
  v[12] ^= (uint32_t)t; v[13] ^= (uint32_t)(t >> 32);
  if (last) v[14] = ~v[14];
#define G(a, b, c, d, x, y)                                   \
  v[a] = v[a] + v[b] + (x); v[d] = rotr32(v[d] ^ v[a], 16);   \
  v[c] = v[c] + v[d];       v[b] = rotr32(v[b] ^ v[c], 12);   \
  v[a] = v[a] + v[b] + (y); v[d] = rotr32(v[d] ^ v[a], 8);    \
  v[c] = v[c] + v[d];       v[b] = rotr32(v[b] ^ v[c], 7);
  for (int r = 0; r < 10; ++r) {
    const uint8_t *s = B2S_SIGMA[r];
    G(0, 4, 8, 12, m[s[0]], m[s[1]]);  G(1, 5, 9, 13, m[s[2]], m[s[3]]);
    G(2, 6, 10, 14, m[s[4]], m[s[5]]); G(3, 7, 11, 15, m[s[6]], m[s[7]]);
    G(0, 5, 10, 15, m[s[8]], m[s[9]]); G(1, 6, 11, 12, m[s[10]], m[s[11]]);
    G(2, 7, 8, 13, m[s[12]], m[s[13]]); G(3, 4, 9, 14, m[s[14]], m[s[15]]);
  }
#undef G
  for (int i = 0; i < 8; ++i) h[i] ^= v[i] ^ v[i + 8];
}

static void blake2s_256(const uint8_t *in, size_t len, uint8_t out[32]) {
  uint32_t h[8];
  for (int i = 0; i < 8; ++i) h[i] = B2S_IV[i];
  h[0] ^= 0x01010020u; /* digest_length=32, key_length=0, fanout=1, depth=1 */
  uint64_t t = 0;
  while (len > 64) { t += 64; b2s_compress(h, in, t, 0); in += 64; len -= 64; }
  uint8_t last[64]; memset(last, 0, 64); if (len) memcpy(last, in, len);
  t += len; b2s_compress(h, last, t, 1);
  for (int i = 0; i < 8; ++i) {
    out[4 * i] = (uint8_t)h[i]; out[4 * i + 1] = (uint8_t)(h[i] >> 8);
    out[4 * i + 2] = (uint8_t)(h[i] >> 16); out[4 * i + 3] = (uint8_t)(h[i] >> 24);
  }
}

EXPORT int orc_blake2s(const uint8_t *in, uint64_t len, uint8_t *out32) {
  blake2s_256(in, (size_t)len, out32);
  return 0;
}

/* ----------------------------------- Merkle (starks/merkle_tree.py:11-56) */

/* merkelize (starks/merkle_tree.py:36-56) on n raw leaves of leaf_len bytes each
 * (already serialised the way :47-53 does: 32-byte big-endian for ints / field
 * elements, raw for bytes).  Outputs
 *   leaves_perm: the n' = 4*(n//4) leaves after permute4 (:11-23) -> tree[n' .. 2n')
 *   nodes      : 32*n' bytes; nodes[32*i..] = tree[i] for 1 <= i < n' (entry 0 zeroed;
 *                the reference keeps b'' there).
 * Returns n'.  Threads only split independent subtrees of one level. */
EXPORT int64_t orc_merkelize(const uint8_t *leaves, uint64_t n, uint64_t leaf_len,
                             uint8_t *leaves_perm, uint8_t *nodes, int threads) {
  uint64_t ld4 = n / 4, np = 4 * ld4;
  for (uint64_t i = 0; i < ld4; ++i)
    for (uint64_t j = 0; j < 4; ++j)
      memcpy(leaves_perm + (4 * i + j) * leaf_len, leaves + (i + j * ld4) * leaf_len, leaf_len);
  if (np == 0) return 0;
  memset(nodes, 0, 32);
  (void)threads;
  /* nodes[i] = blake(nodes[2i] + nodes[2i+1]) for i = n'-1 .. 1 (:54-55).  Processed in
   * descending index order inside each chunk; chunks of one heap level are independent
   * (children always have larger indices and belong to the level below). */
  uint8_t *buf = (uint8_t *)malloc(2 * (leaf_len > 32 ? leaf_len : 32));
  if (threads <= 1) {
    for (uint64_t i = np - 1; i >= 1; --i) {
      const uint8_t *l, *r; size_t ll, rl;
      uint64_t a = 2 * i, b = 2 * i + 1;
      if (a >= np) { l = leaves_perm + (a - np) * leaf_len; ll = leaf_len; } else { l = nodes + 32 * a; ll = 32; }
      if (b >= np) { r = leaves_perm + (b - np) * leaf_len; rl = leaf_len; } else { r = nodes + 32 * b; rl = 32; }
      memcpy(buf, l, ll); memcpy(buf + ll, r, rl);
      blake2s_256(buf, ll + rl, nodes + 32 * i);
    }
  } else {
    /* level by level, highest indices first: [lo, hi) with lo = 2^k */
    uint64_t top = 1; while (top * 2 <= np - 1) top *= 2; /* top = largest power of two <= np-1 */
    for (uint64_t lo = top; lo >= 1; lo /= 2) {
      uint64_t hi = lo * 2 < np ? lo * 2 : np;
#ifdef _OPENMP
#pragma omp parallel num_threads(threads)
#endif
      {
        uint8_t *tb = (uint8_t *)malloc(2 * (leaf_len > 32 ? leaf_len : 32));
#ifdef _OPENMP
#pragma omp for schedule(static)
#endif
        for (int64_t i = (int64_t)lo; i < (int64_t)hi; ++i) {
          const uint8_t *l, *r; size_t ll, rl;
          uint64_t a = 2 * (uint64_t)i, b = a + 1;
          if (a >= np) { l = leaves_perm + (a - np) * leaf_len; ll = leaf_len; } else { l = nodes + 32 * a; ll = 32; }
          if (b >= np) { r = leaves_perm + (b - np) * leaf_len; rl = leaf_len; } else { r = nodes + 32 * b; rl = 32; }
          memcpy(tb, l, ll); memcpy(tb + ll, r, rl);
          blake2s_256(tb, ll + rl, nodes + 32 * (uint64_t)i);
        }
        free(tb);
      }
      if (lo == 1) break;
    }
  }
  free(buf);
  return (int64_t)np;
}

/* Serialise n field elements (8 LE limbs) as 32-byte big-endian (modp.py:94-95). */
EXPORT int orc_to_bytes_be(const uint32_t *limbs, uint64_t n, uint8_t *out) {
  for (uint64_t i = 0; i < n; ++i)
    for (int k = 0; k < 8; ++k) {
      uint32_t w = limbs[8 * i + (7 - k)];
      out[32 * i + 4 * k] = (uint8_t)(w >> 24); out[32 * i + 4 * k + 1] = (uint8_t)(w >> 16);
      out[32 * i + 4 * k + 2] = (uint8_t)(w >> 8); out[32 * i + 4 * k + 3] = (uint8_t)w;
    }
  return 0;
}

/* Leaves of merkelize_polynomial_evaluations (starks/merkle_tree.py:116-118):
 * leaf_i = b''.join(col[i].to_bytes() for col in cols); cols column-major, stride
 * col_stride limbs-elements between columns. */
EXPORT int orc_pack_leaves(const uint32_t *cols, uint64_t n, uint64_t ncols, uint64_t col_stride,
                           uint8_t *out) {
  for (uint64_t i = 0; i < n; ++i)
    for (uint64_t c = 0; c < ncols; ++c)
      orc_to_bytes_be(cols + 8 * (c * col_stride + i), 1, out + 32 * (i * ncols + c));
  return 0;
}

/* ------------------------------ FRI fold (fri.py:236-242, poly_utils.py:301-320,412-440) */

/* One FRI layer's "column": for i < q = n/4, the degree<4 interpolant through
 * (xs[i+q*j], values[i+q*j]), j<4, evaluated at special_x.  Follows multi_interp_4
 * literally: the four numerator cubics eq0..eq3 (:419-427), their values e0..e3 at
 * their own node, ONE batched inversion over all 4q denominators (multi_inv,
 * poly_utils.py:301-320), coefficient assembly (:436-439), then Horner evaluation
 * (Polynomial.__call__, starks/polynomial.py:158-164).  xs is the power cycle of
 * root (utils.py:30-38).  special_x is an arbitrary 256-bit integer (fri.py:229 does
 * not reduce it); products reduce it implicitly, so it is reduced here first. */
EXPORT int orc_fri_fold(const uint32_t *p, const uint32_t *root, const uint32_t *values, uint64_t n,
                        const uint32_t *special_x, uint32_t *column) {
  field_t F; if (field_init(&F, p)) return -1;
  if (n % 4) return -2;
  uint64_t q = n / 4;
  fe w, wm, sx; fe_load(&w, root); f_reduce(&F, &w, &w); f_to_mont(&F, &wm, &w);
  fe_load(&sx, special_x); f_reduce(&F, &sx, &sx); f_to_mont(&F, &sx, &sx);
  /* everything below in Montgomery form */
  fe *xs = (fe *)malloc(n * sizeof(fe)), *ys = (fe *)malloc(n * sizeof(fe));
  xs[0] = F.one;
  for (uint64_t i = 1; i < n; ++i) f_mmul(&F, &xs[i], &xs[i - 1], &wm);
  for (uint64_t i = 0; i < n; ++i) { fe t; fe_load(&t, values + 8 * i); f_reduce(&F, &t, &t); f_to_mont(&F, &ys[i], &t); }
  fe *eq = (fe *)malloc(q * 16 * sizeof(fe));      /* eq[i][k][c], c = coefficient 0..3 */
  fe *targets = (fe *)malloc(n * sizeof(fe));      /* e0..e3 per row */
  fe zero = {{0, 0, 0, 0}};
  for (uint64_t i = 0; i < q; ++i) {
    fe x[4];
    for (int j = 0; j < 4; ++j) x[j] = xs[i + q * j];
    fe x01, x02, x03, x12, x13, x23;
    f_mmul(&F, &x01, &x[0], &x[1]); f_mmul(&F, &x02, &x[0], &x[2]); f_mmul(&F, &x03, &x[0], &x[3]);
    f_mmul(&F, &x12, &x[1], &x[2]); f_mmul(&F, &x13, &x[1], &x[3]); f_mmul(&F, &x23, &x[2], &x[3]);
    fe *e = eq + 16 * i, t, u;
    /* eq0 = [-x12*x3, x12+x13+x23, -x1-x2-x3, 1] */
    f_mmul(&F, &t, &x12, &x[3]); f_sub(&F, &e[0], &zero, &t);
    f_add(&F, &t, &x12, &x13); f_add(&F, &e[1], &t, &x23);
    f_sub(&F, &t, &zero, &x[1]); f_sub(&F, &t, &t, &x[2]); f_sub(&F, &e[2], &t, &x[3]); e[3] = F.one;
    /* eq1 = [-x02*x3, x02+x03+x23, -x0-x2-x3, 1] */
    f_mmul(&F, &t, &x02, &x[3]); f_sub(&F, &e[4], &zero, &t);
    f_add(&F, &t, &x02, &x03); f_add(&F, &e[5], &t, &x23);
    f_sub(&F, &t, &zero, &x[0]); f_sub(&F, &t, &t, &x[2]); f_sub(&F, &e[6], &t, &x[3]); e[7] = F.one;
    /* eq2 = [-x01*x3, x01+x03+x13, -x0-x1-x3, 1] */
    f_mmul(&F, &t, &x01, &x[3]); f_sub(&F, &e[8], &zero, &t);
    f_add(&F, &t, &x01, &x03); f_add(&F, &e[9], &t, &x13);
    f_sub(&F, &t, &zero, &x[0]); f_sub(&F, &t, &t, &x[1]); f_sub(&F, &e[10], &t, &x[3]); e[11] = F.one;
    /* eq3 = [-x01*x2, x01+x02+x12, -x0-x1-x2, 1] */
    f_mmul(&F, &t, &x01, &x[2]); f_sub(&F, &e[12], &zero, &t);
    f_add(&F, &t, &x01, &x02); f_add(&F, &e[13], &t, &x12);
    f_sub(&F, &t, &zero, &x[0]); f_sub(&F, &t, &t, &x[1]); f_sub(&F, &e[14], &t, &x[2]); e[15] = F.one;
    /* e_k = eq_k(x_k), Horner high -> low */
    for (int k = 0; k < 4; ++k) {
      fe acc = e[4 * k + 3];
      for (int c = 2; c >= 0; --c) { f_mmul(&F, &u, &acc, &x[k]); f_add(&F, &acc, &u, &e[4 * k + c]); }
      targets[4 * i + k] = acc;
    }
  }
  /* multi_inv (poly_utils.py:301-320): zeros are skipped (treated as 1, output 0). */
  fe *partials = (fe *)malloc((n + 1) * sizeof(fe)), *invs = (fe *)malloc(n * sizeof(fe));
  partials[0] = F.one;
  for (uint64_t i = 0; i < n; ++i) {
    if (fe_is_zero(&targets[i])) partials[i + 1] = partials[i];
    else f_mmul(&F, &partials[i + 1], &partials[i], &targets[i]);
  }
  fe inv; f_minv(&F, &inv, &partials[n]);
  for (uint64_t i = n; i > 0; --i) {
    if (fe_is_zero(&targets[i - 1])) invs[i - 1] = zero;
    else { f_mmul(&F, &invs[i - 1], &partials[i - 1], &inv); f_mmul(&F, &inv, &inv, &targets[i - 1]); }
  }
  for (uint64_t i = 0; i < q; ++i) {
    fe *e = eq + 16 * i, invy[4], coef[4], t;
    for (int k = 0; k < 4; ++k) f_mmul(&F, &invy[k], &ys[i + q * k], &invs[4 * i + k]);
    for (int c = 0; c < 4; ++c) {
      coef[c] = zero;
      for (int k = 0; k < 4; ++k) { f_mmul(&F, &t, &e[4 * k + c], &invy[k]); f_add(&F, &coef[c], &coef[c], &t); }
    }
    fe acc = coef[3];
    for (int c = 2; c >= 0; --c) { f_mmul(&F, &t, &acc, &sx); f_add(&F, &acc, &t, &coef[c]); }
    fe r; f_from_mont(&F, &r, &acc); fe_store(column + 8 * i, &r);
  }
  free(xs); free(ys); free(eq); free(targets); free(partials); free(invs);
  return 0;
}

/* ------------------------- dense polynomials (starks/polynomial.py:92-164) */
/* Coefficient vectors low -> high, canonical residues in 8-limb form at the ABI.
 * Used by oracle/oracle.py to restate the O(n^2) coefficient-form steps of
 * STARK.mk_proof (starks/stark.py:38-104) literally. */

/* Polynomial.__mul__ (starks/polynomial.py:110-121): schoolbook product. */
EXPORT int orc_poly_mul(const uint32_t *p, const uint32_t *a, uint64_t na, const uint32_t *b,
                        uint64_t nb, uint32_t *out /* na+nb-1 */) {
  field_t F; if (field_init(&F, p)) return -1;
  if (!na || !nb) return 0;
  fe *am = (fe *)malloc(na * sizeof(fe)), *bv = (fe *)malloc(nb * sizeof(fe));
  fe *r = (fe *)calloc(na + nb - 1, sizeof(fe));
  for (uint64_t i = 0; i < na; ++i) { fe t; fe_load(&t, a + 8 * i); f_to_mont(&F, &am[i], &t); }
  for (uint64_t i = 0; i < nb; ++i) fe_load(&bv[i], b + 8 * i);
  for (uint64_t i = 0; i < na; ++i)
    for (uint64_t j = 0; j < nb; ++j) { fe t; f_mmul(&F, &t, &am[i], &bv[j]); f_add(&F, &r[i + j], &r[i + j], &t); }
  for (uint64_t i = 0; i < na + nb - 1; ++i) fe_store(out + 8 * i, &r[i]);
  free(am); free(bv); free(r);
  return 0;
}

/* Polynomial.__divmod__ (starks/polynomial.py:128-143): schoolbook long division.
 * quo gets na-nb+1 coefficients, rem gets na (high ones zero).  b's leading
 * coefficient b[nb-1] must be non-zero; requires na >= nb. */
EXPORT int orc_poly_divmod(const uint32_t *p, const uint32_t *a, uint64_t na, const uint32_t *b,
                           uint64_t nb, uint32_t *quo, uint32_t *rem) {
  field_t F; if (field_init(&F, p)) return -1;
  if (nb == 0 || na < nb) return -2;
  fe *r = (fe *)malloc(na * sizeof(fe)), *bm = (fe *)malloc(nb * sizeof(fe));
  for (uint64_t i = 0; i < na; ++i) fe_load(&r[i], a + 8 * i);
  for (uint64_t i = 0; i < nb; ++i) { fe t; fe_load(&t, b + 8 * i); f_to_mont(&F, &bm[i], &t); }
  if (fe_is_zero(&bm[nb - 1])) { free(r); free(bm); return -2; }
  fe lead_inv; f_minv(&F, &lead_inv, &bm[nb - 1]); /* Montgomery form */
  for (uint64_t k = na - nb + 1; k-- > 0;) {
    fe qk; f_mmul(&F, &qk, &r[k + nb - 1], &lead_inv); /* plain residue */
    fe_store(quo + 8 * k, &qk);
    if (!fe_is_zero(&qk))
      for (uint64_t j = 0; j < nb; ++j) { fe t; f_mmul(&F, &t, &bm[j], &qk); f_sub(&F, &r[k + j], &r[k + j], &t); }
  }
  for (uint64_t i = 0; i < na; ++i) fe_store(rem + 8 * i, &r[i]);
  free(r); free(bm);
  return 0;
}

/* Polynomial.__call__ on a field element (starks/polynomial.py:158-164): Horner. */
EXPORT int orc_poly_eval(const uint32_t *p, const uint32_t *a, uint64_t na, const uint32_t *x,
                         uint64_t nx, uint32_t *out) {
  field_t F; if (field_init(&F, p)) return -1;
  /* the points are independent: spread them over the host threads when the polynomial is long */
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 1) if (na * nx > (1u << 22))
#endif
  for (int64_t k = 0; k < (int64_t)nx; ++k) {
    fe xv, xm, acc = {{0, 0, 0, 0}};
    fe_load(&xv, x + 8 * k); f_reduce(&F, &xv, &xv); f_to_mont(&F, &xm, &xv);
    for (uint64_t i = na; i-- > 0;) {
      fe c, t; fe_load(&c, a + 8 * i);
      f_mmul(&F, &t, &acc, &xm); f_add(&F, &acc, &t, &c);
    }
    fe_store(out + 8 * k, &acc);
  }
  return 0;
}
