// blake2s.cuh -- BLAKE2s-256 compression on the ALU pipe (RFC 7693), unkeyed, as used by
// the reference through hashlib.blake2s (starks/merkle_tree.py:1-5).  All sixteen message
// words and the sixteen state words live in registers; the ten rounds are fully unrolled
// so the sigma permutation is resolved at compile time.
#pragma once
#include <stdint.h>

namespace stk {

__device__ __forceinline__ uint32_t b2s_iv(int i) {
  constexpr uint32_t IV[8] = {0x6A09E667u, 0xBB67AE85u, 0x3C6EF372u, 0xA54FF53Au,
                              0x510E527Fu, 0x9B05688Cu, 0x1F83D9ABu, 0x5BE0CD19u};
  return IV[i];
}

__device__ __forceinline__ void b2s_init(uint32_t h[8]) {
#pragma unroll
  for (int i = 0; i < 8; ++i) h[i] = b2s_iv(i);
  h[0] ^= 0x01010020u;  // digest_length = 32, key_length = 0, fanout = depth = 1
}

__device__ __forceinline__ uint32_t rotr32(uint32_t x, int n) { return __funnelshift_r(x, x, n); }

// Pipe balance.  A G function is 4 xors + 4 rotates (ALU pipe only) and 6 additions.  ptxas turns
// a + b + m into one IADD3 -- on the ALU pipe as well -- which leaves the kernel ALU-bound (840 of
// its ~1040 instructions per compression) with the FMA pipe mostly idle.  The message addition is
// therefore written as a multiply-add by an opaque 1 (a constant-bank word ptxas cannot fold): an
// IMAD on the FMA pipe.  ALU work per compression drops from 840 to 680 instructions.
static __constant__ uint32_t stk_b2s_one = 1u;
__device__ __forceinline__ uint32_t b2s_add_fma(uint32_t a, uint32_t x) {
  uint32_t r;
  asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(x), "r"(stk_b2s_one), "r"(a));
  return r;
}

#define STK_B2S_G(a, b, c, d, x, y)    \
  a = b2s_add_fma(a + b, (x)); d = rotr32(d ^ a, 16); \
  c = c + d;       b = rotr32(b ^ c, 12); \
  a = b2s_add_fma(a + b, (y)); d = rotr32(d ^ a, 8);  \
  c = c + d;       b = rotr32(b ^ c, 7);

// h <- F(h, m, t, last)
__device__ __forceinline__ void b2s_compress(uint32_t h[8], const uint32_t m[16], uint32_t t_lo, bool last) {
  constexpr uint8_t S[10][16] = {
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
  uint32_t v0 = h[0], v1 = h[1], v2 = h[2], v3 = h[3], v4 = h[4], v5 = h[5], v6 = h[6], v7 = h[7];
  uint32_t v8 = b2s_iv(0), v9 = b2s_iv(1), v10 = b2s_iv(2), v11 = b2s_iv(3);
  uint32_t v12 = b2s_iv(4) ^ t_lo, v13 = b2s_iv(5);  // messages here are < 4 GiB: t_hi = 0
  uint32_t v14 = last ? ~b2s_iv(6) : b2s_iv(6), v15 = b2s_iv(7);
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    STK_B2S_G(v0, v4, v8, v12, m[S[r][0]], m[S[r][1]]);
    STK_B2S_G(v1, v5, v9, v13, m[S[r][2]], m[S[r][3]]);
    STK_B2S_G(v2, v6, v10, v14, m[S[r][4]], m[S[r][5]]);
    STK_B2S_G(v3, v7, v11, v15, m[S[r][6]], m[S[r][7]]);
    STK_B2S_G(v0, v5, v10, v15, m[S[r][8]], m[S[r][9]]);
    STK_B2S_G(v1, v6, v11, v12, m[S[r][10]], m[S[r][11]]);
    STK_B2S_G(v2, v7, v8, v13, m[S[r][12]], m[S[r][13]]);
    STK_B2S_G(v3, v4, v9, v14, m[S[r][14]], m[S[r][15]]);
  }
  h[0] ^= v0 ^ v8;  h[1] ^= v1 ^ v9;  h[2] ^= v2 ^ v10; h[3] ^= v3 ^ v11;
  h[4] ^= v4 ^ v12; h[5] ^= v5 ^ v13; h[6] ^= v6 ^ v14; h[7] ^= v7 ^ v15;
}

// Message words of a field element's 32-byte big-endian serialisation
// (IntegerModP.to_bytes, starks/modp.py:94-95): word k (little-endian load of bytes
// 4k..4k+3) is the byte-swapped limb 7-k.
__device__ __forceinline__ uint32_t bswap32(uint32_t x) { return __byte_perm(x, 0, 0x0123); }

}  // namespace stk
