// hostblake2s.h -- unkeyed BLAKE2s-256 on the host (RFC 7693) for the Fiat-Shamir index
// derivation inside the FRI driver (get_pseudorandom_indices, starks/utils.py:60-90: a few
// hundred bytes per layer).  Bulk hashing happens on the device (blake2s.cuh).
#pragma once
#include <stdint.h>
#include <string.h>

namespace stk {
namespace host {

inline uint32_t b2_rotr(uint32_t x, int n) { return (x >> n) | (x << (32 - n)); }

inline void blake2s_256(const uint8_t* in, size_t len, uint8_t out[32]) {
  static const uint32_t IV[8] = {0x6A09E667u, 0xBB67AE85u, 0x3C6EF372u, 0xA54FF53Au,
                                 0x510E527Fu, 0x9B05688Cu, 0x1F83D9ABu, 0x5BE0CD19u};
  static const uint8_t S[10][16] = {
      {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15}, {14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3},
      {11, 8, 12, 0, 5, 2, 15, 13, 10, 14, 3, 6, 7, 1, 9, 4}, {7, 9, 3, 1, 13, 12, 11, 14, 2, 6, 5, 10, 4, 0, 15, 8},
      {9, 0, 5, 7, 2, 4, 10, 15, 14, 1, 11, 12, 6, 8, 3, 13}, {2, 12, 6, 10, 0, 11, 8, 3, 4, 13, 7, 5, 15, 14, 1, 9},
      {12, 5, 1, 15, 14, 13, 4, 10, 0, 7, 6, 3, 9, 2, 8, 11}, {13, 11, 7, 14, 12, 1, 3, 9, 5, 0, 15, 4, 8, 6, 2, 10},
      {6, 15, 14, 9, 11, 3, 0, 8, 12, 2, 13, 7, 1, 4, 10, 5}, {10, 2, 8, 4, 7, 6, 1, 5, 15, 11, 9, 14, 3, 12, 13, 0}};
  uint32_t h[8];
  for (int i = 0; i < 8; ++i) h[i] = IV[i];
  h[0] ^= 0x01010020u;
  size_t off = 0;
  bool done = false;
  while (!done) {
    uint8_t blk[64];
    memset(blk, 0, 64);
    size_t take = len - off < 64 ? len - off : 64;
    if (take) memcpy(blk, in + off, take);
    off += take;
    done = off == len;  // the last (possibly empty or full) block carries the final flag
    uint32_t m[16], v[16];
    for (int i = 0; i < 16; ++i)
      m[i] = (uint32_t)blk[4 * i] | ((uint32_t)blk[4 * i + 1] << 8) | ((uint32_t)blk[4 * i + 2] << 16) |
             ((uint32_t)blk[4 * i + 3] << 24);
    for (int i = 0; i < 8; ++i) { v[i] = h[i]; v[8 + i] = IV[i]; }
    v[12] ^= (uint32_t)off;
    v[13] ^= (uint32_t)((uint64_t)off >> 32);
    if (done) v[14] = ~v[14];
#define STK_HG(a, b, c, d, x, y)                                  \
  v[a] = v[a] + v[b] + (x); v[d] = b2_rotr(v[d] ^ v[a], 16);      \
  v[c] = v[c] + v[d];       v[b] = b2_rotr(v[b] ^ v[c], 12);      \
  v[a] = v[a] + v[b] + (y); v[d] = b2_rotr(v[d] ^ v[a], 8);       \
  v[c] = v[c] + v[d];       v[b] = b2_rotr(v[b] ^ v[c], 7);
    for (int r = 0; r < 10; ++r) {
      STK_HG(0, 4, 8, 12, m[S[r][0]], m[S[r][1]]);
      STK_HG(1, 5, 9, 13, m[S[r][2]], m[S[r][3]]);
      STK_HG(2, 6, 10, 14, m[S[r][4]], m[S[r][5]]);
      STK_HG(3, 7, 11, 15, m[S[r][6]], m[S[r][7]]);
      STK_HG(0, 5, 10, 15, m[S[r][8]], m[S[r][9]]);
      STK_HG(1, 6, 11, 12, m[S[r][10]], m[S[r][11]]);
      STK_HG(2, 7, 8, 13, m[S[r][12]], m[S[r][13]]);
      STK_HG(3, 4, 9, 14, m[S[r][14]], m[S[r][15]]);
    }
#undef STK_HG
    for (int i = 0; i < 8; ++i) h[i] ^= v[i] ^ v[8 + i];
  }
  for (int i = 0; i < 8; ++i) {
    out[4 * i] = (uint8_t)h[i]; out[4 * i + 1] = (uint8_t)(h[i] >> 8);
    out[4 * i + 2] = (uint8_t)(h[i] >> 16); out[4 * i + 3] = (uint8_t)(h[i] >> 24);
  }
}

}  // namespace host
}  // namespace stk
