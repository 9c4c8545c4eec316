/* Plain-C client of libstarks_b200.so: no Python, no CUDA headers -- only include/starks_b200.h.
 * Forward + inverse NTT of 4 columns of 2^16 elements and a Merkle commitment through the C ABI.
 *   gcc -O2 -Iinclude examples/abi_roundtrip.c -o /tmp/abi_roundtrip -Lstarks_b200 -lstarks_b200 \
 *       -Wl,-rpath,$PWD/starks_b200
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "starks_b200.h"

#define CHECK(call)                                                              \
  do {                                                                           \
    int rc_ = (call);                                                            \
    if (rc_ != STK_OK) {                                                         \
      fprintf(stderr, "%s -> %d: %s\n", #call, rc_, stk_last_error(ctx));        \
      return 1;                                                                  \
    }                                                                            \
  } while (0)

/* w = 7^((p-1)/2^16) mod p for p = 2^256 - 351*2^32 + 1, little-endian limbs
 * (value checked against tests/golden/fft.json, logn = 16). */
static const uint32_t W16[8] = {0x982b76ccu, 0x922259c0u, 0x2e5a87d6u, 0x84967c1eu,
                                0x6e066724u, 0x853f7c61u, 0x6ecc9d72u, 0x5caa5220u};

int main(void) {
  stk_ctx* ctx = NULL;
  if (stk_init(0, &ctx) != STK_OK) {
    fprintf(stderr, "stk_init failed: a CUDA device is required (no CPU fallback)\n");
    return 2;
  }
  const uint64_t n = 1u << 16, cols = 4, words = cols * n * 8;
  uint32_t* h = (uint32_t*)malloc(words * 4);
  uint32_t* back = (uint32_t*)malloc(words * 4);
  uint64_t s = 88172645463325252ull;
  for (uint64_t i = 0; i < words; ++i) {
    s ^= s << 13; s ^= s >> 7; s ^= s << 17;
    h[i] = (uint32_t)s;
    if ((i & 7) == 7) h[i] &= 0x7fffffffu; /* canonical residues */
  }
  void *d_in, *d_ev, *d_back, *d_nodes;
  CHECK(stk_dev_alloc(ctx, words * 4, &d_in));
  CHECK(stk_dev_alloc(ctx, words * 4, &d_ev));
  CHECK(stk_dev_alloc(ctx, words * 4, &d_back));
  CHECK(stk_dev_alloc(ctx, 32 * n, &d_nodes));
  CHECK(stk_memcpy_h2d(ctx, d_in, h, words * 4));
  CHECK(stk_ntt(ctx, (const uint32_t*)d_in, n, n, (uint32_t*)d_ev, n, n, cols, W16, 0));
  uint8_t root[32];
  CHECK(stk_merkle_commit(ctx, (const uint32_t*)d_ev, n, cols, n, (uint8_t*)d_nodes, root));
  CHECK(stk_ntt(ctx, (const uint32_t*)d_ev, n, n, (uint32_t*)d_back, n, n, cols, W16, 1));
  CHECK(stk_memcpy_d2h(ctx, back, d_back, words * 4));
  if (memcmp(h, back, words * 4) != 0) {
    fprintf(stderr, "inverse(forward(x)) != x\n");
    return 1;
  }
  /* an over-long input is an index error, as in the reference */
  if (stk_ntt(ctx, (const uint32_t*)d_in, n + 1, n + 1, (uint32_t*)d_ev, n, n, 1, W16, 0) != STK_EINDEX) return 1;
  printf("abi roundtrip ok, root=");
  for (int i = 0; i < 32; ++i) printf("%02x", root[i]);
  printf("\n");
  stk_dev_free(ctx, d_in); stk_dev_free(ctx, d_ev); stk_dev_free(ctx, d_back); stk_dev_free(ctx, d_nodes);
  stk_destroy(ctx);
  free(h); free(back);
  return 0;
}
