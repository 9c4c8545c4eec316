// ntt_pass_stark_hash_t11.cu -- (2048-element tiles; see ntt_pass_stark_hash.cu) final-pass instantiations with the fused Merkle bottom level (HASH in
// ntt.cuh): the last pass of the LDE's forward transform also hashes the leaf pairs of every
// tile as soon as all columns of the tile are stored, so the evaluations are hashed out of L2
// while other CTAs still run butterflies (BLAKE2s is ALU work, the butterflies are bound by the
// FMA pipe) instead of being re-read from HBM by a separate kernel afterwards.
#include <algorithm>
#include "ctx.h"
#include "ntt.cuh"

using namespace stk;

template <int MAXT, int MINB>
static int launch_hash_pass(stk_ctx* c, cudaStream_t s, const NttPass& P) {
  static bool attr_done[64] = {};  // cudaFuncSetAttribute is per device
  auto kern = ntt_pass_kernel<StarkField, 3, MAXT, MINB, false, true>;
  if (!attr_done[c->device & 63]) {
    STK_CUDA(c, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 32 * 8 * MAXT));
    attr_done[c->device & 63] = true;
  }
  const uint32_t T = 1u << P.logT;
  if (T > 8u * MAXT) return stk_fail(c, STK_EUNSUPPORTED, "tile larger than this instantiation");
  if (P.c_is_col || !P.final_pass || !P.grid_swap || P.nrounds < 2 || P.k < 2)
    return stk_fail(c, STK_EUNSUPPORTED, "fused leaf hash needs the final pass of a multi-pass transform");
  const uint64_t tiles = (1ull << P.n) >> P.logT;
  if (tiles > 65535) return stk_fail(c, STK_EUNSUPPORTED, "more than 65535 tiles");
  dim3 grid(P.batch, (unsigned)tiles);
  kern<<<grid, std::max(1u, T >> 3), (size_t)32 * T, s>>>(P, StarkField());
  STK_CUDA(c, cudaGetLastError());
  return STK_OK;
}

int stk_launch_pass_stark_hash_t11(stk_ctx* c, cudaStream_t s, const NttPass& P) {
  return launch_hash_pass<256, 2>(c, s, P);
}
