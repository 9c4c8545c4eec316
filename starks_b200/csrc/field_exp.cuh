// field_exp.cuh -- EXPERIMENTAL multiply for the STARK prime, measured against the production
// StarkField::mul_tw by the K0 microbenchmarks (stk_microbench_variant); not used by any
// product kernel unless a measurement says it should be.
//
// What the SASS of the production multiply shows (cuobjdump of field_kernel<StarkField,5>, per
// multiply): 71 IMAD.WIDE.U32(.X) + ~2 IMAD/IMAD.HI pairs, and ~22 further instructions that
// ptxas ALSO places on the FMA pipe (15 IMAD.MOV -- half of them zero-initialisations of
// accumulator limbs, half register-pair realignments of the 351*H row that starts on an odd
// limb --, 3.5 IMAD.X, 1.75 IMAD.IADD), against ~67 IADD3 on the ALU pipe.  At 4 clk per wide
// multiply and 2 clk per other FMA-pipe instruction that is ~340 of the measured 368 clk per
// warp multiply: the multiply is FMA-pipe bound with the ALU pipe about one third busy.  This
// variant moves the avoidable FMA-pipe work to the ALU pipe:
//   * no zero-initialised accumulators: a row's fresh top pair takes a literal-zero addend,
//     and the zero above a carry limb is produced by an add-with-carry whose carry is provably
//     clear (an ALU instruction the compiler cannot fold into a move);
//   * 351*H as two aligned rows (even limbs / odd limbs) merged by add chains, so no row starts
//     on an odd limb and no register pair has to be realigned.
#pragma once
#include "field.cuh"

namespace stk {

#ifdef __CUDA_ARCH__
// acc[0..7] += a[0,2,4,6]*b, carry -> acc[8], runtime zero -> acc[9]
__device__ __forceinline__ void mad_row_c(uint32_t* acc, const uint32_t* a, uint32_t b) {
  asm volatile("mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;"
      : "+&r"(acc[0]), "+r"(acc[1]) : "r"(a[0]), "r"(b));
#pragma unroll
  for (int j = 2; j < 8; j += 2)
    asm volatile("madc.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;"
        : "+&r"(acc[j]), "+r"(acc[j + 1]) : "r"(a[j]), "r"(b));
  asm volatile("addc.cc.u32 %0, 0, 0; addc.u32 %1, 0, 0;" : "=r"(acc[8]), "=r"(acc[9]));
}
// same, carry limb only (last odd-column row: limb 16 does not exist)
__device__ __forceinline__ void mad_row_c1(uint32_t* acc, const uint32_t* a, uint32_t b) {
  asm volatile("mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;"
      : "+&r"(acc[0]), "+r"(acc[1]) : "r"(a[0]), "r"(b));
#pragma unroll
  for (int j = 2; j < 8; j += 2)
    asm volatile("madc.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;"
        : "+&r"(acc[j]), "+r"(acc[j + 1]) : "r"(a[j]), "r"(b));
  asm volatile("addc.u32 %0, 0, 0;" : "=r"(acc[8]));
}
// acc[0..7] += a[0,2,4,6]*b where acc[7] == 0 on entry: the sum cannot carry out
__device__ __forceinline__ void mad_row_nc(uint32_t* acc, const uint32_t* a, uint32_t b) {
  asm volatile("mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;"
      : "+&r"(acc[0]), "+r"(acc[1]) : "r"(a[0]), "r"(b));
  asm volatile("madc.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;"
      : "+&r"(acc[2]), "+r"(acc[3]) : "r"(a[2]), "r"(b));
  asm volatile("madc.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;"
      : "+&r"(acc[4]), "+r"(acc[5]) : "r"(a[4]), "r"(b));
  asm volatile("madc.lo.cc.u32 %0, %2, %3, %0; madc.hi.u32 %1, %2, %3, %1;"
      : "+&r"(acc[6]), "+r"(acc[7]) : "r"(a[6]), "r"(b));
}
// acc[0..5] += a[0,2,4]*b and the FRESH pair acc[6..7] = a[6]*b + carry: cannot carry out
__device__ __forceinline__ void mad_row_fresh(uint32_t* acc, const uint32_t* a, uint32_t b) {
  asm volatile("mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;"
      : "+&r"(acc[0]), "+r"(acc[1]) : "r"(a[0]), "r"(b));
  asm volatile("madc.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;"
      : "+&r"(acc[2]), "+r"(acc[3]) : "r"(a[2]), "r"(b));
  asm volatile("madc.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;"
      : "+&r"(acc[4]), "+r"(acc[5]) : "r"(a[4]), "r"(b));
  asm volatile("madc.lo.cc.u32 %0, %2, %3, 0; madc.hi.u32 %1, %2, %3, 0;"
      : "=&r"(acc[6]), "=r"(acc[7]) : "r"(a[6]), "r"(b));
}

// Same product as mul512 (field.cuh).  E holds the columns with even limb index sum, O (one limb
// up) the odd ones.  Row by row, which limbs are live and why no carry is lost:
//   E: j=0 sets 0..7 | j=1 adds 2..7, sets 8..9 (fresh: no carry out) | j=2 adds 2..9, carry
//      -> 10, zero -> 11 | j=3 adds 4..11 (limb 11 was zero: no carry out) | j=4 adds 4..11,
//      carry -> 12, zero -> 13 | j=5 adds 6..13 | j=6 adds 6..13, carry -> 14, zero -> 15 |
//      j=7 adds 8..15.
//   O: j=0 sets 0..7 | j=1 adds 0..7, carry -> 8, zero -> 9 | j=2 adds 2..9 | j=3 adds 2..9,
//      carry -> 10, zero -> 11 | j=4 adds 4..11 | j=5 adds 4..11, carry -> 12, zero -> 13 |
//      j=6 adds 6..13 | j=7 adds 6..13, carry -> 14.
// A pair (carry, 0) plus a 64-bit product plus a carry-in is below 2^64, hence "no carry out".
__device__ __forceinline__ void mul512_x(uint32_t* T, const uint32_t* a, const uint32_t* b) {
  uint32_t E[16], O[16];
  mul_row_first(E, a, b[0]);
  mul_row_first(O, a + 1, b[0]);
  mad_row_fresh(E + 2, a + 1, b[1]);  mad_row_c(O, a, b[1]);
  mad_row_c(E + 2, a, b[2]);          mad_row_nc(O + 2, a + 1, b[2]);
  mad_row_nc(E + 4, a + 1, b[3]);     mad_row_c(O + 2, a, b[3]);
  mad_row_c(E + 4, a, b[4]);          mad_row_nc(O + 4, a + 1, b[4]);
  mad_row_nc(E + 6, a + 1, b[5]);     mad_row_c(O + 4, a, b[5]);
  mad_row_c(E + 6, a, b[6]);          mad_row_nc(O + 6, a + 1, b[6]);
  mad_row_nc(E + 8, a + 1, b[7]);     mad_row_c1(O + 6, a, b[7]);
  T[0] = E[0];
  asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(T[1]) : "r"(E[1]), "r"(O[0]));
#pragma unroll
  for (int i = 2; i < 15; ++i) asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(T[i]) : "r"(E[i]), "r"(O[i - 1]));
  asm volatile("addc.u32 %0, %1, %2;" : "=r"(T[15]) : "r"(E[15]), "r"(O[14]));
}

// T (16 limbs) mod p, canonical: StarkField::reduce512 with the 351*H rows aligned.
__device__ __forceinline__ fe reduce512_x(const uint32_t* T) {
  const uint32_t* L = T;
  const uint32_t* H = T + 8;
  uint32_t UE[8], UO[8];
  mul_row_first(UE, H, 351u);       // limbs 0..7 of 351*H (even limbs of H)
  mul_row_first(UO, H + 1, 351u);   // limbs 1..8 (odd limbs of H)
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
  asm volatile("sub.cc.u32 %0, %0, %1;" : "+r"(B[0]) : "r"(H[0]));
#pragma unroll
  for (int i = 1; i < 8; ++i) asm volatile("subc.cc.u32 %0, %0, %1;" : "+r"(B[i]) : "r"(H[i]));
  asm volatile("subc.cc.u32 %0, %0, 0;" : "+r"(B[8]));
  asm volatile("subc.u32 %0, %0, 0;" : "+r"(B[9]));
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
  asm volatile("subc.u32 %0, 0, 0;" : "=r"(k2));
  uint32_t m = 0u - (k1 + k2);
  asm volatile("add.cc.u32 %0, %0, %1;" : "+r"(B[0]) : "r"(m));
  asm volatile("addc.cc.u32 %0, %0, %1;" : "+r"(B[1]) : "r"(m & 350u));
  asm volatile("addc.u32 %0, %0, 0;" : "+r"(B[2]));
  StarkField::canon(B);
  fe r;
#pragma unroll
  for (int i = 0; i < 8; ++i) r.v[i] = B[i];
  return r;
}
#endif

// VARIANT bit 0: mul512_x, bit 1: reduce512_x
template <int VARIANT>
struct StarkFieldX : StarkField {
  __device__ __forceinline__ fe mul_tw(const fe& a, const fe& t) const {
#ifdef __CUDA_ARCH__
    uint32_t T[16];
    if (VARIANT & 1) mul512_x(T, a.v, t.v); else mul512(T, a.v, t.v);
    if (VARIANT & 2) return reduce512_x(T);
    return reduce512(T);
#else
    return StarkField::mul_tw(a, t);
#endif
  }
};

}  // namespace stk
