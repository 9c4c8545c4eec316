// fri.cu -- FRI fold-by-4 (the "column" of one FRI layer).  Replaces the per-row Lagrange
// interpolation of SmoothSubgroupFRI.generate_proximity_proof (starks/fri.py:236-242) via
// multi_interp_4 / multi_inv (starks/poly_utils.py:301-320, 412-440) and Polynomial.__call__
// (starks/polynomial.py:158-164) by its closed form:
//
//   q = n/4, iota = w^q (primitive 4th root), t = x * w^(-i)
//   column[i] = 1/4 * sum_k t^k * sum_j values[i + q*j] * iota^(-jk)
//
// i.e. a 4-point inverse DFT in the variable X / w^i (multiplications by iota only) followed
// by Horner at t.  Exact arithmetic mod p, hence bit-identical to the interpolation route.
#include "ctx.h"

using namespace stk;

namespace {

template <class F>
__global__ void __launch_bounds__(256) fri_fold4_kernel(const fe* __restrict__ vals, uint64_t q,
                                                        const fe* __restrict__ Winv, uint64_t wstride, fe x_plain, fe iota_tw,
                                                        fe quarter_tw, fe* __restrict__ out, const F f) {
  const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i >= q) return;
  fe v0 = fe_load(vals + i), v1 = fe_load(vals + i + q), v2 = fe_load(vals + i + 2 * q), v3 = fe_load(vals + i + 3 * q);
  fe s02 = f.add(v0, v2), d02 = f.sub(v0, v2), s13 = f.add(v1, v3), d13 = f.sub(v1, v3);
  fe m = f.mul_tw(d13, iota_tw);
  fe c0 = f.add(s02, s13), c2 = f.sub(s02, s13);  // 4*b0, 4*b2
  fe c1 = f.sub(d02, m), c3 = f.add(d02, m);      // 4*b1, 4*b3
  fe t = f.mul_tw(x_plain, fe_load_ro(Winv + i * wstride));  // x * w^-i, plain
  fe t_tw = f.to_tw(t);
  fe r = f.add(f.mul_tw(c3, t_tw), c2);
  r = f.add(f.mul_tw(r, t_tw), c1);
  r = f.add(f.mul_tw(r, t_tw), c0);
  fe_store(out + i, f.mul_tw(r, quarter_tw));
}

}  // namespace

extern "C" __attribute__((visibility("default"))) int stk_fri_fold4(stk_ctx* c, const uint32_t* d_vals, uint64_t n,
                                                                     const uint32_t root[8], const uint32_t special_x[8],
                                                                     uint32_t* d_out) {
  if (!c || !d_vals || !d_out || !root || !special_x) return STK_EINVAL;
  if (n == 0 || (n & 3)) return stk_fail(c, STK_EINVAL, "fold-by-4 needs a length divisible by 4");
  const uint64_t q = n / 4;
  fe w = stk_load_fe(root);
  fe one = host::reduce(host::from_u64(1), c->p);
  if (!fe_eq(stk_h_pow(c, w, n), one) || fe_eq(stk_h_pow(c, w, n / 2), one))
    return stk_fail(c, STK_EINVAL, "root is not a primitive n-th root of unity");
  fe winv = stk_h_inv(c, w);
  const fe* Winv;
  uint64_t wstride = 1;
  STK_TRY(stk_get_table_strided(c, winv, n, &Winv, &wstride));
  fe x = host::reduce(stk_load_fe(special_x), c->p);  // fri.py:229 does not reduce; products do
  fe iota_tw = stk_h_to_tw(c, stk_h_pow(c, w, q));
  fe quarter_tw = stk_h_to_tw(c, stk_h_inv(c, host::reduce(host::from_u64(4), c->p)));
  unsigned blocks = (unsigned)((q + 255) / 256);
  if (c->is_stark)
    fri_fold4_kernel<StarkField><<<blocks, 256, 0, c->stream>>>((const fe*)d_vals, q, Winv, wstride, x, iota_tw, quarter_tw,
                                                                (fe*)d_out, StarkField());
  else
    fri_fold4_kernel<MontField><<<blocks, 256, 0, c->stream>>>((const fe*)d_vals, q, Winv, wstride, x, iota_tw, quarter_tw,
                                                               (fe*)d_out, c->mont);
  STK_CUDA(c, cudaGetLastError());
  return STK_OK;
}
