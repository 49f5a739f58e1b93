// Value types and operations used by the generated eval_check kernels (csrc/gen/eval_check_*.cu).
// E4 = one BabyBear degree-4 extension element held in four registers (risc0/core/src/field/baby_bear.rs:375,744-757).
#pragma once
#include "fp.cuh"

namespace r0 {

struct E4 {
  uint32_t a, b, c, d;
};

__device__ __forceinline__ E4 e_add(const E4& x, const E4& y) {
  return E4{fp_add(x.a, y.a), fp_add(x.b, y.b), fp_add(x.c, y.c), fp_add(x.d, y.d)};
}
__device__ __forceinline__ E4 e_sub(const E4& x, const E4& y) {
  return E4{fp_sub(x.a, y.a), fp_sub(x.b, y.b), fp_sub(x.c, y.c), fp_sub(x.d, y.d)};
}
__device__ __forceinline__ E4 e_add_fp(const E4& x, uint32_t f) { return E4{fp_add(x.a, f), x.b, x.c, x.d}; }
__device__ __forceinline__ E4 e_sub_fp(const E4& x, uint32_t f) { return E4{fp_sub(x.a, f), x.b, x.c, x.d}; }
__device__ __forceinline__ E4 e_fp_sub(uint32_t f, const E4& x) {
  return E4{fp_sub(f, x.a), fp_neg(x.b), fp_neg(x.c), fp_neg(x.d)};
}
__device__ __forceinline__ E4 e_scale(const E4& x, uint32_t f) {
  return E4{fp_mul(x.a, f), fp_mul(x.b, f), fp_mul(x.c, f), fp_mul(x.d, f)};
}

// 64-bit sum of up to four products of canonical values (< 4 * P^2 < 2^64), reduced once.
__device__ __forceinline__ uint32_t dot_reduce(uint64_t t) {
  // t < 2^64: bring it under P * 2^32 (mont_reduce's precondition) by conditionally subtracting 2P*2^32 and P*2^32
  uint32_t hi = (uint32_t)(t >> 32);
  uint32_t lo = (uint32_t)t;
  hi = umin32(hi, hi - 2u * P);  // hi < 2^32 < 3P... after this hi < 2P
  hi = umin32(hi, hi - P);
  return mont_reduce(((uint64_t)hi << 32) | lo);
}

// (x0 + x1 X + x2 X^2 + x3 X^3)(y0 + ...) mod X^4 + 11: each output coefficient is one 64-bit dot product; the wrapped
// terms use -11*y (computed once per product: 3 Montgomery products) so that every term is a plain canonical product.
__device__ __forceinline__ E4 e_mul(const E4& x, const E4& y) {
  const uint32_t nb = fp_mul(y.b, FP_NBETA), nc = fp_mul(y.c, FP_NBETA), nd = fp_mul(y.d, FP_NBETA);
  E4 r;
  r.a = dot_reduce((uint64_t)x.a * y.a + (uint64_t)x.b * nd + (uint64_t)x.c * nc + (uint64_t)x.d * nb);
  r.b = dot_reduce((uint64_t)x.a * y.b + (uint64_t)x.b * y.a + (uint64_t)x.c * nd + (uint64_t)x.d * nc);
  r.c = dot_reduce((uint64_t)x.a * y.c + (uint64_t)x.b * y.b + (uint64_t)x.c * y.a + (uint64_t)x.d * nd);
  r.d = dot_reduce((uint64_t)x.a * y.d + (uint64_t)x.b * y.c + (uint64_t)x.c * y.b + (uint64_t)x.d * y.a);
  return r;
}

}  // namespace r0
