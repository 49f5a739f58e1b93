// BabyBear (P = 15*2^27 + 1) and its degree-4 extension, Montgomery form (R = 2^32), for sm_100a device code and
// the host-side driver. Every stored value is canonical (< P) so results are bit-identical with the reference
// (risc0/core/src/field/baby_bear.rs:324-361 add/sub/mul, :744-757 ExtElem mul, :448-487 ExtElem inv).
//
// Montgomery product: t = a*b (IMAD.WIDE.U32), m = lo(t) * (-P^-1 mod 2^32) (IMAD), u = t + m*P (IMAD.WIDE.U32 with
// 64-bit accumulate), result = hi(u) in [0, 2P) then one min-trick conditional subtract (IADD3 + IMNMX.U32).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define R0_HD __host__ __device__ __forceinline__
#else
#define R0_HD inline
#endif

namespace r0 {

constexpr uint32_t P = 0x78000001u;
constexpr uint32_t MONT_NINV = 0x77FFFFFFu;  // -P^-1 mod 2^32
constexpr uint32_t MONT_R2 = 1172168163u;    // 2^64 mod P
constexpr uint32_t MONT_ONE = 0x0FFFFFFEu;   // 2^32 mod P
constexpr uint32_t FP_INVALID = 0xFFFFFFFFu;

R0_HD uint32_t umin32(uint32_t a, uint32_t b) { return a < b ? a : b; }

// Note on pipes (sm_100a, ncu in profiles/): IMAD-class instructions issue on the fma pipe, IADD3 / VIADDMNMX / LOP3 on
// the alu pipe. Kernels whose mix is alu-heavy (Poseidon2's linear layers) move some plain adds to the fma pipe by
// writing them as x * 1 + y with an opaque 1 (poseidon2.cu); doing that for EVERY add over-loads the fma pipe and is
// slower (measured: hash_rows 2.01 -> 1.68 Gperm/s), so the generic helpers below stay plain.
#define R0_ADD32(a, b) ((a) + (b))
#define R0_SUB32(a, b) ((a) - (b))

R0_HD uint32_t fp_add(uint32_t a, uint32_t b) {
  uint32_t r = R0_ADD32(a, b);
  return umin32(r, r - P);
}
R0_HD uint32_t fp_sub(uint32_t a, uint32_t b) {
  uint32_t r = R0_SUB32(a, b);
  return umin32(r, r + P);
}
R0_HD uint32_t fp_neg(uint32_t a) { return fp_sub(0u, a); }
// Montgomery reduction of t < 2^32 * P : returns t / 2^32 mod P, canonical.
// Subtractive form: m = lo(t) * P^-1, so lo(m * P) == lo(t) and t - m*P = 2^32 * (hi(t) - hi(m*P)) exactly - no carry
// to propagate. hi(t) < P and hi(m*P) < P, so the difference is in (-P, P); the wrapped-unsigned "min(r, r + P)" maps it
// to [0, P). On sm_100a this is IMAD, IMAD.HI.U32, IADD3, VIADDMNMX.U32 (the additive form t + (-m)*P needs the
// carry out of the low word on top).
constexpr uint32_t MONT_PINV = 0x88000001u;  // P^-1 mod 2^32
R0_HD uint32_t mont_reduce(uint64_t t) {
  const uint32_t m = (uint32_t)t * MONT_PINV;
#if defined(__CUDA_ARCH__)
  const uint32_t h = __umulhi(m, P);
#else
  const uint32_t h = (uint32_t)(((uint64_t)m * P) >> 32);
#endif
  const uint32_t r = R0_SUB32((uint32_t)(t >> 32), h);
  return umin32(r, r + P);
}
R0_HD uint32_t fp_mul(uint32_t a, uint32_t b) { return mont_reduce((uint64_t)a * b); }
R0_HD uint32_t fp_encode(uint32_t x) { return fp_mul(MONT_R2, x % P); }
R0_HD uint32_t fp_decode(uint32_t m) { return fp_mul(1u, m); }
R0_HD uint32_t fp_pow(uint32_t x, uint64_t n) {
  uint32_t tot = MONT_ONE;
  while (n) {
    if (n & 1) tot = fp_mul(tot, x);
    n >>= 1;
    x = fp_mul(x, x);
  }
  return tot;
}
// x^(P-2), P - 2 = 0x77FFFFFF = 111 0 111 followed by 24 ones: a fixed chain through x^7 (windows of three bits) -
// 30 squarings + 11 products instead of the 31 + 29 of square-and-multiply. The witness accumulator inverts an
// extension element (one base inversion each) some fifty times per cycle, so this chain is most of its arithmetic.
R0_HD uint32_t fp_inv(uint32_t x) {
  const uint32_t x2 = fp_mul(x, x);
  const uint32_t x3 = fp_mul(x2, x);
  const uint32_t x6 = fp_mul(x3, x3);
  const uint32_t x7 = fp_mul(x6, x);
  uint32_t acc = x7;                       // exponent 0b111
  for (int i = 0; i < 4; i++) acc = fp_mul(acc, acc);
  acc = fp_mul(acc, x7);                   // 0b1110111 = 0x77
  for (int g = 0; g < 8; g++) {            // append 24 ones, three at a time
    acc = fp_mul(acc, acc);
    acc = fp_mul(acc, acc);
    acc = fp_mul(acc, acc);
    acc = fp_mul(acc, x7);
  }
  return acc;
}

// Montgomery constants of small integers (x * 2^32 mod P), usable in constant expressions
constexpr uint32_t mont_const(uint64_t x) { return (uint32_t)(((x % 0x78000001ull) << 32) % 0x78000001ull); }
constexpr uint32_t FP_BETA = mont_const(11);
constexpr uint32_t FP_NBETA = mont_const(0x78000001ull - 11);
constexpr uint32_t FP_THREE = mont_const(3);

struct alignas(16) FpExt {
  uint32_t c[4];
};

R0_HD FpExt ext_zero() { return FpExt{{0, 0, 0, 0}}; }
R0_HD FpExt ext_one() { return FpExt{{MONT_ONE, 0, 0, 0}}; }
R0_HD FpExt ext_from_fp(uint32_t a) { return FpExt{{a, 0, 0, 0}}; }
R0_HD bool ext_eq(const FpExt& a, const FpExt& b) {
  return a.c[0] == b.c[0] && a.c[1] == b.c[1] && a.c[2] == b.c[2] && a.c[3] == b.c[3];
}
R0_HD FpExt ext_add(const FpExt& a, const FpExt& b) {
  return FpExt{{fp_add(a.c[0], b.c[0]), fp_add(a.c[1], b.c[1]), fp_add(a.c[2], b.c[2]), fp_add(a.c[3], b.c[3])}};
}
R0_HD FpExt ext_sub(const FpExt& a, const FpExt& b) {
  return FpExt{{fp_sub(a.c[0], b.c[0]), fp_sub(a.c[1], b.c[1]), fp_sub(a.c[2], b.c[2]), fp_sub(a.c[3], b.c[3])}};
}
R0_HD FpExt ext_neg(const FpExt& a) { return FpExt{{fp_neg(a.c[0]), fp_neg(a.c[1]), fp_neg(a.c[2]), fp_neg(a.c[3])}}; }
R0_HD FpExt ext_scale(const FpExt& a, uint32_t s) {
  return FpExt{{fp_mul(a.c[0], s), fp_mul(a.c[1], s), fp_mul(a.c[2], s), fp_mul(a.c[3], s)}};
}
// (a0 + a1 X + a2 X^2 + a3 X^3)(b0 + ...) mod X^4 + 11. Each output coefficient is a sum of <= 4 products of
// canonical values; products whose X-power wraps are scaled by NBETA = P - 11.
R0_HD FpExt ext_mul(const FpExt& x, const FpExt& y) {
  const uint32_t* a = x.c;
  const uint32_t* b = y.c;
  FpExt r;
  r.c[0] = fp_add(fp_mul(a[0], b[0]),
                  fp_mul(FP_NBETA, fp_add(fp_add(fp_mul(a[1], b[3]), fp_mul(a[2], b[2])), fp_mul(a[3], b[1]))));
  r.c[1] = fp_add(fp_add(fp_mul(a[0], b[1]), fp_mul(a[1], b[0])),
                  fp_mul(FP_NBETA, fp_add(fp_mul(a[2], b[3]), fp_mul(a[3], b[2]))));
  r.c[2] = fp_add(fp_add(fp_add(fp_mul(a[0], b[2]), fp_mul(a[1], b[1])), fp_mul(a[2], b[0])),
                  fp_mul(FP_NBETA, fp_mul(a[3], b[3])));
  r.c[3] = fp_add(fp_add(fp_mul(a[0], b[3]), fp_mul(a[1], b[2])), fp_add(fp_mul(a[2], b[1]), fp_mul(a[3], b[0])));
  return r;
}
R0_HD FpExt ext_pow(FpExt x, uint64_t n) {
  FpExt tot = ext_one();
  while (n) {
    if (n & 1) tot = ext_mul(tot, x);
    n >>= 1;
    x = ext_mul(x, x);
  }
  return tot;
}
R0_HD FpExt ext_inv(const FpExt& x) {
  const uint32_t* a = x.c;
  uint32_t b0 = fp_add(fp_mul(a[0], a[0]),
                       fp_mul(FP_BETA, fp_sub(fp_mul(a[1], fp_add(a[3], a[3])), fp_mul(a[2], a[2]))));
  uint32_t b2 = fp_add(fp_sub(fp_mul(a[0], fp_add(a[2], a[2])), fp_mul(a[1], a[1])), fp_mul(FP_BETA, fp_mul(a[3], a[3])));
  uint32_t c = fp_add(fp_mul(b0, b0), fp_mul(fp_mul(FP_BETA, b2), b2));
  uint32_t ic = fp_inv(c);
  b0 = fp_mul(b0, ic);
  b2 = fp_mul(b2, ic);
  FpExt r;
  r.c[0] = fp_add(fp_mul(a[0], b0), fp_mul(fp_mul(FP_BETA, a[2]), b2));
  r.c[1] = fp_add(fp_mul(fp_neg(a[1]), b0), fp_mul(fp_mul(FP_NBETA, a[3]), b2));
  r.c[2] = fp_add(fp_mul(fp_neg(a[0]), b2), fp_mul(a[2], b0));
  r.c[3] = fp_sub(fp_mul(a[1], b2), fp_mul(a[3], b0));
  return r;
}

// ---- lazy 64-bit accumulation of products (device hot loops): acc stays < 2^63 --------------------------------
// acc += a*b with a, b canonical (< P): product < 2^61.8. If the running sum reaches 2^63 we subtract P*2^32
// (== 0 mod P after the final Montgomery reduction), which only touches the high word.
R0_HD void lazy_mac(uint64_t& acc, uint32_t a, uint32_t b) {
  acc += (uint64_t)a * b;
  uint32_t hi = (uint32_t)(acc >> 32);
  hi = umin32(hi, hi - P);  // hi >= P  <=>  acc >= P*2^32 ; keeps acc < P*2^32 + 2^61.8 < 2^63... see note
  acc = ((uint64_t)hi << 32) | (uint32_t)acc;
}
// After lazy_mac the accumulator is < P*2^32, which is exactly mont_reduce's precondition.
R0_HD uint32_t lazy_finish(uint64_t acc) { return mont_reduce(acc); }

R0_HD uint32_t brev32(uint32_t x) {
#if defined(__CUDA_ARCH__)
  return __brev(x);
#else
  x = ((x & 0xaaaaaaaau) >> 1) | ((x & 0x55555555u) << 1);
  x = ((x & 0xccccccccu) >> 2) | ((x & 0x33333333u) << 2);
  x = ((x & 0xf0f0f0f0u) >> 4) | ((x & 0x0f0f0f0fu) << 4);
  x = ((x & 0xff00ff00u) >> 8) | ((x & 0x00ff00ffu) << 8);
  return (x << 16) | (x >> 16);
#endif
}
R0_HD uint32_t brev_bits(uint32_t x, int bits) { return bits == 0 ? 0u : brev32(x) >> (32 - bits); }

}  // namespace r0
