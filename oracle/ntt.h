// ORACLE — TEST INFRASTRUCTURE ONLY. CPU restatement of the reference NTT; never linked into the product.
//
// Follows /root/reference/risc0/zkp/src/core/ntt.rs:
//   bit_rev_32 :34-45, bit_reverse :64-73, fwd_butterfly :91-112, rev_butterfly :115-133,
//   interpolate_ntt :232-282, evaluate_ntt :284-330, expand :334-343
// and core/poly.rs: poly_eval :23-32, poly_interpolate :38-76, poly_divide :81-89.
#pragma once
#include <vector>

#include "field.h"
#include "tables/field_tables.h"

namespace oracle {

inline Fp rou_fwd(unsigned k) { return Fp(R0_ROU_FWD[k]); }
inline Fp rou_rev(unsigned k) { return Fp(R0_ROU_REV[k]); }

inline uint32_t bit_rev_32(uint32_t x) {
  x = ((x & 0xaaaaaaaau) >> 1) | ((x & 0x55555555u) << 1);
  x = ((x & 0xccccccccu) >> 2) | ((x & 0x33333333u) << 2);
  x = ((x & 0xf0f0f0f0u) >> 4) | ((x & 0x0f0f0f0fu) << 4);
  x = ((x & 0xff00ff00u) >> 8) | ((x & 0x00ff00ffu) << 8);
  return (x << 16) | (x >> 16);
}
inline uint32_t brev(uint32_t i, unsigned bits) { return bits == 0 ? 0 : bit_rev_32(i) >> (32 - bits); }

template <typename T>
void bit_reverse(T* io, size_t len) {
  unsigned n = log2_ceil(len);
  for (size_t i = 0; i < len; i++) {
    size_t r = brev(uint32_t(i), n);
    if (i < r) {
      T t = io[i];
      io[i] = io[r];
      io[r] = t;
    }
  }
}

// T is Fp or FpExt; twiddles are always base-field.
template <typename T>
void fwd_butterfly(T* io, unsigned n, unsigned expand_bits) {
  if (n == 0 || n == expand_bits) return;
  size_t half = size_t(1) << (n - 1);
  fwd_butterfly(io, n - 1, expand_bits);
  fwd_butterfly(io + half, n - 1, expand_bits);
  Fp step = rou_fwd(n);
  Fp cur(1);
  for (size_t i = 0; i < half; i++) {
    T a = io[i];
    T b = io[i + half] * cur;
    io[i] = a + b;
    io[i + half] = a - b;
    cur *= step;
  }
}

template <typename T>
void rev_butterfly(T* io, unsigned n) {
  if (n == 0) return;
  size_t half = size_t(1) << (n - 1);
  Fp step = rou_rev(n);
  Fp cur(1);
  for (size_t i = 0; i < half; i++) {
    T a = io[i];
    T b = io[i + half];
    io[i] = a + b;
    io[i + half] = (a - b) * cur;
    cur *= step;
  }
  rev_butterfly(io, n - 1);
  rev_butterfly(io + half, n - 1);
}

template <typename T>
void interpolate_ntt(T* io, size_t size) {
  unsigned n = log2_ceil(size);
  rev_butterfly(io, n);
  Fp norm = Fp::from_u64(size).inv();
  for (size_t i = 0; i < size; i++) io[i] = io[i] * norm;
}

template <typename T>
void evaluate_ntt(T* io, size_t size, unsigned expand_bits) {
  fwd_butterfly(io, log2_ceil(size), expand_bits);
}

template <typename T>
void expand(T* out, const T* in, size_t in_len, unsigned expand_bits) {
  size_t size_out = in_len << expand_bits;
  for (size_t i = 0; i < size_out; i++) out[i] = in[i >> expand_bits];
}

inline FpExt poly_eval(const FpExt* coeffs, size_t n, FpExt x) {
  FpExt mul = FpExt::one(), tot;
  for (size_t i = 0; i < n; i++) {
    tot += coeffs[i] * mul;
    mul *= x;
  }
  return tot;
}

inline FpExt poly_divide(FpExt* p, size_t n, FpExt z) {
  FpExt cur;
  for (size_t i = n; i-- > 0;) {
    FpExt next = z * cur + p[i];
    p[i] = cur;
    cur = next;
  }
  return cur;
}

// out has `out_len` valid entries starting at the register's position (the reference clears the whole tail slice).
inline void poly_interpolate(FpExt* out, size_t out_len, const FpExt* x, const FpExt* fx, size_t size) {
  if (size == 1) {
    out[0] = fx[0];
    return;
  }
  if (size == 2) {
    out[1] = (fx[1] - fx[0]) * (x[1] - x[0]).inv();
    out[0] = fx[0] - out[1] * x[0];
    return;
  }
  std::vector<FpExt> ft(size + 1);
  ft[0] = FpExt::one();
  for (size_t i = 0; i < size; i++) {
    for (size_t j = i + 1; j-- > 0;) {
      FpExt value = ft[j];
      ft[j + 1] += value;
      ft[j] *= -x[i];
    }
  }
  for (size_t i = 0; i < out_len; i++) out[i] = FpExt();
  for (size_t i = 0; i < size; i++) {
    std::vector<FpExt> fr = ft;
    poly_divide(fr.data(), fr.size(), x[i]);
    FpExt fr_xi = poly_eval(fr.data(), fr.size(), x[i]);
    FpExt mul = fx[i] * fr_xi.inv();
    for (size_t j = 0; j < size; j++) out[j] += mul * fr[j];
  }
}

}  // namespace oracle
