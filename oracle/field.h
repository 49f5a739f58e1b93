// ORACLE — TEST INFRASTRUCTURE ONLY. CPU restatement of the reference's BabyBear field; never linked into the product.
//
// Follows /root/reference/risc0/core/src/field/baby_bear.rs:
//   P, M, R2                 :41-42, :84
//   INVALID marker           :93
//   inv = x^(P-2), inv(0)=0  :98-107
//   add / sub / mul (Montgomery, R = 2^32)  :324-361
//   ExtElem = Fp[X]/(X^4+11), mul :744-757, inv :448-487
// Pinned by the reference KATs in tests/test_oracle_field.py (5^1000 = 589699054, baby_bear.rs:893-894; FpExt
// product vectors :815-853).
#pragma once
#include <cstddef>
#include <cstdint>

namespace oracle {

constexpr uint32_t P = 15u * (1u << 27) + 1u;  // 0x78000001
constexpr uint32_t M = 0x88000001u;            // -P^-1 ... (baby_bear.rs:41) used as in mul() below
constexpr uint32_t R2 = 1172168163u;           // 2^64 mod P
constexpr uint32_t INVALID_RAW = 0xffffffffu;

constexpr uint32_t fp_add_raw(uint32_t a, uint32_t b) {
  uint32_t x = a + b;
  return x >= P ? x - P : x;
}
constexpr uint32_t fp_sub_raw(uint32_t a, uint32_t b) {
  uint32_t x = a - b;
  return x > P ? x + P : x;
}
constexpr uint32_t fp_mul_raw(uint32_t a, uint32_t b) {
  uint64_t o64 = uint64_t(a) * uint64_t(b);
  uint32_t low = 0u - uint32_t(o64);
  uint32_t red = M * low;
  o64 += uint64_t(red) * uint64_t(P);
  uint32_t ret = uint32_t(o64 >> 32);
  return ret >= P ? ret - P : ret;
}

struct Fp {
  uint32_t v;  // Montgomery form, canonical (< P), or INVALID_RAW
  constexpr Fp() : v(0) {}
  constexpr explicit Fp(uint32_t normal) : v(fp_mul_raw(R2, normal % P)) {}
  static constexpr Fp raw(uint32_t m) {
    Fp r;
    r.v = m;
    return r;
  }
  static constexpr Fp from_u64(uint64_t x) { return Fp(uint32_t(x % P)); }
  static constexpr Fp invalid() { return raw(INVALID_RAW); }
  constexpr uint32_t as_u32() const { return fp_mul_raw(1, v); }  // decode
  constexpr bool is_valid() const { return v != INVALID_RAW; }
  constexpr Fp valid_or_zero() const { return is_valid() ? *this : Fp(); }
  constexpr Fp operator+(Fp o) const { return raw(fp_add_raw(v, o.v)); }
  constexpr Fp operator-(Fp o) const { return raw(fp_sub_raw(v, o.v)); }
  constexpr Fp operator*(Fp o) const { return raw(fp_mul_raw(v, o.v)); }
  constexpr Fp operator-() const { return raw(fp_sub_raw(0, v)); }
  Fp& operator+=(Fp o) { return *this = *this + o; }
  Fp& operator-=(Fp o) { return *this = *this - o; }
  Fp& operator*=(Fp o) { return *this = *this * o; }
  constexpr bool operator==(Fp o) const { return v == o.v; }
  constexpr bool operator!=(Fp o) const { return v != o.v; }
  constexpr Fp pow(uint64_t n) const {
    Fp tot(1), x = *this;
    while (n) {
      if (n & 1) tot = tot * x;
      n >>= 1;
      x = x * x;
    }
    return tot;
  }
  constexpr Fp inv() const { return pow(P - 2); }
};

struct FpExt {
  Fp e[4];
  constexpr FpExt() : e{} {}
  constexpr explicit FpExt(Fp a) : e{a, Fp(), Fp(), Fp()} {}
  constexpr FpExt(Fp a, Fp b, Fp c, Fp d) : e{a, b, c, d} {}
  static constexpr FpExt one() { return FpExt(Fp(1)); }
  constexpr FpExt operator+(const FpExt& o) const { return {e[0] + o.e[0], e[1] + o.e[1], e[2] + o.e[2], e[3] + o.e[3]}; }
  constexpr FpExt operator-(const FpExt& o) const { return {e[0] - o.e[0], e[1] - o.e[1], e[2] - o.e[2], e[3] - o.e[3]}; }
  constexpr FpExt operator-() const { return {-e[0], -e[1], -e[2], -e[3]}; }
  constexpr FpExt operator*(Fp s) const { return {e[0] * s, e[1] * s, e[2] * s, e[3] * s}; }
  constexpr FpExt operator*(const FpExt& o) const {
    // baby_bear.rs:744-757 : X^4 = -11
    const Fp NBETA(P - 11);
    const Fp* a = e;
    const Fp* b = o.e;
    return {a[0] * b[0] + NBETA * (a[1] * b[3] + a[2] * b[2] + a[3] * b[1]),
            a[0] * b[1] + a[1] * b[0] + NBETA * (a[2] * b[3] + a[3] * b[2]),
            a[0] * b[2] + a[1] * b[1] + a[2] * b[0] + NBETA * (a[3] * b[3]),
            a[0] * b[3] + a[1] * b[2] + a[2] * b[1] + a[3] * b[0]};
  }
  FpExt& operator+=(const FpExt& o) { return *this = *this + o; }
  FpExt& operator-=(const FpExt& o) { return *this = *this - o; }
  FpExt& operator*=(const FpExt& o) { return *this = *this * o; }
  constexpr bool operator==(const FpExt& o) const {
    return e[0] == o.e[0] && e[1] == o.e[1] && e[2] == o.e[2] && e[3] == o.e[3];
  }
  constexpr bool operator!=(const FpExt& o) const { return !(*this == o); }
  constexpr FpExt pow(uint64_t n) const {
    FpExt tot = one(), x = *this;
    while (n) {
      if (n & 1) tot = tot * x;
      n >>= 1;
      x = x * x;
    }
    return tot;
  }
  constexpr FpExt inv() const {
    // baby_bear.rs:448-487
    const Fp BETA(11), NBETA(P - 11);
    const Fp* a = e;
    Fp b0 = a[0] * a[0] + BETA * (a[1] * (a[3] + a[3]) - a[2] * a[2]);
    Fp b2 = a[0] * (a[2] + a[2]) - a[1] * a[1] + BETA * (a[3] * a[3]);
    Fp c = b0 * b0 + BETA * b2 * b2;
    Fp ic = c.inv();
    b0 = b0 * ic;
    b2 = b2 * ic;
    return {a[0] * b0 + BETA * a[2] * b2, -a[1] * b0 + NBETA * a[3] * b2, -a[0] * b2 + a[2] * b0,
            a[1] * b2 - a[3] * b0};
  }
};

static_assert(sizeof(Fp) == 4 && sizeof(FpExt) == 16, "layout must match the reference's repr(transparent) types");

inline unsigned log2_ceil(size_t x) {
  unsigned r = 0;
  while ((size_t(1) << r) < x) r++;
  return r;
}

}  // namespace oracle
