// ORACLE — TEST INFRASTRUCTURE ONLY. CPU restatement of the reference's hash suites; never linked into the product.
//
// Poseidon2 (t=24, rate 16, out 8) follows /root/reference/risc0/zkp/src/core/hash/poseidon2/mod.rs:
//   hash_pair :46-58, to_digest :94-100, round-constant adds :102-110, sbox x^7 :112-117,
//   multiply_by_m_int :131-137, 4x4 circulant :139-151, multiply_by_m_ext :153-177,
//   poseidon2_mix :193-216, unpadded_hash :221-244; RNG: poseidon2/rng.rs:26-89.
// SHA-256 suite follows core/hash/sha/cpu.rs:56-76 (hash_raw_data_slice: no padding, zero-filled tail block,
// byte-swapped state words), :79-98 (compress = hash_pair), sha/rng.rs:27-101 (ShaRng) and
// baby_bear.rs:109-139 (Elem::random: six u32 folded mod P).
// Pinned by KATs: permutation + hash KATs poseidon2/mod.rs:330-401, RNG KAT prove/merkle.rs:161-172,
// SHA hash_rows KAT hal/cpu.rs:726-733 (tests/test_oracle_kats.py).
#pragma once
#include <array>
#include <cstring>
#include <memory>
#include <vector>

#include "field.h"
#include "tables/poseidon2_tables.h"

namespace oracle {

struct Digest {
  uint32_t w[8];
  bool operator==(const Digest& o) const { return std::memcmp(w, o.w, 32) == 0; }
  bool operator!=(const Digest& o) const { return !(*this == o); }
};

// ------------------------------------------------------------------ Poseidon2
constexpr int P2_CELLS = 24, P2_RATE = 16, P2_OUT = 8;
using P2State = std::array<Fp, P2_CELLS>;

inline Fp p2_sbox(Fp x) {
  Fp x2 = x * x;
  Fp x4 = x2 * x2;
  Fp x6 = x4 * x2;
  return x6 * x;
}

inline void p2_m_ext(P2State& c) {
  P2State old = c;
  Fp tmp[4] = {};
  const Fp two(2), four(4);
  for (int i = 0; i < P2_CELLS / 4; i++) {
    const Fp* x = &old[i * 4];
    Fp t0 = x[0] + x[1];
    Fp t1 = x[2] + x[3];
    Fp t2 = two * x[1] + t1;
    Fp t3 = two * x[3] + t0;
    Fp t4 = four * t1 + t3;
    Fp t5 = four * t0 + t2;
    Fp t6 = t3 + t5;
    Fp t7 = t2 + t4;
    Fp out[4] = {t6, t5, t7, t4};
    for (int j = 0; j < 4; j++) {
      tmp[j] += out[j];
      c[i * 4 + j] = out[j];
    }
  }
  for (int i = 0; i < P2_CELLS; i++) c[i] += tmp[i % 4];
}

inline void p2_m_int(P2State& c) {
  Fp sum;
  for (int i = 0; i < P2_CELLS; i++) sum += c[i];
  for (int i = 0; i < P2_CELLS; i++) c[i] = sum + Fp(R0_P2_DIAG[i]) * c[i];
}

inline void poseidon2_mix(P2State& c) {
  p2_m_ext(c);
  int full = 0;
  auto full_round = [&]() {
    for (int i = 0; i < P2_CELLS; i++) c[i] = p2_sbox(c[i] + Fp(R0_P2_RC_FULL[full * P2_CELLS + i]));
    p2_m_ext(c);
    full++;
  };
  for (int r = 0; r < 4; r++) full_round();
  for (int r = 0; r < 21; r++) {
    c[0] = p2_sbox(c[0] + Fp(R0_P2_RC_PARTIAL[r]));
    p2_m_int(c);
  }
  for (int r = 0; r < 4; r++) full_round();
}

// Unpadded overwrite-mode sponge over `n` elements read with `stride`.
inline Digest poseidon2_hash_elems(const Fp* data, size_t n, size_t stride = 1) {
  P2State st{};
  size_t unmixed = 0;
  for (size_t i = 0; i < n; i++) {
    st[unmixed++] = data[i * stride];
    if (unmixed == P2_RATE) {
      poseidon2_mix(st);
      unmixed = 0;
    }
  }
  if (unmixed != 0 || n == 0) {
    for (size_t i = unmixed; i < P2_RATE; i++) st[i] = Fp();
    poseidon2_mix(st);
  }
  Digest d;
  for (int i = 0; i < P2_OUT; i++) d.w[i] = st[i].v;
  return d;
}

inline Digest poseidon2_hash_pair(const Digest& a, const Digest& b) {
  Fp both[16];
  for (int i = 0; i < 8; i++) {
    both[i] = Fp::raw(a.w[i]);
    both[8 + i] = Fp::raw(b.w[i]);
  }
  return poseidon2_hash_elems(both, 16);
}

// ------------------------------------------------------------------ SHA-256 (FIPS 180-4 compression function)
namespace sha_detail {
static const uint32_t K[64] = {
    0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5, 0xd807aa98,
    0x12835b01, 0x243185be, 0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174, 0xe49b69c1, 0xefbe4786,
    0x0fc19dc6, 0x240ca1cc, 0x2de92c6f, 0x4a7484aa, 0x5cb0a9dc, 0x76f988da, 0x983e5152, 0xa831c66d, 0xb00327c8,
    0xbf597fc7, 0xc6e00bf3, 0xd5a79147, 0x06ca6351, 0x14292967, 0x27b70a85, 0x2e1b2138, 0x4d2c6dfc, 0x53380d13,
    0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85, 0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3, 0xd192e819,
    0xd6990624, 0xf40e3585, 0x106aa070, 0x19a4c116, 0x1e376c08, 0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a,
    0x5b9cca4f, 0x682e6ff3, 0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208, 0x90befffa, 0xa4506ceb, 0xbef9a3f7,
    0xc67178f2};
static const uint32_t INIT[8] = {0x6a09e667, 0xbb67ae85, 0x3c6ef372, 0xa54ff53a,
                                 0x510e527f, 0x9b05688c, 0x1f83d9ab, 0x5be0cd19};
inline uint32_t rotr(uint32_t x, int n) { return (x >> n) | (x << (32 - n)); }
inline uint32_t bswap(uint32_t x) { return __builtin_bswap32(x); }
// state: native big-endian-value words; block: 16 message words (already big-endian-decoded)
inline void compress(uint32_t st[8], const uint32_t m[16]) {
  uint32_t w[64];
  for (int i = 0; i < 16; i++) w[i] = m[i];
  for (int i = 16; i < 64; i++) {
    uint32_t s0 = rotr(w[i - 15], 7) ^ rotr(w[i - 15], 18) ^ (w[i - 15] >> 3);
    uint32_t s1 = rotr(w[i - 2], 17) ^ rotr(w[i - 2], 19) ^ (w[i - 2] >> 10);
    w[i] = w[i - 16] + s0 + w[i - 7] + s1;
  }
  uint32_t a = st[0], b = st[1], c = st[2], d = st[3], e = st[4], f = st[5], g = st[6], h = st[7];
  for (int i = 0; i < 64; i++) {
    uint32_t S1 = rotr(e, 6) ^ rotr(e, 11) ^ rotr(e, 25);
    uint32_t ch = (e & f) ^ (~e & g);
    uint32_t t1 = h + S1 + ch + K[i] + w[i];
    uint32_t S0 = rotr(a, 2) ^ rotr(a, 13) ^ rotr(a, 22);
    uint32_t mj = (a & b) ^ (a & c) ^ (b & c);
    uint32_t t2 = S0 + mj;
    h = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
  }
  st[0] += a; st[1] += b; st[2] += c; st[3] += d; st[4] += e; st[5] += f; st[6] += g; st[7] += h;
}
}  // namespace sha_detail

// Unpadded SHA of little-endian u32 words read with `stride` (the reference hashes the raw bytes of the slice).
inline Digest sha_hash_raw_words(const uint32_t* data, size_t n, size_t stride = 1) {
  using namespace sha_detail;
  uint32_t st[8];
  std::memcpy(st, INIT, 32);
  uint32_t blk[16];
  size_t cur = 0;
  for (size_t i = 0; i < n; i++) {
    blk[cur++] = bswap(data[i * stride]);  // bytes of a LE word, read as a BE message word
    if (cur == 16) {
      compress(st, blk);
      cur = 0;
    }
  }
  if (cur != 0) {
    for (size_t i = cur; i < 16; i++) blk[i] = 0;
    compress(st, blk);
  }
  Digest d;
  for (int i = 0; i < 8; i++) d.w[i] = bswap(st[i]);
  return d;
}
inline Digest sha_hash_pair(const Digest& a, const Digest& b) {
  uint32_t words[16];
  std::memcpy(words, a.w, 32);
  std::memcpy(words + 8, b.w, 32);
  return sha_hash_raw_words(words, 16);
}
// Standard padded SHA-256 of a byte string, digest as the reference stores it (ne-bytes words of the byte digest).
inline Digest sha_hash_bytes(const uint8_t* bytes, size_t len) {
  using namespace sha_detail;
  std::vector<uint8_t> msg(bytes, bytes + len);
  msg.push_back(0x80);
  while (msg.size() % 64 != 56) msg.push_back(0);
  uint64_t bits = uint64_t(len) * 8;
  for (int i = 7; i >= 0; i--) msg.push_back(uint8_t(bits >> (8 * i)));
  uint32_t st[8];
  std::memcpy(st, INIT, 32);
  for (size_t off = 0; off < msg.size(); off += 64) {
    uint32_t blk[16];
    for (int i = 0; i < 16; i++)
      blk[i] = (uint32_t(msg[off + 4 * i]) << 24) | (uint32_t(msg[off + 4 * i + 1]) << 16) |
               (uint32_t(msg[off + 4 * i + 2]) << 8) | uint32_t(msg[off + 4 * i + 3]);
    compress(st, blk);
  }
  Digest d;
  for (int i = 0; i < 8; i++) d.w[i] = bswap(st[i]);
  return d;
}

// ------------------------------------------------------------------ suites (HashFn + Rng), core/hash/mod.rs:30-77
struct Rng {
  virtual ~Rng() {}
  virtual void mix(const Digest& d) = 0;
  virtual uint32_t random_bits(unsigned bits) = 0;
  virtual Fp random_elem() = 0;
  FpExt random_ext_elem() {
    Fp a = random_elem(), b = random_elem(), c = random_elem(), d = random_elem();
    return FpExt(a, b, c, d);
  }
};

struct Poseidon2Rng : Rng {
  P2State cells{};
  size_t pool_used = 0;
  void mix(const Digest& d) override {
    if (pool_used != 0) {
      poseidon2_mix(cells);
      pool_used = 0;
    }
    for (int i = 0; i < P2_OUT; i++) cells[i] += Fp::raw(d.w[i]);
    poseidon2_mix(cells);
  }
  Fp random_elem() override {
    if (pool_used == P2_RATE) {
      poseidon2_mix(cells);
      pool_used = 0;
    }
    return cells[pool_used++];
  }
  uint32_t random_bits(unsigned bits) override {
    uint32_t val = random_elem().as_u32();
    for (int i = 0; i < 3; i++) {
      uint32_t nv = random_elem().as_u32();
      if (val == 0) val = nv;
    }
    return ((1u << bits) - 1) & val;
  }
};

struct ShaRng : Rng {
  Digest pool0, pool1;
  size_t pool_used = 0;
  ShaRng() {
    pool0 = sha_hash_bytes(reinterpret_cast<const uint8_t*>("Hello"), 5);
    pool1 = sha_hash_bytes(reinterpret_cast<const uint8_t*>("World"), 5);
  }
  void step() {
    pool0 = sha_hash_pair(pool0, pool1);
    pool1 = sha_hash_pair(pool0, pool1);
    pool_used = 0;
  }
  uint32_t next_u32() {
    if (pool_used == 8) step();
    return pool0.w[pool_used++];
  }
  void mix(const Digest& d) override {
    for (int i = 0; i < 8; i++) pool0.w[i] ^= d.w[i];
    step();
  }
  uint32_t random_bits(unsigned bits) override { return ((1u << bits) - 1) & next_u32(); }
  Fp random_elem() override {
    uint64_t val = 0;
    for (int i = 0; i < 6; i++) {
      val <<= 32;
      val += next_u32();
      val %= P;
    }
    return Fp(uint32_t(val));
  }
};

enum class HashKind { Poseidon2 = 0, Sha256 = 1 };

struct HashSuite {
  HashKind kind;
  explicit HashSuite(HashKind k) : kind(k) {}
  Digest hash_elem_slice(const Fp* data, size_t n, size_t stride = 1) const {
    return kind == HashKind::Poseidon2 ? poseidon2_hash_elems(data, n, stride)
                                       : sha_hash_raw_words(reinterpret_cast<const uint32_t*>(data), n, stride);
  }
  Digest hash_ext_elem_slice(const FpExt* data, size_t n) const {
    return hash_elem_slice(reinterpret_cast<const Fp*>(data), n * 4);
  }
  Digest hash_pair(const Digest& a, const Digest& b) const {
    return kind == HashKind::Poseidon2 ? poseidon2_hash_pair(a, b) : sha_hash_pair(a, b);
  }
  std::unique_ptr<Rng> new_rng() const {
    if (kind == HashKind::Poseidon2) return std::unique_ptr<Rng>(new Poseidon2Rng());
    return std::unique_ptr<Rng>(new ShaRng());
  }
};

}  // namespace oracle
