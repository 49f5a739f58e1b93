// Host side of the Fiat-Shamir transcript used by the segment prover driver (prover.cu): the two hash suites'
// small-input hashing and their random-number generators. The reference keeps this on the host too
// (Hal::get_hash_suite, risc0/zkp/src/hal/mod.rs:63-65): only bulk hashing (hash_rows / hash_fold) runs on the GPU.
//
// Restates:
//   Poseidon2 permutation     risc0/zkp/src/core/hash/poseidon2/mod.rs:102-216 (consts.rs tables)
//   unpadded sponge / pair    poseidon2/mod.rs:46-58,221-244
//   Poseidon2Rng              poseidon2/rng.rs:26-89
//   SHA-256 suite             core/hash/sha/cpu.rs:56-98, sha/rng.rs:27-101, baby_bear.rs:109-139 (Elem::random)
//   WriteIOP                  prove/write_iop.rs:25-76
#pragma once
#include <string.h>

#include <memory>
#include <vector>

#include "fp.cuh"
#include "tables/poseidon2_tables.h"

namespace r0 {

struct Digest {
  uint32_t w[8];
};

// ---------------------------------------------------------------- Poseidon2 on the host (Montgomery words throughout)
struct P2Host {
  static uint32_t sbox(uint32_t x) {
    uint32_t x2 = fp_mul(x, x), x4 = fp_mul(x2, x2);
    return fp_mul(fp_mul(x4, x2), x);
  }
  static void external(uint32_t* c) {
    uint32_t col[4] = {0, 0, 0, 0};
    for (int g = 0; g < 6; g++) {
      uint32_t* x = c + 4 * g;
      uint32_t t0 = fp_add(x[0], x[1]), t1 = fp_add(x[2], x[3]);
      uint32_t t2 = fp_add(fp_add(x[1], x[1]), t1), t3 = fp_add(fp_add(x[3], x[3]), t0);
      uint32_t t1x4 = fp_add(t1, t1), t0x4 = fp_add(t0, t0);
      t1x4 = fp_add(t1x4, t1x4);
      t0x4 = fp_add(t0x4, t0x4);
      uint32_t t4 = fp_add(t1x4, t3), t5 = fp_add(t0x4, t2);
      x[0] = fp_add(t3, t5);
      x[1] = t5;
      x[2] = fp_add(t2, t4);
      x[3] = t4;
      for (int j = 0; j < 4; j++) col[j] = fp_add(col[j], x[j]);
    }
    for (int i = 0; i < 24; i++) c[i] = fp_add(c[i], col[i & 3]);
  }
  static void permute(uint32_t* c) {
    external(c);
    for (int round = 0; round < 29; round++) {
      const bool full = round < 4 || round >= 25;
      if (full) {
        const uint32_t* rc = R0_P2_RC_FULL_MONT + 24 * (round < 4 ? round : round - 21);
        for (int i = 0; i < 24; i++) c[i] = sbox(fp_add(c[i], rc[i]));
        external(c);
      } else {
        c[0] = sbox(fp_add(c[0], R0_P2_RC_PARTIAL_MONT[round - 4]));
        uint32_t sum = 0;
        for (int i = 0; i < 24; i++) sum = fp_add(sum, c[i]);
        for (int i = 0; i < 24; i++) c[i] = fp_add(sum, fp_mul(R0_P2_DIAG_MONT[i], c[i]));
      }
    }
  }
  static Digest hash_words(const uint32_t* data, size_t n) {
    uint32_t st[24] = {0};
    size_t fill = 0;
    for (size_t i = 0; i < n; i++) {
      st[fill++] = data[i];
      if (fill == 16) {
        permute(st);
        fill = 0;
      }
    }
    if (fill != 0 || n == 0) {
      for (size_t i = fill; i < 16; i++) st[i] = 0;
      permute(st);
    }
    Digest d;
    memcpy(d.w, st, 32);
    return d;
  }
};

// ---------------------------------------------------------------- SHA-256 on the host
struct ShaHost {
  static uint32_t ror(uint32_t x, int n) { return (x >> n) | (x << (32 - n)); }
  static void compress(uint32_t* h, const uint32_t* block) {
    static const uint32_t K[64] = {
        0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5,
        0xd807aa98, 0x12835b01, 0x243185be, 0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174,
        0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc, 0x2de92c6f, 0x4a7484aa, 0x5cb0a9dc, 0x76f988da,
        0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147, 0x06ca6351, 0x14292967,
        0x27b70a85, 0x2e1b2138, 0x4d2c6dfc, 0x53380d13, 0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85,
        0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3, 0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070,
        0x19a4c116, 0x1e376c08, 0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f, 0x682e6ff3,
        0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208, 0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2};
    uint32_t w[64];
    for (int t = 0; t < 16; t++) w[t] = block[t];
    for (int t = 16; t < 64; t++) {
      uint32_t a = w[t - 15], b = w[t - 2];
      w[t] = w[t - 16] + (ror(a, 7) ^ ror(a, 18) ^ (a >> 3)) + w[t - 7] + (ror(b, 17) ^ ror(b, 19) ^ (b >> 10));
    }
    uint32_t v[8];
    memcpy(v, h, 32);
    for (int t = 0; t < 64; t++) {
      uint32_t e = v[4], a = v[0];
      uint32_t t1 = v[7] + (ror(e, 6) ^ ror(e, 11) ^ ror(e, 25)) + ((e & v[5]) ^ (~e & v[6])) + K[t] + w[t];
      uint32_t t2 = (ror(a, 2) ^ ror(a, 13) ^ ror(a, 22)) + ((a & v[1]) ^ (a & v[2]) ^ (v[1] & v[2]));
      for (int j = 7; j > 0; j--) v[j] = v[j - 1];
      v[4] += t1;
      v[0] = t1 + t2;
    }
    for (int j = 0; j < 8; j++) h[j] += v[j];
  }
  static void init(uint32_t* h) {
    static const uint32_t IV[8] = {0x6a09e667, 0xbb67ae85, 0x3c6ef372, 0xa54ff53a,
                                   0x510e527f, 0x9b05688c, 0x1f83d9ab, 0x5be0cd19};
    memcpy(h, IV, 32);
  }
  // no padding; a short last block is zero filled; words are hashed as their little-endian bytes
  static Digest hash_words(const uint32_t* data, size_t n) {
    uint32_t h[8], blk[16];
    init(h);
    size_t fill = 0;
    for (size_t i = 0; i < n; i++) {
      blk[fill++] = __builtin_bswap32(data[i]);
      if (fill == 16) {
        compress(h, blk);
        fill = 0;
      }
    }
    if (fill) {
      for (size_t i = fill; i < 16; i++) blk[i] = 0;
      compress(h, blk);
    }
    Digest d;
    for (int j = 0; j < 8; j++) d.w[j] = __builtin_bswap32(h[j]);
    return d;
  }
  // standard padded SHA-256 of a short byte string (only used to seed ShaRng with "Hello" / "World")
  static Digest hash_short_bytes(const char* s) {
    size_t len = strlen(s);
    uint8_t m[64] = {0};
    memcpy(m, s, len);
    m[len] = 0x80;
    m[63] = (uint8_t)(len * 8);
    m[62] = (uint8_t)((len * 8) >> 8);
    uint32_t h[8], blk[16];
    init(h);
    for (int i = 0; i < 16; i++) blk[i] = (uint32_t(m[4 * i]) << 24) | (uint32_t(m[4 * i + 1]) << 16) | (uint32_t(m[4 * i + 2]) << 8) | m[4 * i + 3];
    compress(h, blk);
    Digest d;
    for (int j = 0; j < 8; j++) d.w[j] = __builtin_bswap32(h[j]);
    return d;
  }
};

// ---------------------------------------------------------------- suite + transcript
struct HostSuite {
  int kind;  // R0B200_HASH_POSEIDON2 = 0, R0B200_HASH_SHA256 = 1
  Digest hash_words(const uint32_t* data, size_t n) const {
    return kind == 0 ? P2Host::hash_words(data, n) : ShaHost::hash_words(data, n);
  }
};

class Transcript {
 public:
  std::vector<uint32_t> proof;

  explicit Transcript(int kind) : kind_(kind) {
    memset(cells_, 0, sizeof(cells_));
    if (kind_ == 1) {
      pool0_ = ShaHost::hash_short_bytes("Hello");
      pool1_ = ShaHost::hash_short_bytes("World");
    }
  }
  void write(const uint32_t* p, size_t n) { proof.insert(proof.end(), p, p + n); }
  void commit(const Digest& d) {
    if (kind_ == 0) {
      if (used_ != 0) {
        P2Host::permute(cells_);
        used_ = 0;
      }
      for (int i = 0; i < 8; i++) cells_[i] = fp_add(cells_[i], d.w[i]);
      P2Host::permute(cells_);
    } else {
      for (int i = 0; i < 8; i++) pool0_.w[i] ^= d.w[i];
      sha_step();
    }
  }
  uint32_t random_elem() {  // Montgomery word
    if (kind_ == 0) {
      if (used_ == 16) {
        P2Host::permute(cells_);
        used_ = 0;
      }
      return cells_[used_++];
    }
    uint64_t v = 0;
    for (int i = 0; i < 6; i++) v = ((v << 32) + sha_next()) % P;
    return fp_encode((uint32_t)v);
  }
  FpExt random_ext() {
    FpExt e;
    for (int i = 0; i < 4; i++) e.c[i] = random_elem();
    return e;
  }
  uint32_t random_bits(unsigned bits) {
    if (kind_ == 1) return ((1u << bits) - 1u) & sha_next();
    uint32_t val = fp_decode(random_elem());
    for (int i = 0; i < 3; i++) {
      uint32_t nv = fp_decode(random_elem());
      if (val == 0) val = nv;
    }
    return ((1u << bits) - 1u) & val;
  }

 private:
  void sha_step() {
    uint32_t both[16];
    memcpy(both, pool0_.w, 32);
    memcpy(both + 8, pool1_.w, 32);
    pool0_ = ShaHost::hash_words(both, 16);
    memcpy(both, pool0_.w, 32);
    pool1_ = ShaHost::hash_words(both, 16);
    used_ = 0;
  }
  uint32_t sha_next() {
    if (used_ == 8) sha_step();
    return pool0_.w[used_++];
  }
  int kind_;
  uint32_t cells_[24];
  size_t used_ = 0;
  Digest pool0_{}, pool1_{};
};

}  // namespace r0
