// ORACLE — TEST INFRASTRUCTURE ONLY. CPU restatement of the reference STARK verifier; never linked into the product.
//
// Follows /root/reference/risc0/zkp/src/verify:
//   read_iop.rs:24-84    ReadIOP
//   merkle.rs:84-187     MerkleTreeVerifier::{new, verify}
//   fri.rs:34-155        VerifyRoundInfo, verify_query, fri_verify
//   mod.rs:246-286       fri_eval_taps (DEEP quotient per query)
//   mod.rs:293-453       verify_validity
//   mod.rs:461-481       read_slice_with_po2
//   mod.rs:500-560       verify (rv32im / recursion / keccak protocol)
//   mod.rs:566-612       verify_v3 (golden-seal protocol, `verify/proof.bin`, HelloCircuit mod.rs:614-727)
// Used (a) to pin the oracle against the reference's golden seal and (b) to accept seals made by the oracle prover
// and by the CUDA product path. `poly_ext` may be empty: the rv32im `poly_ext.rs` is a missing blob in the reference
// snapshot, so for rv32im seals everything EXCEPT the constraint-evaluation equality is checked.
#pragma once
#include <functional>
#include <string>
#include <vector>

#include "prover.h"

namespace oracle {

struct VerifyError : std::runtime_error {
  using std::runtime_error::runtime_error;
};

struct ReadIOP {
  const uint32_t* p;
  size_t left;
  std::unique_ptr<Rng> rng;
  ReadIOP(const HashSuite& s, const uint32_t* seal, size_t n) : p(seal), left(n), rng(s.new_rng()) {}
  const uint32_t* read_u32s(size_t n) {
    if (n > left) throw VerifyError("seal truncated");
    const uint32_t* r = p;
    p += n;
    left -= n;
    return r;
  }
  const Fp* read_elems(size_t n) {
    const uint32_t* r = read_u32s(n);
    for (size_t i = 0; i < n; i++)
      if (r[i] >= P) throw VerifyError("invalid field element in seal");
    return reinterpret_cast<const Fp*>(r);
  }
  const FpExt* read_ext(size_t n) { return reinterpret_cast<const FpExt*>(read_elems(4 * n)); }
  const Digest* read_digests(size_t n) { return reinterpret_cast<const Digest*>(read_u32s(8 * n)); }
  void commit(const Digest& d) { rng->mix(d); }
};

struct MerkleTreeVerifier {
  MerkleTreeParams params;
  std::vector<Digest> node;  // heap index 1 .. 2*top_size-1
  MerkleTreeVerifier(ReadIOP& iop, const HashSuite& suite, size_t rows, size_t cols, size_t queries)
      : params(rows, cols, queries), node(2 * params.top_size) {
    const Digest* top = iop.read_digests(params.top_size);
    for (size_t i = 0; i < params.top_size; i++) node[params.top_size + i] = top[i];
    for (size_t i = params.top_size; i-- > 1;) node[i] = suite.hash_pair(node[2 * i], node[2 * i + 1]);
    iop.commit(root());
  }
  const Digest& root() const { return node[1]; }
  const Fp* verify(ReadIOP& iop, const HashSuite& suite, size_t idx) const {
    if (idx >= params.row_size) throw VerifyError("merkle query out of range");
    const Fp* out = iop.read_elems(params.col_size);
    Digest cur = suite.hash_elem_slice(out, params.col_size);
    idx += params.row_size;
    while (idx >= 2 * params.top_size) {
      size_t low_bit = idx % 2;
      const Digest* other = iop.read_digests(1);
      idx /= 2;
      cur = low_bit == 1 ? suite.hash_pair(*other, cur) : suite.hash_pair(cur, *other);
    }
    if (node[idx] != cur) throw VerifyError("merkle path mismatch");
    return out;
  }
};

// poly_ext(poly_mix, eval_u, args) -> tot
using PolyExtFn = std::function<FpExt(FpExt poly_mix, const std::vector<FpExt>& eval_u,
                                      const std::vector<std::vector<Fp>>& args)>;

struct Verifier {
  const TapSet& taps;
  const HashSuite& suite;
  ReadIOP iop;
  size_t po2 = 0, tot_cycles = 0;
  std::vector<std::unique_ptr<MerkleTreeVerifier>> merkle;
  std::vector<Digest> roots;  // group roots in read order, then check root, then FRI roots
  bool validity_checked = false;

  Verifier(const TapSet& t, const HashSuite& s, const uint32_t* seal, size_t n)
      : taps(t), suite(s), iop(s, seal, n), merkle(t.num_groups()) {}

  void commit_info16(const char* info) {
    Fp elems[16];
    for (int i = 0; i < 16; i++) elems[i] = Fp(uint32_t(uint8_t(info[i])));
    iop.commit(suite.hash_elem_slice(elems, 16));
  }
  void verify_group(size_t g) {
    if (merkle[g]) throw VerifyError("group verified twice");
    merkle[g].reset(new MerkleTreeVerifier(iop, suite, INV_RATE * tot_cycles, taps.group_size(g), QUERIES));
    roots.push_back(merkle[g]->root());
  }
  std::vector<Fp> read_rng(size_t n) {
    std::vector<Fp> v(n);
    for (auto& x : v) x = iop.rng->random_elem();
    return v;
  }
  std::vector<Fp> read_slice_with_po2(size_t size) {
    const Fp* s = iop.read_elems(size + 1);
    iop.commit(suite.hash_elem_slice(s, size + 1));
    po2 = s[size].v;  // stored as a RAW word (rv32im/src/prove/hal/mod.rs:203)
    if (po2 > 24) throw VerifyError("po2 too large");
    tot_cycles = size_t(1) << po2;
    return std::vector<Fp>(s, s + size);
  }
  static FpExt poly_eval_(const FpExt* c, size_t n, FpExt x) { return poly_eval(c, n, x); }

  void verify_validity(const PolyExtFn& poly_ext, const std::vector<std::vector<Fp>>& args) {
    for (auto& m : merkle)
      if (!m) throw VerifyError("missing group");
    FpExt poly_mix = iop.rng->random_ext_elem();
    size_t domain = INV_RATE * tot_cycles;
    MerkleTreeVerifier check_merkle(iop, suite, domain, CHECK_SIZE, QUERIES);
    roots.push_back(check_merkle.root());
    FpExt z = iop.rng->random_ext_elem();
    Fp back_one = rou_rev(unsigned(po2));
    size_t num_taps = taps.tap_size();
    const FpExt* coeff_u = iop.read_ext(num_taps + CHECK_SIZE);
    iop.commit(suite.hash_ext_elem_slice(coeff_u, num_taps + CHECK_SIZE));
    std::vector<FpExt> eval_u;
    std::vector<size_t> regs = taps.all_regs();
    {
      size_t cur_pos = 0;
      for (size_t r : regs) {
        size_t sz = taps.taps[r].skip;
        for (size_t i = 0; i < sz; i++) {
          FpExt x = z * back_one.pow(taps.taps[r + i].back);
          eval_u.push_back(poly_eval(&coeff_u[cur_pos], sz, x));
        }
        cur_pos += sz;
      }
    }
    if (poly_ext) {
      FpExt result = poly_ext(poly_mix, eval_u, args);
      FpExt check;
      const size_t remap[4] = {0, 2, 1, 3};
      for (size_t i = 0; i < 4; i++) {
        size_t rmi = remap[i];
        FpExt zi = z.pow(i);
        check += coeff_u[num_taps + rmi] * zi * FpExt(Fp(1), Fp(), Fp(), Fp());
        check += coeff_u[num_taps + rmi + 4] * zi * FpExt(Fp(), Fp(1), Fp(), Fp());
        check += coeff_u[num_taps + rmi + 8] * zi * FpExt(Fp(), Fp(), Fp(1), Fp());
        check += coeff_u[num_taps + rmi + 12] * zi * FpExt(Fp(), Fp(), Fp(), Fp(1));
      }
      check *= (FpExt(Fp(3)) * z).pow(tot_cycles) - FpExt::one();
      if (check != result) throw VerifyError("check != result (constraint polynomial)");
      validity_checked = true;
    }
    FpExt mix = iop.rng->random_ext_elem();
    size_t tot_combo_backs = taps.combo_taps.size();
    std::vector<FpExt> combo_u(tot_combo_backs + 1);
    std::vector<FpExt> tap_mix_pows, check_mix_pows;
    FpExt cur_mix = FpExt::one();
    size_t cur_pos = 0;
    for (size_t r : regs) {
      size_t sz = taps.taps[r].skip;
      for (size_t i = 0; i < sz; i++) combo_u[taps.combo_begin[taps.taps[r].combo] + i] += cur_mix * coeff_u[cur_pos + i];
      tap_mix_pows.push_back(cur_mix);
      cur_mix *= mix;
      cur_pos += sz;
    }
    for (size_t i = 0; i < CHECK_SIZE; i++) {
      combo_u[tot_combo_backs] += cur_mix * coeff_u[cur_pos];
      cur_pos++;
      check_mix_pows.push_back(cur_mix);
      cur_mix *= mix;
    }
    Fp gen = rou_fwd(log2_ceil(domain));
    size_t combo_count = taps.combos_count;
    fri_verify([&](size_t idx) -> FpExt {
      Fp xb = gen.pow(idx);
      std::vector<const Fp*> rows;
      for (auto& m : merkle) rows.push_back(m->verify(iop, suite, idx));
      const Fp* check_row = check_merkle.verify(iop, suite, idx);
      // fri_eval_taps
      std::vector<FpExt> tot(combo_count + 1);
      FpExt x(xb);
      for (size_t k = 0; k < regs.size(); k++) {
        const Tap& t = taps.taps[regs[k]];
        tot[t.combo] += tap_mix_pows[k] * rows[t.group][t.offset];
      }
      for (size_t i = 0; i < CHECK_SIZE; i++) tot[combo_count] += check_mix_pows[i] * check_row[i];
      FpExt ret;
      for (size_t i = 0; i < combo_count; i++) {
        size_t b = taps.combo_begin[i], e = taps.combo_begin[i + 1];
        FpExt num = tot[i] - poly_eval(&combo_u[b], e - b, x);
        FpExt divisor = FpExt::one();
        for (size_t k = b; k < e; k++) divisor *= x - z * back_one.pow(taps.combo_taps[k]);
        ret += num * divisor.inv();
      }
      FpExt check_num = tot[combo_count] - combo_u[tot_combo_backs];
      FpExt check_div = x - z.pow(INV_RATE);
      ret += check_num * check_div.inv();
      return ret;
    });
  }

  void fri_verify(const std::function<FpExt(size_t)>& inner) {
    size_t degree = tot_cycles;
    size_t orig_domain = INV_RATE * degree;
    size_t domain = orig_domain;
    struct Round {
      size_t domain;
      std::unique_ptr<MerkleTreeVerifier> merkle;
      FpExt mix;
    };
    std::vector<Round> rounds;
    while (degree > FRI_MIN_DEGREE) {
      Round r;
      r.domain = domain / FRI_FOLD;
      r.merkle.reset(new MerkleTreeVerifier(iop, suite, r.domain, FRI_FOLD * EXT_SIZE, QUERIES));
      roots.push_back(r.merkle->root());
      r.mix = iop.rng->random_ext_elem();
      rounds.push_back(std::move(r));
      domain /= FRI_FOLD;
      degree /= FRI_FOLD;
    }
    const Fp* final_coeffs = iop.read_elems(EXT_SIZE * degree);
    iop.commit(suite.hash_elem_slice(final_coeffs, EXT_SIZE * degree));
    Fp gen = rou_fwd(log2_ceil(domain));
    std::vector<FpExt> poly_buf(degree);
    for (size_t i = 0; i < degree; i++)
      poly_buf[i] = FpExt(final_coeffs[0 * degree + i], final_coeffs[1 * degree + i], final_coeffs[2 * degree + i],
                          final_coeffs[3 * degree + i]);
    for (size_t q = 0; q < QUERIES; q++) {
      size_t pos = iop.rng->random_bits(log2_ceil(orig_domain));
      FpExt goal = inner(pos);
      for (auto& r : rounds) {
        size_t quot = pos / r.domain, group = pos % r.domain;
        const Fp* data = r.merkle->verify(iop, suite, group);
        FpExt data_ext[FRI_FOLD];
        for (size_t i = 0; i < FRI_FOLD; i++)
          data_ext[i] = FpExt(data[0 * FRI_FOLD + i], data[1 * FRI_FOLD + i], data[2 * FRI_FOLD + i], data[3 * FRI_FOLD + i]);
        if (data_ext[quot] != goal) throw VerifyError("FRI goal mismatch");
        unsigned root_po2 = log2_ceil(FRI_FOLD * r.domain);
        Fp inv_wk = rou_rev(root_po2).pow(group);
        interpolate_ntt(data_ext, FRI_FOLD);
        bit_reverse(data_ext, FRI_FOLD);
        goal = poly_eval(data_ext, FRI_FOLD, r.mix * inv_wk);
        pos = group;
      }
      Fp x = gen.pow(pos);
      if (poly_eval(poly_buf.data(), degree, FpExt(x)) != goal) throw VerifyError("FRI final polynomial mismatch");
    }
  }
  void verify_complete() {
    if (iop.left != 0) throw VerifyError("trailing words in seal");
  }
};

// verify/mod.rs:500-560 ; `version_words` leading u32s are skipped (rv32im writes RV32IM_SEAL_VERSION first and the
// rv32im crate strips it before calling verify).
inline void verify_standard(const TapSet& taps, const HashSuite& suite, const uint32_t* seal, size_t n,
                            const char* circuit_info, size_t output_size, size_t mix_size, size_t version_words,
                            const PolyExtFn& poly_ext, Verifier** out_v = nullptr) {
  if (n <= version_words) throw VerifyError("empty seal");
  static thread_local std::unique_ptr<Verifier> keep;
  keep.reset(new Verifier(taps, suite, seal + version_words, n - version_words));
  Verifier& v = *keep;
  if (out_v) *out_v = &v;
  v.commit_info16("RISC0_STARK:v1__");
  v.commit_info16(circuit_info);
  std::vector<Fp> out = v.read_slice_with_po2(output_size);
  v.verify_group(1);  // code
  v.verify_group(2);  // data
  std::vector<Fp> mix = v.read_rng(mix_size);
  v.verify_group(0);  // accum
  v.verify_validity(poly_ext, {out, mix});
  v.verify_complete();
}

// verify/mod.rs:566-612 with all-zero GroupInfo (HelloCircuit)
inline void verify_v3_simple(const TapSet& taps, const HashSuite& suite, const uint32_t* seal, size_t n, size_t po2,
                             const PolyExtFn& poly_ext, Verifier** out_v = nullptr) {
  if (n == 0) throw VerifyError("empty seal");
  static thread_local std::unique_ptr<Verifier> keep;
  keep.reset(new Verifier(taps, suite, seal, n));
  Verifier& v = *keep;
  if (out_v) *out_v = &v;
  v.po2 = po2;
  v.tot_cycles = size_t(1) << po2;
  for (size_t g = 0; g < taps.num_groups(); g++) v.verify_group(g);
  v.verify_validity(poly_ext, {{}, {}});
  v.verify_complete();
}

}  // namespace oracle
