// ORACLE — TEST INFRASTRUCTURE ONLY. CPU restatement of the reference STARK prover driver; never linked into the product.
//
// Follows /root/reference/risc0/zkp/src:
//   taps.rs:21-342              TapSet accessors (group_taps, regs, group_size, combos)
//   merkle.rs:32-67             MerkleTreeParams (layers, top_size)
//   prove/write_iop.rs:25-76    WriteIOP
//   prove/merkle.rs:54-138      MerkleTreeProver::{new, commit, prove}
//   prove/poly_group.rs:63-83   PolyGroup::new
//   prove/fri.rs:39-126         ProveRoundInfo::new / prove_query, fri_prove
//   prove/prover.rs:38-48       make_coeffs
//   prove/prover.rs:81-108      Prover::commit_group
//   prove/prover.rs:111-393     Prover::finalize
// Every step calls the CpuHal restatement in hal_cpu.h op by op, in the reference's order.
#pragma once
#include <functional>
#include <memory>
#include <string>
#include <vector>

#include "hal_cpu.h"

namespace oracle {

struct Tap {
  uint16_t offset, back, group, combo, skip;
};

struct TapSet {
  std::vector<Tap> taps;
  std::vector<uint16_t> combo_taps, combo_begin;
  std::vector<uint32_t> group_begin;  // num_groups + 1
  size_t combos_count = 0, reg_count = 0;

  static TapSet from_tables(const uint16_t* flat, size_t ntaps, const uint16_t* ctaps, size_t nctaps,
                            const uint16_t* cbegin, size_t ncombos, const uint32_t* gbegin, size_t ngroups,
                            size_t nregs) {
    TapSet t;
    for (size_t i = 0; i < ntaps; i++)
      t.taps.push_back({flat[5 * i], flat[5 * i + 1], flat[5 * i + 2], flat[5 * i + 3], flat[5 * i + 4]});
    t.combo_taps.assign(ctaps, ctaps + nctaps);
    t.combo_begin.assign(cbegin, cbegin + ncombos + 1);
    t.group_begin.assign(gbegin, gbegin + ngroups + 1);
    t.combos_count = ncombos;
    t.reg_count = nregs;
    return t;
  }
  size_t num_groups() const { return group_begin.size() - 1; }
  size_t tap_size() const { return group_begin.back(); }
  size_t group_size(size_t g) const { return size_t(taps[group_begin[g + 1] - 1].offset) + 1; }
  // register start indices (taps.rs RegisterIter) in [begin, end)
  std::vector<size_t> regs(size_t begin, size_t end) const {
    std::vector<size_t> r;
    size_t cur = begin;
    while (cur < taps.size()) {
      size_t next = cur + taps[cur].skip;
      if (next > end) break;
      r.push_back(cur);
      cur = next;
    }
    return r;
  }
  std::vector<size_t> all_regs() const { return regs(0, tap_size()); }
  std::vector<size_t> group_regs(size_t g) const { return regs(group_begin[g], group_begin[g + 1]); }
};

struct MerkleTreeParams {
  size_t row_size, col_size, queries, layers, top_layer, top_size;
  MerkleTreeParams(size_t rows, size_t cols, size_t q) : row_size(rows), col_size(cols), queries(q) {
    layers = log2_ceil(rows);
    if ((size_t(1) << layers) != rows) throw std::runtime_error("rows not a power of two");
    top_layer = 0;
    for (size_t i = 1; i < layers; i++) {
      if ((size_t(1) << i) > queries) break;
      top_layer = i;
    }
    top_size = size_t(1) << top_layer;
  }
};

struct WriteIOP {
  std::vector<uint32_t> proof;
  std::unique_ptr<Rng> rng;
  explicit WriteIOP(const HashSuite& s) : rng(s.new_rng()) {}
  void write_u32(const uint32_t* p, size_t n) { proof.insert(proof.end(), p, p + n); }
  void write_elems(const Fp* p, size_t n) { write_u32(reinterpret_cast<const uint32_t*>(p), n); }
  void write_ext(const FpExt* p, size_t n) { write_u32(reinterpret_cast<const uint32_t*>(p), 4 * n); }
  void write_digests(const Digest* p, size_t n) { write_u32(reinterpret_cast<const uint32_t*>(p), 8 * n); }
  void commit(const Digest& d) { rng->mix(d); }
  uint32_t random_bits(unsigned b) { return rng->random_bits(b); }
  Fp random_elem() { return rng->random_elem(); }
  FpExt random_ext_elem() { return rng->random_ext_elem(); }
};

// Records what a test wants to compare beyond the seal itself.
struct ProveTrace {
  std::vector<Digest> roots;           // every committed Merkle root, in commit order
  std::vector<uint32_t> query_pos;     // the 50 drawn FRI positions
};

struct MerkleTreeProver {
  MerkleTreeParams params;
  const Fp* matrix;
  std::vector<Digest> nodes;
  Digest root;
  MerkleTreeProver(const HashSuite& suite, const Fp* mat, size_t rows, size_t cols, size_t queries)
      : params(rows, cols, queries), matrix(mat), nodes(rows * 2) {
    hash_rows(suite, nodes.data() + rows, rows, matrix, rows * cols);
    for (size_t i = params.layers; i-- > 0;) {
      size_t layer_size = size_t(1) << i;
      hash_fold(suite, nodes.data(), layer_size * 2, layer_size);
    }
    root = nodes[1];
  }
  void commit(WriteIOP& iop, ProveTrace* tr) const {
    iop.write_digests(nodes.data() + params.top_size, params.top_size);
    iop.commit(root);
    if (tr) tr->roots.push_back(root);
  }
  void prove(WriteIOP& iop, size_t idx) const {
    std::vector<Fp> out(params.col_size);
    gather_sample(out.data(), matrix, idx, params.col_size, params.row_size);
    iop.write_elems(out.data(), out.size());
    idx += params.row_size;
    while (idx >= 2 * params.top_size) {
      size_t low_bit = idx % 2;
      idx /= 2;
      size_t other_idx = 2 * idx + (1 - low_bit);
      iop.write_digests(&nodes[other_idx], 1);
    }
  }
};

struct PolyGroup {
  std::vector<Fp> coeffs;  // natural order after construction
  size_t count;
  std::vector<Fp> evaluated;
  std::unique_ptr<MerkleTreeProver> merkle;
  PolyGroup(const HashSuite& suite, std::vector<Fp>&& c, size_t count_, size_t size) : coeffs(std::move(c)), count(count_) {
    size_t domain = size * INV_RATE;
    evaluated.resize(count * domain);
    batch_expand_into_evaluate_ntt(evaluated.data(), evaluated.size(), coeffs.data(), coeffs.size(), count,
                                   log2_ceil(INV_RATE));
    batch_bit_reverse(coeffs.data(), coeffs.size(), count);
    merkle.reset(new MerkleTreeProver(suite, evaluated.data(), domain, count, QUERIES));
  }
};

inline std::vector<Fp> make_coeffs(const Fp* witness, size_t size, size_t count) {
  std::vector<Fp> coeffs(size);
  eltwise_copy_elem(coeffs.data(), witness, size);
  batch_interpolate_ntt(coeffs.data(), size, count);
  zk_shift(coeffs.data(), size, count);
  return coeffs;
}

// CircuitHal::eval_check (hal/mod.rs:279-289): check[k*domain + i] for k<4. groups are the `evaluated` matrices in
// tap-group order (accum, code, data); globals in the order the circuit passes them.
using EvalCheckFn = std::function<void(Fp* check, const std::vector<const Fp*>& groups,
                                       const std::vector<const Fp*>& globals, FpExt poly_mix, size_t po2, size_t steps)>;

struct FriRound {
  size_t domain;
  std::vector<Fp> evaluated;
  std::vector<Fp> coeffs;
  std::unique_ptr<MerkleTreeProver> merkle;
};

inline void fri_prove(const HashSuite& suite, WriteIOP& iop, const std::vector<Fp>& in_coeffs,
                      const std::function<void(WriteIOP&, size_t)>& inner, ProveTrace* tr) {
  size_t orig_domain = in_coeffs.size() / EXT_SIZE * INV_RATE;
  std::vector<std::unique_ptr<FriRound>> rounds;
  const std::vector<Fp>* coeffs = &in_coeffs;
  while (coeffs->size() / EXT_SIZE > FRI_MIN_DEGREE) {
    std::unique_ptr<FriRound> r(new FriRound());
    size_t size = coeffs->size() / EXT_SIZE;
    r->domain = size * INV_RATE;
    r->evaluated.resize(r->domain * EXT_SIZE);
    batch_expand_into_evaluate_ntt(r->evaluated.data(), r->evaluated.size(), coeffs->data(), coeffs->size(), EXT_SIZE,
                                   log2_ceil(INV_RATE));
    r->merkle.reset(new MerkleTreeProver(suite, r->evaluated.data(), r->domain / FRI_FOLD, FRI_FOLD * EXT_SIZE, QUERIES));
    r->merkle->commit(iop, tr);
    FpExt fold_mix = iop.random_ext_elem();
    r->coeffs.resize(size / FRI_FOLD * EXT_SIZE);
    fri_fold(r->coeffs.data(), r->coeffs.size(), coeffs->data(), fold_mix);
    rounds.push_back(std::move(r));
    coeffs = &rounds.back()->coeffs;
  }
  std::vector<Fp> final_coeffs(coeffs->size());
  eltwise_copy_elem(final_coeffs.data(), coeffs->data(), coeffs->size());
  batch_bit_reverse(final_coeffs.data(), final_coeffs.size(), EXT_SIZE);
  iop.write_elems(final_coeffs.data(), final_coeffs.size());
  iop.commit(suite.hash_elem_slice(final_coeffs.data(), final_coeffs.size()));
  for (size_t q = 0; q < QUERIES; q++) {
    size_t pos = iop.random_bits(log2_ceil(orig_domain));
    if (tr) tr->query_pos.push_back(uint32_t(pos));
    inner(iop, pos);
    for (auto& r : rounds) {
      size_t group = pos % (r->domain / FRI_FOLD);
      r->merkle->prove(iop, group);
      pos = group;
    }
  }
}

struct Prover {
  const HashSuite& suite;
  const TapSet& taps;
  WriteIOP iop;
  std::vector<std::unique_ptr<PolyGroup>> groups;
  size_t cycles = 0, po2 = 0;
  ProveTrace* trace = nullptr;

  Prover(const HashSuite& s, const TapSet& t) : suite(s), taps(t), iop(s), groups(t.num_groups()) {}
  void set_po2(size_t p) {
    po2 = p;
    cycles = size_t(1) << p;
  }
  void commit_group(size_t g, const Fp* witness, size_t witness_size) {
    size_t group_size = taps.group_size(g);
    if (witness_size != group_size * cycles) throw std::runtime_error("commit_group: bad witness size");
    if (groups[g]) throw std::runtime_error("group committed twice");
    std::vector<Fp> coeffs = make_coeffs(witness, witness_size, group_size);
    groups[g].reset(new PolyGroup(suite, std::move(coeffs), group_size, cycles));
    groups[g]->merkle->commit(iop, trace);
  }

  std::vector<uint32_t> finalize(const std::vector<const Fp*>& globals, const EvalCheckFn& eval_check) {
    FpExt poly_mix = iop.random_ext_elem();
    size_t domain = cycles * INV_RATE;
    std::vector<Fp> check_poly(EXT_SIZE * domain);
    std::vector<const Fp*> gev;
    for (auto& g : groups) gev.push_back(g->evaluated.data());
    eval_check(check_poly.data(), gev, globals, poly_mix, po2, cycles);
    batch_interpolate_ntt(check_poly.data(), check_poly.size(), EXT_SIZE);
    PolyGroup check_group(suite, std::move(check_poly), CHECK_SIZE, cycles);
    check_group.merkle->commit(iop, trace);
    FpExt z = iop.random_ext_elem();
    FpExt back_one(rou_rev(unsigned(po2)));
    std::vector<FpExt> all_xs, eval_u;
    for (size_t id = 0; id < groups.size(); id++) {
      std::vector<uint32_t> which;
      std::vector<FpExt> xs;
      for (size_t t = taps.group_begin[id]; t < taps.group_begin[id + 1]; t++) {
        which.push_back(taps.taps[t].offset);
        FpExt x = back_one.pow(taps.taps[t].back) * z;
        xs.push_back(x);
        all_xs.push_back(x);
      }
      std::vector<FpExt> out(which.size());
      batch_evaluate_any(groups[id]->coeffs.data(), groups[id]->coeffs.size(), groups[id]->count, which.data(),
                         xs.data(), out.data(), which.size());
      eval_u.insert(eval_u.end(), out.begin(), out.end());
    }
    std::vector<FpExt> coeff_u(eval_u.size());
    {
      size_t pos = 0;
      for (size_t r : taps.all_regs()) {
        size_t sz = taps.taps[r].skip;
        poly_interpolate(&coeff_u[pos], coeff_u.size() - pos, &all_xs[pos], &eval_u[pos], sz);
        pos += sz;
      }
    }
    FpExt z_pow = z.pow(EXT_SIZE);
    {
      std::vector<uint32_t> which(CHECK_SIZE);
      for (size_t i = 0; i < CHECK_SIZE; i++) which[i] = uint32_t(i);
      std::vector<FpExt> xs(CHECK_SIZE, z_pow), out(CHECK_SIZE);
      batch_evaluate_any(check_group.coeffs.data(), check_group.coeffs.size(), CHECK_SIZE, which.data(), xs.data(),
                         out.data(), CHECK_SIZE);
      coeff_u.insert(coeff_u.end(), out.begin(), out.end());
      iop.write_ext(coeff_u.data(), coeff_u.size());
      iop.commit(suite.hash_ext_elem_slice(coeff_u.data(), coeff_u.size()));
    }
    FpExt mix = iop.random_ext_elem();
    size_t combo_count = taps.combos_count;
    std::vector<FpExt> combos(cycles * (combo_count + 1));
    {
      FpExt cur_mix = FpExt::one();
      for (size_t id = 0; id < groups.size(); id++) {
        size_t group_size = taps.group_size(id);
        std::vector<uint32_t> which;
        for (size_t r : taps.group_regs(id)) which.push_back(taps.taps[r].combo);
        mix_poly_coeffs(combos.data(), combos.size(), cur_mix, mix, groups[id]->coeffs.data(), which.data(),
                        group_size, cycles);
        cur_mix *= mix.pow(group_size);
      }
      std::vector<uint32_t> which(CHECK_SIZE, uint32_t(combo_count));
      mix_poly_coeffs(combos.data(), combos.size(), cur_mix, mix, check_group.coeffs.data(), which.data(), CHECK_SIZE,
                      cycles);
    }
    {
      std::vector<uint32_t> reg_sizes, reg_combo_ids;
      for (size_t r : taps.all_regs()) {
        reg_sizes.push_back(taps.taps[r].skip);
        reg_combo_ids.push_back(taps.taps[r].combo);
      }
      combos_prepare(combos.data(), coeff_u.data(), combo_count, cycles, reg_sizes.data(), reg_combo_ids.data(),
                     reg_sizes.size(), mix);
      std::vector<std::vector<FpExt>> chunks;
      for (size_t i = 0; i < combo_count; i++) {
        std::vector<FpExt> pows;
        for (size_t k = taps.combo_begin[i]; k < taps.combo_begin[i + 1]; k++)
          pows.push_back(z * back_one.pow(taps.combo_taps[k]));
        chunks.push_back(pows);
      }
      chunks.push_back({z_pow});
      if (!combos_divide(combos.data(), chunks, cycles)) throw std::runtime_error("combos_divide: nonzero remainder");
    }
    std::vector<Fp> final_poly_coeffs(cycles * EXT_SIZE);
    eltwise_sum_extelem(final_poly_coeffs.data(), final_poly_coeffs.size(), combos.data(), combos.size());
    batch_bit_reverse(final_poly_coeffs.data(), final_poly_coeffs.size(), EXT_SIZE);
    fri_prove(
        suite, iop, final_poly_coeffs,
        [&](WriteIOP& w, size_t idx) {
          for (auto& g : groups) g->merkle->prove(w, idx);
          check_group.merkle->prove(w, idx);
        },
        trace);
    return iop.proof;
  }
};

}  // namespace oracle
