// Segment prover driver: the callers of the Hal on the hot path (SURVEY §8f "callers either side"), i.e. what
// `SegmentProverImpl::prove_core`'s prove_inner block does with a committed witness, as one stream-ordered C++ routine.
//
// Mirrors, op for op and in the same transcript order:
//   risc0/circuit/rv32im/src/prove/hal/mod.rs:171-222   prove_core (header, group commits, mix draw, finalize)
//   risc0/zkp/src/prove/prover.rs:38-48,81-108,111-393  make_coeffs, commit_group, finalize
//   risc0/zkp/src/prove/poly_group.rs:63-83             PolyGroup::new
//   risc0/zkp/src/prove/merkle.rs:54-138                MerkleTreeProver::{new, commit, prove}
//   risc0/zkp/src/prove/fri.rs:39-126                   fri_prove
//   risc0/zkp/src/taps.rs:21-342                        TapSet accessors
//   risc0/zkp/src/core/poly.rs:38-89                    poly_interpolate (host, <= 6 points per register)
// What is different from the reference's flow: the host only synchronises where Fiat-Shamir needs a digest or an
// evaluation back (one small D2H per commitment); make_coeffs fuses interpolate + zk_shift; each Merkle tree is one
// call; all 350 query openings are gathered by two launches after the 50 positions have been drawn (drawing them has
// no side effects on the transcript other than the RNG itself, so the seal is unchanged).
#include <algorithm>
#include <memory>
#include <stdexcept>
#include <vector>

#include "../../include/r0b200.h"
#include "ctx.h"
#include "launchers.h"
#include "tables/circuit_recursion.h"
#include "tables/circuit_rv32im.h"
#include "tables/field_tables.h"
#include "transcript.h"

namespace r0 {
namespace {

constexpr size_t QUERIES = 50, INV_RATE = 4, FRI_FOLD = 16, FRI_MIN_DEGREE = 256, CHECK_SIZE = 16, EXT = 4;

struct DevBuf {
  Ctx* c = nullptr;
  uint32_t* p = nullptr;
  size_t words = 0;
  DevBuf() {}
  // `alloc_stream`: the stream the allocation is ordered on (default: the compute stream). The buffer is always
  // released on the compute stream, after the kernels that used it.
  DevBuf(Ctx* c_, size_t words_, cudaStream_t alloc_stream = nullptr) : c(c_), words(words_) {
    R0_CUDA(r0_malloc_async(c, &p, (words ? words : 4) * 4, alloc_stream ? alloc_stream : c->stream));
    c->bytes_allocated += words * 4;
    if (c->bytes_allocated > c->bytes_peak) c->bytes_peak = c->bytes_allocated;
  }
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  DevBuf(DevBuf&& o) noexcept { *this = std::move(o); }
  DevBuf& operator=(DevBuf&& o) noexcept {
    release();
    c = o.c;
    p = o.p;
    words = o.words;
    o.p = nullptr;
    return *this;
  }
  void release() {
    if (p) {
      cudaFreeAsync(p, c->stream);
      c->bytes_allocated -= words * 4;
      p = nullptr;
    }
  }
  ~DevBuf() { release(); }
};

struct Tap {
  uint16_t offset, back, group, combo, skip;
};

struct TapSet {
  std::vector<Tap> taps;
  std::vector<uint16_t> combo_taps, combo_begin;
  std::vector<uint32_t> group_begin;
  std::vector<uint32_t> group_sizes;
  size_t combos_count = 0;
  size_t num_groups() const { return group_begin.size() - 1; }
  size_t tap_size() const { return group_begin.back(); }
  std::vector<size_t> regs(size_t begin, size_t end) const {  // first tap of each register in [begin, end)
    std::vector<size_t> r;
    for (size_t cur = begin; cur < end; cur += taps[cur].skip) r.push_back(cur);
    return r;
  }
};

// Everything circuit-specific the driver needs: tap set, sizes, the info string hashed into the transcript, whether
// a seal version word is written first (rv32im/src/prove/hal/mod.rs:181-183 does, the recursion prover does not,
// recursion/src/prove/mod.rs:179-224), and the circuit's eval_check launcher.
struct CircuitDesc {
  const char* info;
  const uint16_t* taps;
  size_t ntaps;
  const uint16_t* combo_taps;
  size_t ncombo_taps;
  const uint16_t* combo_begin;
  size_t ncombos;
  const uint32_t* group_begin;
  const uint32_t* group_sizes;
  size_t output_size, mix_size;
  bool has_version;
  uint32_t version;
  void (*eval_check)(Ctx*, uint32_t*, const uint32_t*, const uint32_t*, const uint32_t*, const uint32_t*, const uint32_t*,
                     const FpExt&, uint32_t);
};

const CircuitDesc kRv32im = {RV32IM_CIRCUIT_INFO, RV32IM_TAPS, RV32IM_NUM_TAPS, RV32IM_COMBO_TAPS, RV32IM_TOT_COMBO_BACKS,
                             RV32IM_COMBO_BEGIN, RV32IM_NUM_COMBOS, RV32IM_GROUP_BEGIN, RV32IM_GROUP_SIZES,
                             RV32IM_OUTPUT_SIZE, RV32IM_MIX_SIZE, true, RV32IM_SEAL_VERSION, r0_eval_check_rv32im};
const CircuitDesc kRecursion = {RECURSION_CIRCUIT_INFO, RECURSION_TAPS, RECURSION_NUM_TAPS, RECURSION_COMBO_TAPS,
                                RECURSION_TOT_COMBO_BACKS, RECURSION_COMBO_BEGIN, RECURSION_NUM_COMBOS,
                                RECURSION_GROUP_BEGIN, RECURSION_GROUP_SIZES, RECURSION_OUTPUT_SIZE, RECURSION_MIX_SIZE,
                                false, 0, r0_eval_check_recursion};

TapSet make_taps(const CircuitDesc& d) {
  TapSet t;
  for (size_t i = 0; i < d.ntaps; i++) {
    const uint16_t* f = d.taps + 5 * i;
    t.taps.push_back({f[0], f[1], f[2], f[3], f[4]});
  }
  t.combo_taps.assign(d.combo_taps, d.combo_taps + d.ncombo_taps);
  t.combo_begin.assign(d.combo_begin, d.combo_begin + d.ncombos + 1);
  t.group_begin.assign(d.group_begin, d.group_begin + 4);
  t.group_sizes.assign(d.group_sizes, d.group_sizes + 3);
  t.combos_count = d.ncombos;
  return t;
}

size_t log2_exact(size_t n) {
  size_t k = 0;
  while ((size_t(1) << k) < n) k++;
  R0_CHECK((size_t(1) << k) == n, "size is not a power of two");
  return k;
}

// ---- host-side interpolation over <= 6 points (poly.rs:38-76); exact arithmetic makes the method irrelevant, so
// this is plain Lagrange: out = sum_i fx[i] * prod_{j != i} (X - x_j) / (x_i - x_j)
void poly_interpolate(FpExt* out, const FpExt* x, const FpExt* fx, size_t n) {
  for (size_t k = 0; k < n; k++) out[k] = ext_zero();
  std::vector<FpExt> basis(n);
  for (size_t i = 0; i < n; i++) {
    for (size_t k = 0; k < n; k++) basis[k] = ext_zero();
    basis[0] = ext_one();
    size_t deg = 0;
    FpExt denom = ext_one();
    for (size_t j = 0; j < n; j++) {
      if (j == i) continue;
      // basis *= (X - x_j)
      for (size_t k = deg + 1; k-- > 0;) {
        FpExt up = basis[k];
        if (k + 1 < n) basis[k + 1] = ext_add(basis[k + 1], up);
        basis[k] = ext_mul(up, ext_neg(x[j]));
      }
      deg++;
      denom = ext_mul(denom, ext_sub(x[i], x[j]));
    }
    FpExt scale = ext_mul(fx[i], ext_inv(denom));
    for (size_t k = 0; k < n; k++) out[k] = ext_add(out[k], ext_mul(scale, basis[k]));
  }
}

struct Opening {  // one MerkleTreeProver::prove call, resolved after the batched gathers
  size_t val_off, cols, dig_off, ndig;
};

struct Openings {
  std::vector<GatherJob> vjobs;
  std::vector<DigestJob> djobs;
  std::vector<Opening> list;
  size_t val_words = 0;
};

struct MerkleTree {
  DevBuf nodes;
  const uint32_t* matrix = nullptr;
  size_t rows = 0, cols = 0, top_size = 0;
  Digest root{};

  MerkleTree() {}
  MerkleTree(Ctx* c, int hash, const uint32_t* mat, size_t rows_, size_t cols_) : nodes(c, 2 * rows_ * 8), matrix(mat), rows(rows_), cols(cols_) {
    size_t layers = log2_exact(rows);
    size_t top_layer = 0;
    for (size_t i = 1; i < layers; i++) {
      if ((size_t(1) << i) > QUERIES) break;
      top_layer = i;
    }
    top_size = size_t(1) << top_layer;
    const char* e = r0b200_merkle_build((r0b200_ctx*)c, hash, nodes.p, matrix, rows, cols);
    if (e) {
      std::string msg(e);
      r0b200_free_error(e);
      throw std::runtime_error(msg);
    }
  }
  // writes nodes[top_size .. 2*top_size) to the proof and commits the root (merkle.rs:83-96)
  void commit(Ctx* c, Transcript& iop, std::vector<Digest>* roots) {
    NvtxRange range("commit");   // merkle.rs:85
    std::vector<uint32_t> top(8 * top_size);
    {
      // fetch nodes [1, 2*top_size) in one copy: node 1 is the root, the last top_size of them are the top layer
      std::vector<uint32_t> head(8 * 2 * top_size);
      R0_CUDA(cudaMemcpyAsync(head.data() + 8, nodes.p + 8, (2 * top_size - 1) * 32, cudaMemcpyDeviceToHost, c->stream));
      R0_CUDA(cudaStreamSynchronize(c->stream));
      memcpy(root.w, head.data() + 8, 32);
      memcpy(top.data(), head.data() + 8 * top_size, 32 * top_size);
    }
    iop.write(top.data(), top.size());
    iop.commit(root);
    if (roots) roots->push_back(root);
  }
  void prove(Openings& o, size_t idx) const {  // merkle.rs:108-138
    Opening op{o.val_words, cols, o.djobs.size(), 0};
    o.vjobs.push_back(GatherJob{matrix, (uint64_t)idx, (uint64_t)rows, (uint32_t)cols, (uint32_t)o.val_words});
    o.val_words += cols;
    idx += rows;
    while (idx >= 2 * top_size) {
      size_t low = idx & 1;
      idx >>= 1;
      o.djobs.push_back(DigestJob{nodes.p, (uint64_t)(2 * idx + (1 - low))});
      op.ndig++;
    }
    o.list.push_back(op);
  }
};

struct PolyGroup {
  DevBuf coeffs, evaluated;
  MerkleTree merkle;
  size_t count = 0;
};

}  // namespace
}  // namespace r0

// A witness whose host->device upload has been enqueued on the context's copy stream (r0b200_witness_upload): the
// three coefficient buffers a proof starts from, plus one event per uploaded column chunk. Allocation and copies are
// ordered on the copy stream only, so they overlap whatever the compute stream is doing (the previous proof).
struct r0b200_witness {
  r0::Ctx* c = nullptr;
  uint32_t po2 = 0;
  int circuit = 0;  // 0 = rv32im, 1 = recursion
  uint32_t* coeffs[3] = {nullptr, nullptr, nullptr};
  size_t words[3] = {0, 0, 0};
  std::vector<cudaEvent_t> events[3];
};

namespace r0 {
namespace {

class SegmentProver {
 public:
  SegmentProver(Ctx* c, const CircuitDesc& desc, int hash, size_t po2)
      : c_(c), desc_(desc), hash_(hash), po2_(po2), cycles_(size_t(1) << po2), iop_(hash), suite_{hash}, taps_(make_taps(desc)) {
    groups_.resize(taps_.num_groups());
  }
  // Error paths: pending H2D chunks may still be reading the caller's host buffers and writing the coefficient
  // buffers that are about to be freed on the compute stream, so drain the copy stream first.
  ~SegmentProver() {
    bool pending = false;
    for (auto& v : uploads_) {
      for (cudaEvent_t ev : v) {
        pending = true;
        cudaEventDestroy(ev);
      }
      v.clear();
    }
    if (pending) cudaStreamSynchronize(c_->copy_stream);
  }
  SegmentProver(const SegmentProver&) = delete;
  SegmentProver& operator=(const SegmentProver&) = delete;
  Transcript& iop() { return iop_; }
  bool has_group(size_t g) const { return groups_[g].coeffs.p != nullptr; }
  std::vector<Digest> roots;
  std::vector<uint32_t> query_pos;

  // Host-resident witness: allocate the group's coefficient buffer now and enqueue its upload on the copy stream in
  // chunks of columns, one event per chunk. commit_group() later makes the compute stream wait chunk by chunk, so the
  // PCIe transfer of chunk j+1 (and of the next group) overlaps the NTTs / hashing of what is already on the device.
  void prefetch_group(size_t g, const uint32_t* witness_host) {
    const size_t count = taps_.group_sizes[g];
    PolyGroup& pg = groups_[g];
    pg.count = count;
    pg.coeffs = DevBuf(c_, count * cycles_);
    cudaEvent_t ready;
    R0_CUDA(cudaEventCreateWithFlags(&ready, cudaEventDisableTiming));
    R0_CUDA(cudaEventRecord(ready, c_->stream));            // the allocation is ordered on the compute stream
    R0_CUDA(cudaStreamWaitEvent(c_->copy_stream, ready, 0));
    R0_CUDA(cudaEventDestroy(ready));
    const size_t chunk = upload_chunk_cols();
    for (size_t c0 = 0; c0 < count; c0 += chunk) {
      const size_t nc = std::min(chunk, count - c0);
      R0_CUDA(cudaMemcpyAsync(pg.coeffs.p + c0 * cycles_, witness_host + c0 * cycles_, nc * cycles_ * 4,
                              cudaMemcpyHostToDevice, c_->copy_stream));
      cudaEvent_t ev;
      R0_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
      R0_CUDA(cudaEventRecord(ev, c_->copy_stream));
      uploads_[g].push_back(ev);
    }
  }
  // take over group g's buffer and chunk events from an uploaded witness (r0b200_witness_upload)
  void adopt_group(size_t g, r0b200_witness* w) {
    PolyGroup& pg = groups_[g];
    pg.count = taps_.group_sizes[g];
    R0_CHECK(w->words[g] == pg.count * cycles_, "uploaded witness does not match the circuit / po2");
    pg.coeffs.c = c_;
    pg.coeffs.p = w->coeffs[g];
    pg.coeffs.words = w->words[g];
    c_->bytes_allocated += w->words[g] * 4;
    if (c_->bytes_allocated > c_->bytes_peak) c_->bytes_peak = c_->bytes_allocated;
    w->coeffs[g] = nullptr;
    uploads_[g] = std::move(w->events[g]);
    w->events[g].clear();
  }
  size_t upload_chunk_cols() const {
    size_t cols = (size_t(64) << 20) / (cycles_ * 4);   // about 64 MB per chunk
    return cols ? cols : 1;
  }

  // Prover::commit_group (prover.rs:81-108). `witness` is a device pointer, or NULL when prefetch_group() was used.
  void commit_group(size_t g, const uint32_t* witness_dev) {
    NvtxRange range("commit_group");   // prover.rs:82
    const size_t count = taps_.group_sizes[g];
    PolyGroup& pg = groups_[g];
    std::unique_ptr<NvtxRange> coeffs_range(new NvtxRange("make_coeffs"));   // prover.rs:39
    if (witness_dev) {
      pg.count = count;
      pg.coeffs = DevBuf(c_, count * cycles_);
      // make_coeffs' copy (poly_group.rs:63-83) is folded into the transform: it reads the witness and writes coeffs
      r0_ntt_interpolate(c_, pg.coeffs.p, count, (int)po2_, /*zk=*/true, 0, witness_dev);
    } else {
      const size_t chunk = upload_chunk_cols();
      size_t j = 0;
      for (size_t c0 = 0; c0 < count; c0 += chunk, j++) {
        const size_t nc = std::min(chunk, count - c0);
        R0_CUDA(cudaStreamWaitEvent(c_->stream, uploads_[g][j], 0));
        R0_CUDA(cudaEventDestroy(uploads_[g][j]));
        r0_ntt_interpolate(c_, pg.coeffs.p + c0 * cycles_, nc, (int)po2_, /*zk=*/true, 0);
      }
      uploads_[g].clear();
    }
    coeffs_range.reset();
    finish_group(pg);
    pg.merkle.commit(c_, iop_, &roots);
  }

  // Prover::finalize (prover.rs:111-393); globals: mix (36 words) and out (90 words), host Montgomery words
  void finalize(const uint32_t* mix_host, const uint32_t* out_host) {
    NvtxRange range("finalize");   // prover.rs:115
    const size_t domain = cycles_ * INV_RATE;
    const FpExt poly_mix = iop_.random_ext();
    PolyGroup check;
    check.count = CHECK_SIZE;
    check.coeffs = DevBuf(c_, EXT * domain);
    {
      NvtxRange r2("eval_check");   // rv32im/src/prove/hal/cuda.rs:180
    desc_.eval_check(c_, check.coeffs.p, groups_[0].evaluated.p, groups_[1].evaluated.p, groups_[2].evaluated.p, out_host,
                     mix_host, poly_mix, (uint32_t)po2_);
    }
    r0_ntt_interpolate(c_, check.coeffs.p, EXT, (int)(po2_ + 2), /*zk=*/false, 0);
    finish_group(check);
    check.merkle.commit(c_, iop_, &roots);

    const FpExt z = iop_.random_ext();
    const FpExt back_one = ext_from_fp(R0_ROU_REV_MONT[po2_]);
    // evaluate every tap at z * back_one^back (prover.rs:176-211)
    const size_t ntaps = taps_.tap_size();
    std::vector<FpExt> all_xs(ntaps), eval_u(ntaps + CHECK_SIZE);
    std::vector<uint32_t> which(ntaps + CHECK_SIZE);
    for (size_t t = 0; t < ntaps; t++) {
      which[t] = taps_.taps[t].offset;
      all_xs[t] = ext_mul(ext_pow(back_one, taps_.taps[t].back), z);
    }
    const FpExt z_pow = ext_pow(z, EXT);
    for (size_t i = 0; i < CHECK_SIZE; i++) which[ntaps + i] = (uint32_t)i;
    {
      NvtxRange r2("eval_u");   // prover.rs:208
      std::vector<FpExt> xs(all_xs);
      xs.resize(ntaps + CHECK_SIZE, z_pow);
      DevBuf d_which(c_, which.size()), d_xs(c_, xs.size() * 4), d_out(c_, xs.size() * 4);
      R0_CUDA(cudaMemcpyAsync(d_which.p, which.data(), which.size() * 4, cudaMemcpyHostToDevice, c_->stream));
      R0_CUDA(cudaMemcpyAsync(d_xs.p, xs.data(), xs.size() * 16, cudaMemcpyHostToDevice, c_->stream));
      for (size_t g = 0; g <= groups_.size(); g++) {
        const size_t b = g < groups_.size() ? taps_.group_begin[g] : ntaps;
        const size_t e = g < groups_.size() ? taps_.group_begin[g + 1] : ntaps + CHECK_SIZE;
        const PolyGroup& pg = g < groups_.size() ? groups_[g] : check;
        r0_batch_evaluate_any(c_, pg.coeffs.p, cycles_, d_which.p + b, d_xs.p + 4 * b, d_out.p + 4 * b, e - b, pg.count, which.data() + b);
      }
      R0_CUDA(cudaMemcpyAsync(eval_u.data(), d_out.p, eval_u.size() * 16, cudaMemcpyDeviceToHost, c_->stream));
      R0_CUDA(cudaStreamSynchronize(c_->stream));
    }
    // coeff_u: per-register interpolation, then the 16 check evaluations verbatim (prover.rs:213-246)
    std::vector<FpExt> coeff_u(ntaps + CHECK_SIZE);
    const std::vector<size_t> regs = taps_.regs(0, ntaps);
    {
      NvtxRange r2("poly_interpolate");   // prover.rs:233
      for (size_t r : regs) poly_interpolate(&coeff_u[r], &all_xs[r], &eval_u[r], taps_.taps[r].skip);
    }
    for (size_t i = 0; i < CHECK_SIZE; i++) coeff_u[ntaps + i] = eval_u[ntaps + i];
    iop_.write((const uint32_t*)coeff_u.data(), coeff_u.size() * 4);
    iop_.commit(suite_.hash_words((const uint32_t*)coeff_u.data(), coeff_u.size() * 4));

    // DEEP combination (prover.rs:248-352)
    const FpExt mix = iop_.random_ext();
    const size_t combo_count = taps_.combos_count;
    DevBuf combos(c_, EXT * cycles_ * (combo_count + 1));
    r0_fill(c_, combos.p, 0, combos.words);
    FpExt cur_mix = ext_one();
    std::unique_ptr<NvtxRange> mix_range(new NvtxRange("mix_poly_coeffs"));   // prover.rs:284
    for (size_t g = 0; g < groups_.size(); g++) {
      const size_t gsize = taps_.group_sizes[g];
      std::vector<uint32_t> ids;
      for (size_t r : taps_.regs(taps_.group_begin[g], taps_.group_begin[g + 1])) ids.push_back(taps_.taps[r].combo);
      R0_CHECK(ids.size() == gsize, "tap set: register count != group size");
      r0_mix_poly_coeffs(c_, combos.p, cur_mix, mix, groups_[g].coeffs.p, ids.data(), gsize, cycles_);
      cur_mix = ext_mul(cur_mix, ext_pow(mix, gsize));
    }
    {
      std::vector<uint32_t> ids(CHECK_SIZE, (uint32_t)combo_count);
      r0_mix_poly_coeffs(c_, combos.p, cur_mix, mix, check.coeffs.p, ids.data(), CHECK_SIZE, cycles_);
    }
    mix_range.reset();
    {
      NvtxRange r2("load_combos");   // prover.rs:321 (prepare :325, divide :337)
      std::vector<uint32_t> reg_sizes, reg_ids;
      for (size_t r : regs) {
        reg_sizes.push_back(taps_.taps[r].skip);
        reg_ids.push_back(taps_.taps[r].combo);
      }
      r0_combos_prepare(c_, combos.p, coeff_u.data(), coeff_u.size(), (uint32_t)combo_count, cycles_, reg_sizes.data(),
                        reg_ids.data(), (uint32_t)reg_sizes.size(), mix, (uint32_t)CHECK_SIZE);
      std::vector<uint32_t> pow_begin{0};
      std::vector<uint32_t> pows;
      auto push = [&](const FpExt& e) { pows.insert(pows.end(), e.c, e.c + 4); };
      for (size_t i = 0; i < combo_count; i++) {
        for (size_t k = taps_.combo_begin[i]; k < taps_.combo_begin[i + 1]; k++)
          push(ext_mul(z, ext_pow(back_one, taps_.combo_taps[k])));
        pow_begin.push_back((uint32_t)(pows.size() / 4));
      }
      push(z_pow);
      pow_begin.push_back((uint32_t)(pows.size() / 4));
      const char* e = r0b200_combos_divide((r0b200_ctx*)c_, combos.p, combo_count + 1, pow_begin.data(), pows.data(), cycles_);
      if (e) {
        std::string msg(e);
        r0b200_free_error(e);
        throw std::runtime_error(msg);
      }
    }
    DevBuf final_coeffs(c_, EXT * cycles_);
    std::unique_ptr<NvtxRange> sum_range(new NvtxRange("sum"));   // prover.rs:358
    r0_eltwise_sum_ext(c_, final_coeffs.p, combos.p, cycles_, combo_count + 1);
    combos.release();
    r0_bit_reverse(c_, final_coeffs.p, EXT, (int)po2_);
    sum_range.reset();
    fri_prove(std::move(final_coeffs), check);
  }

 private:
  // PolyGroup::new (poly_group.rs:63-83): LDE, coefficient bit reversal, Merkle tree over the evaluations
  void finish_group(PolyGroup& pg) {
    NvtxRange range("poly_group");   // poly_group.rs:70
    const size_t domain = cycles_ * INV_RATE;
    pg.evaluated = DevBuf(c_, pg.count * domain);
    r0_ntt_expand_evaluate(c_, pg.evaluated.p, pg.coeffs.p, pg.count, (int)(po2_ + 2), 2, 0);
    r0_bit_reverse(c_, pg.coeffs.p, pg.count, (int)po2_);
    pg.merkle = MerkleTree(c_, hash_, pg.evaluated.p, domain, pg.count);
  }

  struct FriRound {
    size_t domain;
    DevBuf evaluated, coeffs;
    MerkleTree merkle;
  };

  void fri_prove(DevBuf&& in_coeffs, const PolyGroup& check) {  // fri.rs:77-126
    NvtxRange range("fri_prove");   // fri.rs:94
    const size_t orig_domain = in_coeffs.words / EXT * INV_RATE;
    std::vector<std::unique_ptr<FriRound>> rounds;
    DevBuf first = std::move(in_coeffs);
    const DevBuf* coeffs = &first;
    while (coeffs->words / EXT > FRI_MIN_DEGREE) {
      std::unique_ptr<FriRound> r(new FriRound());
      const size_t size = coeffs->words / EXT;
      r->domain = size * INV_RATE;
      r->evaluated = DevBuf(c_, r->domain * EXT);
      r0_ntt_expand_evaluate(c_, r->evaluated.p, coeffs->p, EXT, (int)log2_exact(r->domain), 2, 0);
      r->merkle = MerkleTree(c_, hash_, r->evaluated.p, r->domain / FRI_FOLD, FRI_FOLD * EXT);
      r->merkle.commit(c_, iop_, &roots);
      const FpExt fold_mix = iop_.random_ext();
      r->coeffs = DevBuf(c_, size / FRI_FOLD * EXT);
      r0_fri_fold(c_, r->coeffs.p, coeffs->p, size / FRI_FOLD, fold_mix);
      rounds.push_back(std::move(r));
      coeffs = &rounds.back()->coeffs;
    }
    {
      DevBuf fin(c_, coeffs->words);
      r0_eltwise_copy(c_, fin.p, coeffs->p, coeffs->words);
      r0_bit_reverse(c_, fin.p, EXT, (int)log2_exact(coeffs->words / EXT));
      std::vector<uint32_t> host(coeffs->words);
      R0_CUDA(cudaMemcpyAsync(host.data(), fin.p, host.size() * 4, cudaMemcpyDeviceToHost, c_->stream));
      R0_CUDA(cudaStreamSynchronize(c_->stream));
      iop_.write(host.data(), host.size());
      iop_.commit(suite_.hash_words(host.data(), host.size()));
    }
    // queries: positions first (RNG only), then every opening in two gathers
    Openings o;
    const unsigned bits = (unsigned)log2_exact(orig_domain);
    for (size_t q = 0; q < QUERIES; q++) {
      size_t pos = iop_.random_bits(bits);
      query_pos.push_back((uint32_t)pos);
      for (auto& g : groups_) g.merkle.prove(o, pos);
      check.merkle.prove(o, pos);
      for (auto& r : rounds) {
        size_t group = pos % (r->domain / FRI_FOLD);
        r->merkle.prove(o, group);
        pos = group;
      }
    }
    DevBuf d_vals(c_, o.val_words), d_digs(c_, o.djobs.size() * 8);
    r0_gather_batched(c_, d_vals.p, o.vjobs.data(), o.vjobs.size());
    r0_gather_digests(c_, d_digs.p, o.djobs.data(), o.djobs.size());
    std::vector<uint32_t> vals(o.val_words), digs(o.djobs.size() * 8);
    R0_CUDA(cudaMemcpyAsync(vals.data(), d_vals.p, vals.size() * 4, cudaMemcpyDeviceToHost, c_->stream));
    R0_CUDA(cudaMemcpyAsync(digs.data(), d_digs.p, digs.size() * 4, cudaMemcpyDeviceToHost, c_->stream));
    R0_CUDA(cudaStreamSynchronize(c_->stream));
    for (const Opening& op : o.list) {
      iop_.write(vals.data() + op.val_off, op.cols);
      iop_.write(digs.data() + 8 * op.dig_off, 8 * op.ndig);
    }
  }

  Ctx* c_;
  const CircuitDesc& desc_;
  int hash_;
  size_t po2_, cycles_;
  Transcript iop_;
  HostSuite suite_;
  TapSet taps_;
  std::vector<PolyGroup> groups_;
  std::vector<cudaEvent_t> uploads_[3];
};

}  // namespace
}  // namespace r0

using namespace r0;

// One proof in flight: prove_core split where the protocol splits it. begin() runs everything up to and including the
// draw of the accum mix (rv32im/src/prove/hal/mod.rs:181-213: version word, info and header commits, code and data
// groups, MIX_SIZE random elements); the caller computes accum FROM that mix (witgen.accum(.., &mix), :213-216) and
// finish() commits it and runs Prover::finalize (:217-222). The transcript lives in the handle between the two.
struct r0b200_proof {
  r0b200_ctx* ctx;
  const CircuitDesc& desc;
  uint32_t po2;
  SegmentProver prover;
  std::vector<uint32_t> header, mix;
  bool on_host = false, uploaded = false;
  r0b200_proof(r0b200_ctx* c, const CircuitDesc& d, int hash, uint32_t po2_)
      : ctx(c), desc(d), po2(po2_), prover(c, d, hash, po2_) {}
};

static r0b200_proof* proof_begin(r0b200_ctx* ctx, const CircuitDesc& desc, int hash, uint32_t po2, const uint32_t* code,
                                 const uint32_t* data, int witness_on_host, r0b200_witness* uploaded,
                                 const uint32_t* global_host) {
  R0_CHECK(ctx != nullptr, "null r0b200 context");
  R0_CUDA(cudaSetDevice(ctx->device));
  R0_CHECK(hash == R0B200_HASH_POSEIDON2 || hash == R0B200_HASH_SHA256, "prove: unknown hash suite");
  R0_CHECK(po2 >= 9 && po2 + 2 <= (uint32_t)MAX_LG, "prove: po2 out of range (9..22)");
  R0_CHECK(global_host != nullptr, "prove: null globals");
  std::unique_ptr<r0b200_proof> p(new r0b200_proof(ctx, desc, hash, po2));
  SegmentProver& prover = p->prover;
  Transcript& iop = prover.iop();
  HostSuite suite{hash};
  NvtxRange range("prove_begin");
  if (desc.has_version) iop.write(&desc.version, 1);
  auto commit_info = [&](const char* s) {
    uint32_t e[16];
    for (int i = 0; i < 16; i++) e[i] = fp_encode((uint32_t)(uint8_t)s[i]);
    iop.commit(suite.hash_words(e, 16));
  };
  commit_info("RISC0_STARK:v1__");
  commit_info(desc.info);
  // header: globals (INVALID -> 0) followed by the raw po2 word (rv32im/src/prove/hal/mod.rs:196-206,
  // recursion/src/prove/mod.rs:193-206)
  std::vector<uint32_t>& header = p->header;
  header.resize(desc.output_size + 1);
  for (size_t i = 0; i < desc.output_size; i++) header[i] = global_host[i] == FP_INVALID ? 0u : global_host[i];
  header[desc.output_size] = po2;
  iop.commit(suite.hash_words(header.data(), header.size()));
  iop.write(header.data(), header.size());
  p->on_host = witness_on_host != 0 || uploaded != nullptr;
  p->uploaded = uploaded != nullptr;
  if (uploaded) {
    R0_CHECK(uploaded->c == ctx && uploaded->po2 == po2, "uploaded witness belongs to another context or size");
    prover.adopt_group(1, uploaded);
    prover.adopt_group(2, uploaded);
    if (uploaded->coeffs[0]) prover.adopt_group(0, uploaded);   // legacy: accum uploaded before the mix was known
  } else if (p->on_host) {
    // uploads are enqueued up front in consumption order; each group's compute waits only for its own chunks
    prover.prefetch_group(1, code);
    prover.prefetch_group(2, data);
  }
  prover.commit_group(1, p->on_host ? nullptr : code);
  prover.commit_group(2, p->on_host ? nullptr : data);
  p->mix.resize(desc.mix_size);
  for (size_t i = 0; i < desc.mix_size; i++) p->mix[i] = iop.random_elem();
  return p.release();
}

static void proof_finish(r0b200_proof* p, const uint32_t* accum, int accum_on_host, uint32_t* seal_out_host,
                         size_t seal_cap, size_t* seal_len, uint32_t* roots_out_host, size_t roots_cap, size_t* nroots,
                         uint32_t* query_pos_out_host) {
  R0_CUDA(cudaSetDevice(p->ctx->device));
  SegmentProver& prover = p->prover;
  Transcript& iop = prover.iop();
  NvtxRange range("prove_finish");
  if (prover.has_group(0)) {
    prover.commit_group(0, nullptr);   // accum came with r0b200_witness_upload
  } else {
    R0_CHECK(accum != nullptr, "prove_finish: null accum");
    if (accum_on_host) prover.prefetch_group(0, accum);
    prover.commit_group(0, accum_on_host ? nullptr : accum);
  }
  prover.finalize(p->mix.data(), p->header.data());
  R0_CUDA(cudaStreamSynchronize(p->ctx->stream));
  if (seal_len) *seal_len = iop.proof.size();
  R0_CHECK(seal_out_host != nullptr && iop.proof.size() <= seal_cap, "prove: seal buffer too small");
  memcpy(seal_out_host, iop.proof.data(), iop.proof.size() * 4);
  if (nroots) *nroots = prover.roots.size();
  if (roots_out_host) {
    R0_CHECK(prover.roots.size() <= roots_cap, "prove: roots buffer too small");
    memcpy(roots_out_host, prover.roots.data(), prover.roots.size() * 32);
  }
  if (query_pos_out_host) memcpy(query_pos_out_host, prover.query_pos.data(), prover.query_pos.size() * 4);
}

static void prove_segment(r0b200_ctx* ctx, const CircuitDesc& desc, int hash, uint32_t po2, const uint32_t* code,
                          const uint32_t* data, const uint32_t* accum, int witness_on_host, r0b200_witness* uploaded,
                          const uint32_t* global_host,
                          uint32_t* seal_out_host, size_t seal_cap, size_t* seal_len, uint32_t* roots_out_host,
                          size_t roots_cap, size_t* nroots, uint32_t* query_pos_out_host) {
  std::unique_ptr<r0b200_proof> p(proof_begin(ctx, desc, hash, po2, code, data, witness_on_host, uploaded, global_host));
  proof_finish(p.get(), accum, witness_on_host, seal_out_host, seal_cap, seal_len, roots_out_host, roots_cap, nroots,
               query_pos_out_host);
}

static const CircuitDesc& circuit_desc(int circuit) {
  R0_CHECK(circuit == R0B200_CIRCUIT_RV32IM || circuit == R0B200_CIRCUIT_RECURSION, "unknown circuit");
  return circuit == R0B200_CIRCUIT_RV32IM ? kRv32im : kRecursion;
}

extern "C" r0b200_err r0b200_prove_begin(r0b200_ctx* ctx, int circuit, int hash, uint32_t po2, const uint32_t* code,
                                         const uint32_t* data, int witness_on_host, r0b200_witness* uploaded,
                                         const uint32_t* global_host, uint32_t* mix_out_host, size_t mix_cap,
                                         r0b200_proof** out) {
  R0_API_BEGIN
  R0_CHECK(out != nullptr && mix_out_host != nullptr, "prove_begin: null output");
  const CircuitDesc& desc = circuit_desc(circuit);
  R0_CHECK(mix_cap >= desc.mix_size, "prove_begin: mix buffer too small");
  if (uploaded) {
    R0_CHECK(uploaded->circuit == circuit, "prove_begin: uploaded witness is for another circuit");
    po2 = uploaded->po2;
  }
  r0b200_proof* p = proof_begin(ctx, desc, hash, po2, code, data, witness_on_host, uploaded, global_host);
  memcpy(mix_out_host, p->mix.data(), desc.mix_size * 4);
  *out = p;
  R0_API_END
}

extern "C" r0b200_err r0b200_prove_finish(r0b200_proof* proof, const uint32_t* accum, int accum_on_host,
                                          uint32_t* seal_out_host, size_t seal_cap, size_t* seal_len,
                                          uint32_t* roots_out_host, size_t roots_cap, size_t* nroots,
                                          uint32_t* query_pos_out_host) {
  R0_API_BEGIN
  R0_CHECK(proof != nullptr, "prove_finish: null proof");
  std::unique_ptr<r0b200_proof> p(proof);   // the handle is consumed, also on failure
  proof_finish(p.get(), accum, accum_on_host, seal_out_host, seal_cap, seal_len, roots_out_host, roots_cap, nroots,
               query_pos_out_host);
  R0_API_END
}

extern "C" void r0b200_prove_abort(r0b200_proof* proof) {
  if (!proof) return;
  cudaSetDevice(proof->ctx->device);
  delete proof;
}

// A segment's prover inputs resident on the device: the preflight trace (sorted), the injector CSR and the globals -
// what PreflightResults carries (prove/witgen/mod.rs:55-61). Uploaded on the copy stream so that segment s + 1 travels
// while segment s is proved (the reference's CPU -> GPU queue of depth 2, r0vm/src/actors/worker.rs:70-76,585).
struct r0b200_segment {
  r0b200_ctx* ctx = nullptr;
  uint32_t po2 = 0;
  r0b200_trace* trace = nullptr;
  uint32_t *d_index = nullptr, *d_offsets = nullptr, *d_values = nullptr;
  size_t index_len = 0, nvals = 0;
  std::vector<uint32_t> global;
  // bigint cycles of the trace (major 12), extracted at upload time: the BigIntAccumState cells of these rows depend on
  // the accum mix and are computed on the host between the two phases (prove/witgen/mod.rs:186-207, byte_poly.rs:403-475)
  struct BigIntCycle {
    uint32_t row, poly_op, coeff;
    uint8_t bytes[16];
  };
  std::vector<BigIntCycle> bigint;
  cudaEvent_t ready = nullptr;
};

namespace {
// RawPreflightCycle / RawMemoryTransaction as the host hands them over (rv32im-sys/src/lib.rs:21-61)
struct HostCycle {
  uint32_t state, pc;
  uint8_t major, minor, machine_mode, padding;
  uint32_t user_cycle, txn_idx, paging_idx, bigint_idx, diff_count[2];
};
struct HostTxn {
  uint32_t addr, cycle, word, prev_cycle, prev_word;
};
static_assert(sizeof(HostCycle) == 36 && sizeof(HostTxn) == 20, "preflight trace layout");
constexpr uint32_t kMajorBigInt = 12;      // 7 + CycleState::BigIntEcall (40) / 8; minor 0 = the ecall cycle, 1 = a step
constexpr uint32_t kPolyReset = 0, kPolyShift = 1, kPolySetTerm = 2, kPolyAddTotal = 3, kPolyCarry1 = 4, kPolyCarry2 = 5,
                   kPolyEqZero = 6;
// accum columns of BigIntAccumState: poly, term, total, four words each (kLayout_TopAccum.user._0.state; pinned by
// tests/test_preflight.py::test_bigint_accum_columns)
constexpr uint32_t kBigIntAccumCols[3] = {0, 4, 8};

// what a bigint row carries into the accum phase: the 16 witness bytes and, for a step cycle, poly_op / coeff of the
// verify-program word it executed - the first recorded transaction of the cycle (prove/witgen/bigint.rs:113-116)
void collect_bigint_cycles(const r0b200_preflight_trace* t, size_t cycles, std::vector<r0b200_segment::BigIntCycle>& out) {
  const HostCycle* cyc = static_cast<const HostCycle*>(t->cycles);
  const HostTxn* txn = static_cast<const HostTxn*>(t->txns);
  for (size_t r = 0; r < cycles; r++) {
    if (cyc[r].major != kMajorBigInt) continue;
    r0b200_segment::BigIntCycle b{};
    b.row = (uint32_t)r;
    R0_CHECK(t->bigint_bytes != nullptr && (size_t)cyc[r].bigint_idx + 16 <= t->bigint_bytes_len,
             "segment_upload: bigint cycle without its 16 witness bytes");
    memcpy(b.bytes, t->bigint_bytes + cyc[r].bigint_idx, 16);
    if (cyc[r].minor == 0) {
      b.poly_op = kPolyReset;
      b.coeff = 0;
    } else {
      R0_CHECK(cyc[r].txn_idx < t->txns_len, "segment_upload: bigint step cycle without transactions");
      const uint32_t insn = txn[cyc[r].txn_idx].word;
      b.poly_op = (insn >> 24) & 0xf;
      b.coeff = (insn >> 21) & 0x7;
      R0_CHECK(b.poly_op <= kPolyEqZero, "segment_upload: invalid poly_op in bigint program");
    }
    out.push_back(b);
  }
}

// BigIntAccum::step over the segment's bigint cycles at the mix point (Montgomery words; the last four of the accum
// mix): fills the scatter arrays for the 12 cells of each row. Throws when the program's identity does not hold.
void bigint_accum_cells(const std::vector<r0b200_segment::BigIntCycle>& rows, const uint32_t* mix, size_t mix_size, size_t cycles,
                        std::vector<uint32_t>& index, std::vector<uint32_t>& offsets, std::vector<uint32_t>& values) {
  const FpExt z{{mix[mix_size - 4], mix[mix_size - 3], mix[mix_size - 2], mix[mix_size - 1]}};
  FpExt powers[17];
  powers[0] = ext_one();
  for (int i = 1; i < 17; i++) powers[i] = ext_mul(powers[i - 1], z);
  FpExt neg_poly = ext_zero();
  for (int i = 0; i < 16; i++) neg_poly = ext_add(neg_poly, ext_scale(powers[i], fp_encode(128)));
  FpExt poly = ext_zero(), term = ext_one(), total = ext_zero();
  index.assign(1, 0u);
  for (const auto& b : rows) {
    FpExt delta = ext_zero();
    for (int i = 0; i < 16; i++) delta = ext_add(delta, ext_scale(powers[i], fp_encode(b.bytes[i])));
    const FpExt new_poly = ext_add(poly, delta);
    bool reset = false;
    switch (b.poly_op) {
      case kPolyReset: reset = true; break;
      case kPolyShift: poly = ext_mul(new_poly, powers[16]); break;
      case kPolySetTerm:
        poly = ext_zero();
        term = new_poly;
        break;
      case kPolyAddTotal:
        total = ext_add(total, ext_mul(ext_scale(term, fp_sub(fp_encode(b.coeff), fp_encode(4))), new_poly));
        poly = ext_zero();
        term = ext_one();
        break;
      case kPolyCarry1: poly = ext_add(poly, ext_scale(ext_sub(delta, neg_poly), fp_encode(64 * 256))); break;
      case kPolyCarry2: poly = ext_add(poly, ext_scale(delta, fp_encode(256))); break;
      default: {  // EqZero
        const FpExt goal = ext_add(total, ext_mul(new_poly, ext_sub(powers[1], ext_from_fp(fp_encode(256)))));
        R0_CHECK((goal.c[0] | goal.c[1] | goal.c[2] | goal.c[3]) == 0, "accum: Invalid eqz in bigint accum");
        reset = true;
      }
    }
    if (reset) {
      poly = ext_zero();
      term = ext_one();
      total = ext_zero();
    }
    const FpExt* regs[3] = {&poly, &term, &total};
    for (int k = 0; k < 3; k++)
      for (int j = 0; j < 4; j++) {
        offsets.push_back((uint32_t)((kBigIntAccumCols[k] + j) * cycles + b.row));
        values.push_back(regs[k]->c[j]);
      }
    index.push_back((uint32_t)offsets.size());
  }
}
}  // namespace

extern "C" void r0b200_segment_free(r0b200_segment* seg) {
  if (!seg) return;
  cudaSetDevice(seg->ctx->device);
  if (seg->ready) cudaEventDestroy(seg->ready);
  r0_trace_free(seg->trace);
  for (uint32_t* p : {seg->d_index, seg->d_offsets, seg->d_values})
    if (p) cudaFreeAsync(p, seg->ctx->stream);
  delete seg;
}

extern "C" r0b200_err r0b200_segment_upload(r0b200_ctx* ctx, uint32_t po2, const r0b200_preflight_trace* trace_host,
                                            const uint32_t* global_host, const uint32_t* inj_index_host,
                                            size_t inj_index_len, const uint32_t* inj_offsets_host,
                                            const uint32_t* inj_values_host, r0b200_segment** out) {
  R0_API_BEGIN
  R0_CHECK(ctx != nullptr && trace_host != nullptr && global_host != nullptr && out != nullptr, "segment_upload: null argument");
  R0_CHECK(po2 >= 9 && po2 + 2 <= (uint32_t)MAX_LG, "segment_upload: po2 out of range (9..22)");
  R0_CUDA(cudaSetDevice(ctx->device));
  const size_t cycles = size_t(1) << po2;
  R0_CHECK(inj_index_len == 0 || inj_index_len == cycles + 1, "segment_upload: injector index must have cycles + 1 entries");
  struct Guard {
    r0b200_segment* s;
    ~Guard() { r0b200_segment_free(s); }
  } g{new r0b200_segment()};
  r0b200_segment* seg = g.s;
  seg->ctx = ctx;
  seg->po2 = po2;
  seg->global.assign(global_host, global_host + kRv32im.output_size);
  collect_bigint_cycles(trace_host, cycles, seg->bigint);
  cudaStream_t cs = ctx->copy_stream;
  seg->trace = r0_trace_upload(ctx, trace_host, (uint32_t)cycles, cs);
  if (inj_index_len >= 2) {
    seg->index_len = inj_index_len;
    seg->nvals = inj_index_host[inj_index_len - 1];
    R0_CUDA(r0_malloc_async(seg->ctx, &seg->d_index, inj_index_len * 4, cs));
    R0_CUDA(r0_malloc_async(seg->ctx, &seg->d_offsets, std::max<size_t>(1, seg->nvals) * 4, cs));
    R0_CUDA(r0_malloc_async(seg->ctx, &seg->d_values, std::max<size_t>(1, seg->nvals) * 4, cs));
    R0_CUDA(cudaMemcpyAsync(seg->d_index, inj_index_host, inj_index_len * 4, cudaMemcpyHostToDevice, cs));
    R0_CUDA(cudaMemcpyAsync(seg->d_offsets, inj_offsets_host, seg->nvals * 4, cudaMemcpyHostToDevice, cs));
    R0_CUDA(cudaMemcpyAsync(seg->d_values, inj_values_host, seg->nvals * 4, cudaMemcpyHostToDevice, cs));
  }
  R0_CUDA(cudaEventCreateWithFlags(&seg->ready, cudaEventDisableTiming));
  R0_CUDA(cudaEventRecord(seg->ready, cs));
  g.s = nullptr;
  *out = seg;
  R0_API_END
}

// prove_core from a PreflightResults (rv32im/src/prove/hal/mod.rs:143-224): WitnessGenerator::new on the device
// (prove/witgen/mod.rs:130-176: INVALID fill, injector scatter, generate_witness, zeroize), the two group commits, the
// mix draw, WitnessGenerator::accum (:178-224: step_accum from that mix, zeroize), the accum commit and finalize.
static void prove_core_rv32im(r0b200_ctx* ctx, int hash, r0b200_segment* seg, uint32_t* seal_out_host, size_t seal_cap,
                              size_t* seal_len, uint32_t* roots_out_host, size_t roots_cap, size_t* nroots,
                              uint32_t* query_pos_out_host, uint32_t* global_out_host) {
  NvtxRange range("prove_core");
  const uint32_t po2 = seg->po2;
  const size_t cycles = size_t(1) << po2;
  const CircuitDesc& d = kRv32im;
  R0_CUDA(cudaStreamWaitEvent(ctx->stream, seg->ready, 0));
  DevBuf data, code, accum, d_global(ctx, d.output_size), d_mix(ctx, d.mix_size);
  std::vector<uint32_t> global(d.output_size);
  {
    NvtxRange r2("witness_generator_new");
    data = DevBuf(ctx, d.group_sizes[2] * cycles);
    r0_fill(ctx, data.p, FP_INVALID, data.words);
    if (seg->index_len >= 2) r0_scatter_dev(ctx, data.p, seg->d_index, seg->index_len - 1, seg->d_offsets, seg->d_values);
    R0_CUDA(cudaMemcpyAsync(d_global.p, seg->global.data(), d.output_size * 4, cudaMemcpyHostToDevice, ctx->stream));
    {
      NvtxRange r3("witgen");
      r0_witgen_rv32im(ctx, seg->trace, d_global.p, data.p, /*sync_check=*/true);
    }
    NvtxRange r4("zeroize");
    r0_eltwise_zeroize(ctx, d_global.p, d.output_size);
    r0_eltwise_zeroize(ctx, data.p, data.words);
    code = DevBuf(ctx, d.group_sizes[1] * cycles);
    r0_fill(ctx, code.p, 0, code.words);   // code is INVALID-initialised then zeroized: all zero (witgen/mod.rs:149,168)
    R0_CUDA(cudaMemcpyAsync(global.data(), d_global.p, d.output_size * 4, cudaMemcpyDeviceToHost, ctx->stream));
    R0_CUDA(cudaStreamSynchronize(ctx->stream));
  }
  if (global_out_host) memcpy(global_out_host, global.data(), d.output_size * 4);
  std::unique_ptr<r0b200_proof> p(proof_begin(ctx, d, hash, po2, code.p, data.p, 0, nullptr, global.data()));
  {
    NvtxRange r2("accumulate");
    accum = DevBuf(ctx, d.group_sizes[0] * cycles);
    r0_fill(ctx, accum.p, FP_INVALID, accum.words);
    DevBuf bi_index, bi_offsets, bi_values;
    std::vector<uint32_t> h_index, h_offsets, h_values;   // outlive the copies below (pageable: staged before return)
    if (!seg->bigint.empty()) {
      // inject BigIntAccumState backs (witgen/mod.rs:186-207): a function of the mix the transcript has just yielded
      bigint_accum_cells(seg->bigint, p->mix.data(), d.mix_size, cycles, h_index, h_offsets, h_values);
      bi_index = DevBuf(ctx, h_index.size());
      bi_offsets = DevBuf(ctx, h_offsets.size());
      bi_values = DevBuf(ctx, h_values.size());
      R0_CUDA(cudaMemcpyAsync(bi_index.p, h_index.data(), h_index.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
      R0_CUDA(cudaMemcpyAsync(bi_offsets.p, h_offsets.data(), h_offsets.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
      R0_CUDA(cudaMemcpyAsync(bi_values.p, h_values.data(), h_values.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
      r0_scatter_dev(ctx, accum.p, bi_index.p, h_index.size() - 1, bi_offsets.p, bi_values.p);
    }
    R0_CUDA(cudaMemcpyAsync(d_mix.p, p->mix.data(), d.mix_size * 4, cudaMemcpyHostToDevice, ctx->stream));
    r0_accum_rv32im(ctx, seg->trace, data.p, accum.p, d_global.p, d_mix.p, /*sync_check=*/true);
    r0_eltwise_zeroize(ctx, accum.p, accum.words);
  }
  proof_finish(p.get(), accum.p, 0, seal_out_host, seal_cap, seal_len, roots_out_host, roots_cap, nroots,
               query_pos_out_host);
}

extern "C" r0b200_err r0b200_prove_segment(r0b200_ctx* ctx, int hash, r0b200_segment* segment, uint32_t* seal_out_host,
                                           size_t seal_cap, size_t* seal_len, uint32_t* roots_out_host, size_t roots_cap,
                                           size_t* nroots, uint32_t* query_pos_out_host, uint32_t* global_out_host) {
  R0_API_BEGIN
  R0_CHECK(ctx != nullptr && segment != nullptr && segment->ctx == ctx, "prove_segment: segment belongs to another context");
  R0_CUDA(cudaSetDevice(ctx->device));
  prove_core_rv32im(ctx, hash, segment, seal_out_host, seal_cap, seal_len, roots_out_host, roots_cap, nroots,
                    query_pos_out_host, global_out_host);
  R0_API_END
}

extern "C" r0b200_err r0b200_prove_segment_rv32im(r0b200_ctx* ctx, int hash, uint32_t po2,
                                                  const r0b200_preflight_trace* trace_host, const uint32_t* global_host,
                                                  const uint32_t* inj_index_host, size_t inj_index_len,
                                                  const uint32_t* inj_offsets_host, const uint32_t* inj_values_host,
                                                  uint32_t* seal_out_host, size_t seal_cap, size_t* seal_len,
                                                  uint32_t* roots_out_host, size_t roots_cap, size_t* nroots,
                                                  uint32_t* query_pos_out_host, uint32_t* global_out_host) {
  r0b200_segment* seg = nullptr;
  r0b200_err e = r0b200_segment_upload(ctx, po2, trace_host, global_host, inj_index_host, inj_index_len, inj_offsets_host,
                                       inj_values_host, &seg);
  if (e) return e;
  e = r0b200_prove_segment(ctx, hash, seg, seal_out_host, seal_cap, seal_len, roots_out_host, roots_cap, nroots,
                           query_pos_out_host, global_out_host);
  r0b200_segment_free(seg);
  return e;
}

extern "C" r0b200_err r0b200_prove_rv32im(r0b200_ctx* ctx, int hash, uint32_t po2, const uint32_t* code,
                                          const uint32_t* data, const uint32_t* accum, int witness_on_host,
                                          const uint32_t* global_host, uint32_t* seal_out_host, size_t seal_cap,
                                          size_t* seal_len, uint32_t* roots_out_host, size_t roots_cap, size_t* nroots,
                                          uint32_t* query_pos_out_host) {
  R0_API_BEGIN
  prove_segment(ctx, kRv32im, hash, po2, code, data, accum, witness_on_host, nullptr, global_host, seal_out_host, seal_cap,
                seal_len, roots_out_host, roots_cap, nroots, query_pos_out_host);
  R0_API_END
}

extern "C" r0b200_err r0b200_prove_recursion(r0b200_ctx* ctx, int hash, uint32_t po2, const uint32_t* ctrl,
                                             const uint32_t* data, const uint32_t* accum, int witness_on_host,
                                             const uint32_t* global_host, uint32_t* seal_out_host, size_t seal_cap,
                                             size_t* seal_len, uint32_t* roots_out_host, size_t roots_cap,
                                             size_t* nroots, uint32_t* query_pos_out_host) {
  R0_API_BEGIN
  prove_segment(ctx, kRecursion, hash, po2, ctrl, data, accum, witness_on_host, nullptr, global_host, seal_out_host,
                seal_cap, seal_len, roots_out_host, roots_cap, nroots, query_pos_out_host);
  R0_API_END
}

// ---- pipelined host witnesses: upload segment s+1 while segment s is being proved ---------------------------------
extern "C" r0b200_err r0b200_witness_upload(r0b200_ctx* ctx, int circuit, uint32_t po2, const uint32_t* code_host,
                                            const uint32_t* data_host, const uint32_t* accum_host, r0b200_witness** out) {
  R0_API_BEGIN
  R0_CHECK(ctx != nullptr && out != nullptr, "witness_upload: null argument");
  R0_CHECK(circuit == 0 || circuit == 1, "witness_upload: unknown circuit");
  R0_CHECK(po2 >= 9 && po2 + 2 <= (uint32_t)MAX_LG, "witness_upload: po2 out of range");
  R0_CUDA(cudaSetDevice(ctx->device));
  const CircuitDesc& desc = circuit == 0 ? kRv32im : kRecursion;
  const size_t cycles = size_t(1) << po2;
  std::unique_ptr<r0b200_witness> w(new r0b200_witness());
  w->c = ctx;
  w->po2 = po2;
  w->circuit = circuit;
  const uint32_t* src[3] = {accum_host, code_host, data_host};   // tap-group order: accum, code, data
  size_t chunk = (size_t(64) << 20) / (cycles * 4);
  if (chunk == 0) chunk = 1;
  const int order[3] = {1, 2, 0};                                // consumption order: code, data, accum
  R0_CHECK(code_host != nullptr && data_host != nullptr, "witness_upload: null code / data");
  for (int oi = 0; oi < 3; oi++) {
    const int g = order[oi];
    if (!src[g]) continue;   // accum is normally not known yet (it depends on the mix drawn by prove_begin)
    const size_t count = desc.group_sizes[g];
    w->words[g] = count * cycles;
    R0_CUDA(r0_malloc_async(ctx, &w->coeffs[g], w->words[g] * 4, ctx->copy_stream));
    for (size_t c0 = 0; c0 < count; c0 += chunk) {
      const size_t nc = std::min(chunk, count - c0);
      R0_CUDA(cudaMemcpyAsync(w->coeffs[g] + c0 * cycles, src[g] + c0 * cycles, nc * cycles * 4, cudaMemcpyHostToDevice,
                              ctx->copy_stream));
      cudaEvent_t ev;
      R0_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
      R0_CUDA(cudaEventRecord(ev, ctx->copy_stream));
      w->events[g].push_back(ev);
    }
  }
  *out = w.release();
  R0_API_END
}

extern "C" void r0b200_witness_free(r0b200_witness* w) {
  if (!w) return;
  cudaSetDevice(w->c->device);
  for (int g = 0; g < 3; g++) {
    for (cudaEvent_t ev : w->events[g]) cudaEventDestroy(ev);
    if (w->coeffs[g]) cudaFreeAsync(w->coeffs[g], w->c->copy_stream);
  }
  delete w;
}

extern "C" r0b200_err r0b200_prove_uploaded(r0b200_ctx* ctx, int hash, r0b200_witness* witness, const uint32_t* global_host,
                                            uint32_t* seal_out_host, size_t seal_cap, size_t* seal_len,
                                            uint32_t* roots_out_host, size_t roots_cap, size_t* nroots,
                                            uint32_t* query_pos_out_host) {
  R0_API_BEGIN
  R0_CHECK(witness != nullptr, "prove_uploaded: null witness");
  prove_segment(ctx, witness->circuit == 0 ? kRv32im : kRecursion, hash, witness->po2, nullptr, nullptr, nullptr, 1, witness,
                global_host, seal_out_host, seal_cap, seal_len, roots_out_host, roots_cap, nroots, query_pos_out_host);
  R0_API_END
}
