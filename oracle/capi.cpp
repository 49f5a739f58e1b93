// ORACLE — TEST INFRASTRUCTURE ONLY. extern "C" surface of the CPU restatement, for ctypes (tests/, smoke(),
// bench.py's cpu_baseline leg). The product (`risc0_b200/`) never links, loads or calls this library.
//
// Error convention mirrors the reference FFI (risc0/sys/src/lib.rs:53-75): NULL = OK, else a strdup'd message the
// caller frees with `orc_free_str`.
#include <dlfcn.h>
#include <omp.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#include "hal_cpu.h"
#include "prover.h"
#include "tables/circuit_recursion.h"
#include "tables/circuit_rv32im.h"
#include "verifier.h"

using namespace oracle;

#define ORC_TRY try {
#define ORC_CATCH                        \
  }                                      \
  catch (const std::exception& e) {      \
    return strdup(e.what());             \
  }                                      \
  return nullptr;

static HashSuite suite_of(int kind) { return HashSuite(kind == 0 ? HashKind::Poseidon2 : HashKind::Sha256); }

// ------------------------------------------------------------------ reference-compiled poly_fp (oracle/_ref)
// risc0_circuit_rv32im_cpu_poly_fp: rv32im-sys/kernels/cxx/eval_check.cpp:30-38
typedef const char* (*rv32im_poly_fp_t)(size_t cycle, size_t steps, FpExt* poly_mix, Fp** args, FpExt* result);
static rv32im_poly_fp_t g_rv32im_poly_fp = nullptr;

extern "C" const char* orc_load_ref(const char* path) {
  ORC_TRY
  void* h = dlopen(path, RTLD_NOW | RTLD_LOCAL);
  if (!h) throw std::runtime_error(std::string("dlopen failed: ") + dlerror());
  g_rv32im_poly_fp = (rv32im_poly_fp_t)dlsym(h, "risc0_circuit_rv32im_cpu_poly_fp");
  if (!g_rv32im_poly_fp) throw std::runtime_error("symbol risc0_circuit_rv32im_cpu_poly_fp not found");
  ORC_CATCH
}
extern "C" int orc_ref_loaded() { return g_rv32im_poly_fp != nullptr; }

static TapSet rv32im_taps() {
  return TapSet::from_tables(RV32IM_TAPS, RV32IM_NUM_TAPS, RV32IM_COMBO_TAPS, RV32IM_TOT_COMBO_BACKS,
                             RV32IM_COMBO_BEGIN, RV32IM_NUM_COMBOS, RV32IM_GROUP_BEGIN, 3, RV32IM_NUM_REGS);
}
static TapSet hello_taps() {
  // verify/mod.rs:627-660
  static const uint16_t flat[15] = {0, 0, 0, 0, 1, 0, 0, 1, 0, 1, 0, 0, 2, 0, 1};
  static const uint16_t ctaps[1] = {0};
  static const uint16_t cbegin[2] = {0, 1};
  static const uint32_t gbegin[4] = {0, 1, 2, 3};
  return TapSet::from_tables(flat, 3, ctaps, 1, cbegin, 1, gbegin, 3, 3);
}

// rv32im CpuCircuitHal::eval_check, rv32im/src/prove/hal/cpu.rs:145-208. groups = [accum, code, data];
// globals = [mix, out] (GLOBAL_MIX = 0, GLOBAL_OUT = 1 as passed by prove_core: finalize(&[&mix.buf, global])).
static void rv32im_eval_check(Fp* check, const Fp* accum, const Fp* data, const Fp* mix, const Fp* out, FpExt poly_mix,
                              size_t po2, size_t steps, size_t begin, size_t end) {
  if (!g_rv32im_poly_fp) throw std::runtime_error("oracle/_ref poly_fp not loaded (call orc_load_ref)");
  size_t domain = steps * INV_RATE;
  std::vector<FpExt> pows(RV32IM_NUM_POLY_MIX_POWERS);
  for (size_t i = 0; i < pows.size(); i++) pows[i] = poly_mix.pow(RV32IM_POLY_MIX_POWERS[i]);
  Fp rou = rou_fwd(unsigned(po2 + 2));
  std::string err;
#pragma omp parallel for schedule(dynamic, 64)
  for (size_t cycle = begin; cycle < end; cycle++) {
    Fp* args[4] = {const_cast<Fp*>(accum), const_cast<Fp*>(data), const_cast<Fp*>(out), const_cast<Fp*>(mix)};
    FpExt tot;
    const char* e = g_rv32im_poly_fp(cycle, domain, pows.data(), args, &tot);
    if (e) {
#pragma omp critical
      err = e;
      continue;
    }
    Fp x = rou.pow(cycle);
    Fp y = (Fp(3) * x).pow(size_t(1) << po2);
    FpExt ret = tot * (y - Fp(1)).inv();
    for (size_t k = 0; k < EXT_SIZE; k++) check[k * domain + cycle] = ret.e[k];
  }
  if (!err.empty()) throw std::runtime_error(err);
}

// Range form so tests / the CPU baseline can run a bounded sample of domain points [begin, end).
extern "C" const char* orc_rv32im_eval_check(uint32_t* check, const uint32_t* accum, const uint32_t* data,
                                             const uint32_t* mix, const uint32_t* out, const uint32_t* poly_mix,
                                             uint32_t po2, uint64_t begin, uint64_t end) {
  ORC_TRY
  FpExt pm(Fp::raw(poly_mix[0]), Fp::raw(poly_mix[1]), Fp::raw(poly_mix[2]), Fp::raw(poly_mix[3]));
  rv32im_eval_check((Fp*)check, (const Fp*)accum, (const Fp*)data, (const Fp*)mix, (const Fp*)out, pm, po2,
                    size_t(1) << po2, begin, end);
  ORC_CATCH
}

// Constraint check of a RAW (un-extended) rv32im witness: is every constraint zero on every row? The generated poly_fp
// reads a tap at row (cycle - INV_RATE * back) mod steps (rust_poly_fp_0.cpp:22,114), so the N trace rows are placed at
// every INV_RATE-th row of a steps = INV_RATE * N matrix and poly_fp is evaluated at those rows only. This is what "the
// witness satisfies the circuit" means; the prover itself never checks it. bad_rows = rows with a non-zero result.
extern "C" const char* orc_rv32im_check_constraints(const uint32_t* accum, const uint32_t* data, const uint32_t* mix,
                                                    const uint32_t* out, const uint32_t* poly_mix, uint32_t po2,
                                                    uint64_t* bad_rows, uint64_t* first_bad) {
  ORC_TRY
  if (!g_rv32im_poly_fp) throw std::runtime_error("oracle/_ref poly_fp not loaded (call orc_load_ref)");
  const size_t N = size_t(1) << po2, steps = N * INV_RATE;
  FpExt pm(Fp::raw(poly_mix[0]), Fp::raw(poly_mix[1]), Fp::raw(poly_mix[2]), Fp::raw(poly_mix[3]));
  std::vector<FpExt> pows(RV32IM_NUM_POLY_MIX_POWERS);
  for (size_t i = 0; i < pows.size(); i++) pows[i] = pm.pow(RV32IM_POLY_MIX_POWERS[i]);
  std::vector<Fp> acc4(103 * steps), dat4(211 * steps);
  for (size_t c = 0; c < 103; c++)
    for (size_t r = 0; r < N; r++) acc4[c * steps + INV_RATE * r] = Fp::raw(accum[c * N + r]);
  for (size_t c = 0; c < 211; c++)
    for (size_t r = 0; r < N; r++) dat4[c * steps + INV_RATE * r] = Fp::raw(data[c * N + r]);
  uint64_t bad = 0, first = UINT64_MAX;
  std::string err;
#pragma omp parallel for schedule(dynamic, 64) reduction(+ : bad) reduction(min : first)
  for (size_t r = 0; r < N; r++) {
    Fp* args[4] = {acc4.data(), dat4.data(), (Fp*)out, (Fp*)mix};
    FpExt tot;
    const char* e = g_rv32im_poly_fp(INV_RATE * r, steps, pows.data(), args, &tot);
    if (e) {
#pragma omp critical
      err = e;
      continue;
    }
    bool zero = true;
    for (size_t k = 0; k < EXT_SIZE; k++) zero = zero && tot.e[k].v == 0;
    if (!zero) {
      bad++;
      if (r < first) first = r;
    }
  }
  if (!err.empty()) throw std::runtime_error(err);
  *bad_rows = bad;
  *first_bad = first;
  ORC_CATCH
}

// HelloCircuit constraints (verify/mod.rs:683-704): u0 = 0, u1 = 0, u2*(u2-1) = 0
static void hello_eval_check(Fp* check, const std::vector<const Fp*>& g, FpExt poly_mix, size_t po2, size_t steps) {
  size_t domain = steps * INV_RATE;
  Fp rou = rou_fwd(unsigned(po2 + 2));
  for (size_t i = 0; i < domain; i++) {
    Fp u0 = g[0][i], u1 = g[1][i], u2 = g[2][i];
    FpExt tot, mul = FpExt::one();
    tot += mul * u0;
    mul *= poly_mix;
    tot += mul * u1;
    mul *= poly_mix;
    tot += mul * (u2 * (u2 - Fp(1)));
    Fp x = rou.pow(i);
    Fp y = (Fp(3) * x).pow(steps);
    FpExt ret = tot * (y - Fp(1)).inv();
    for (size_t k = 0; k < EXT_SIZE; k++) check[k * domain + i] = ret.e[k];
  }
}
static FpExt hello_poly_ext(FpExt mix, const std::vector<FpExt>& u, const std::vector<std::vector<Fp>>&) {
  FpExt tot, mul = FpExt::one();
  tot += mul * u[0];
  mul *= mix;
  tot += mul * u[1];
  mul *= mix;
  tot += mul * (u[2] * (u[2] - FpExt::one()));
  return tot;
}

// ------------------------------------------------------------------ field / hash KAT helpers
extern "C" uint32_t orc_fp_encode(uint32_t x) { return Fp(x).v; }
extern "C" uint32_t orc_fp_decode(uint32_t m) { return Fp::raw(m).as_u32(); }
extern "C" uint32_t orc_fp_add(uint32_t a, uint32_t b) { return (Fp::raw(a) + Fp::raw(b)).v; }
extern "C" uint32_t orc_fp_sub(uint32_t a, uint32_t b) { return (Fp::raw(a) - Fp::raw(b)).v; }
extern "C" uint32_t orc_fp_mul(uint32_t a, uint32_t b) { return (Fp::raw(a) * Fp::raw(b)).v; }
extern "C" uint32_t orc_fp_pow(uint32_t a, uint64_t n) { return Fp::raw(a).pow(n).v; }
extern "C" uint32_t orc_fp_inv(uint32_t a) { return Fp::raw(a).inv().v; }
static FpExt ext_of(const uint32_t* p) { return FpExt(Fp::raw(p[0]), Fp::raw(p[1]), Fp::raw(p[2]), Fp::raw(p[3])); }
static void ext_to(uint32_t* p, const FpExt& x) {
  for (int i = 0; i < 4; i++) p[i] = x.e[i].v;
}
extern "C" void orc_fpext_mul(uint32_t* out, const uint32_t* a, const uint32_t* b) { ext_to(out, ext_of(a) * ext_of(b)); }
extern "C" void orc_fpext_inv(uint32_t* out, const uint32_t* a) { ext_to(out, ext_of(a).inv()); }
extern "C" void orc_fpext_pow(uint32_t* out, const uint32_t* a, uint64_t n) { ext_to(out, ext_of(a).pow(n)); }
extern "C" uint32_t orc_rou_fwd(uint32_t k) { return rou_fwd(k).v; }
extern "C" uint32_t orc_rou_rev(uint32_t k) { return rou_rev(k).v; }

extern "C" void orc_poseidon2_mix(uint32_t* cells_mont) {
  P2State st;
  for (int i = 0; i < 24; i++) st[i] = Fp::raw(cells_mont[i]);
  poseidon2_mix(st);
  for (int i = 0; i < 24; i++) cells_mont[i] = st[i].v;
}
extern "C" void orc_hash_elems(int kind, uint32_t* out8, const uint32_t* data, uint64_t n, uint64_t stride) {
  Digest d = suite_of(kind).hash_elem_slice((const Fp*)data, n, stride);
  memcpy(out8, d.w, 32);
}
extern "C" void orc_hash_pair(int kind, uint32_t* out8, const uint32_t* a, const uint32_t* b) {
  Digest da, db;
  memcpy(da.w, a, 32);
  memcpy(db.w, b, 32);
  Digest d = suite_of(kind).hash_pair(da, db);
  memcpy(out8, d.w, 32);
}
extern "C" void orc_sha_hash_bytes(uint32_t* out8, const uint8_t* bytes, uint64_t n) {
  Digest d = sha_hash_bytes(bytes, n);
  memcpy(out8, d.w, 32);
}
extern "C" void* orc_rng_new(int kind) { return suite_of(kind).new_rng().release(); }
extern "C" void orc_rng_free(void* r) { delete (Rng*)r; }
extern "C" void orc_rng_mix(void* r, const uint32_t* d8) {
  Digest d;
  memcpy(d.w, d8, 32);
  ((Rng*)r)->mix(d);
}
extern "C" uint32_t orc_rng_elem(void* r) { return ((Rng*)r)->random_elem().v; }
extern "C" uint32_t orc_rng_bits(void* r, uint32_t bits) { return ((Rng*)r)->random_bits(bits); }

extern "C" void orc_merkle_params(uint64_t rows, uint64_t cols, uint64_t queries, uint64_t* layers, uint64_t* top_layer,
                                  uint64_t* top_size) {
  MerkleTreeParams p(rows, cols, queries);
  *layers = p.layers;
  *top_layer = p.top_layer;
  *top_size = p.top_size;
}

// ------------------------------------------------------------------ Hal ops (names = Hal trait method names)
extern "C" const char* orc_batch_expand_into_evaluate_ntt(uint32_t* out, uint64_t out_size, const uint32_t* in,
                                                          uint64_t in_size, uint64_t count, uint32_t expand_bits) {
  ORC_TRY batch_expand_into_evaluate_ntt((Fp*)out, out_size, (const Fp*)in, in_size, count, expand_bits);
  ORC_CATCH
}
extern "C" void orc_batch_interpolate_ntt(uint32_t* io, uint64_t size, uint64_t count) {
  batch_interpolate_ntt((Fp*)io, size, count);
}
extern "C" void orc_batch_bit_reverse(uint32_t* io, uint64_t size, uint64_t count) {
  batch_bit_reverse((Fp*)io, size, count);
}
extern "C" void orc_batch_evaluate_any(const uint32_t* coeffs, uint64_t coeffs_size, uint64_t poly_count,
                                       const uint32_t* which, const uint32_t* xs, uint32_t* out, uint64_t eval_count) {
  batch_evaluate_any((const Fp*)coeffs, coeffs_size, poly_count, which, (const FpExt*)xs, (FpExt*)out, eval_count);
}
extern "C" void orc_zk_shift(uint32_t* io, uint64_t size, uint64_t count) { zk_shift((Fp*)io, size, count); }
extern "C" void orc_mix_poly_coeffs(uint32_t* out, uint64_t out_size, const uint32_t* mix_start, const uint32_t* mix,
                                    const uint32_t* in, const uint32_t* combos, uint64_t input_size, uint64_t count) {
  mix_poly_coeffs((FpExt*)out, out_size, ext_of(mix_start), ext_of(mix), (const Fp*)in, combos, input_size, count);
}
extern "C" void orc_eltwise_add_elem(uint32_t* out, const uint32_t* a, const uint32_t* b, uint64_t n) {
  eltwise_add_elem((Fp*)out, (const Fp*)a, (const Fp*)b, n);
}
extern "C" void orc_eltwise_sum_extelem(uint32_t* out, uint64_t out_size, const uint32_t* in, uint64_t in_size) {
  eltwise_sum_extelem((Fp*)out, out_size, (const FpExt*)in, in_size);
}
extern "C" void orc_eltwise_copy_elem(uint32_t* out, const uint32_t* in, uint64_t n) {
  eltwise_copy_elem((Fp*)out, (const Fp*)in, n);
}
extern "C" void orc_eltwise_zeroize_elem(uint32_t* io, uint64_t n) { eltwise_zeroize_elem((Fp*)io, n); }
extern "C" void orc_fri_fold(uint32_t* out, uint64_t out_size, const uint32_t* in, const uint32_t* mix) {
  fri_fold((Fp*)out, out_size, (const Fp*)in, ext_of(mix));
}
extern "C" void orc_hash_rows(int kind, uint32_t* out, uint64_t rows, const uint32_t* matrix, uint64_t matrix_size) {
  hash_rows(suite_of(kind), (Digest*)out, rows, (const Fp*)matrix, matrix_size);
}
extern "C" const char* orc_hash_fold(int kind, uint32_t* io, uint64_t input_size, uint64_t output_size) {
  ORC_TRY hash_fold(suite_of(kind), (Digest*)io, input_size, output_size);
  ORC_CATCH
}
extern "C" void orc_gather_sample(uint32_t* dst, const uint32_t* src, uint64_t idx, uint64_t size, uint64_t stride) {
  gather_sample((Fp*)dst, (const Fp*)src, idx, size, stride);
}
extern "C" void orc_scatter(uint32_t* into, const uint32_t* index, uint64_t index_len, const uint32_t* offsets,
                            const uint32_t* values) {
  scatter((Fp*)into, index, index_len, offsets, (const Fp*)values);
}
extern "C" void orc_eltwise_copy_elem_slice(uint32_t* into, const uint32_t* from, uint64_t from_rows,
                                            uint64_t from_cols, uint64_t from_offset, uint64_t from_stride,
                                            uint64_t into_offset, uint64_t into_stride) {
  eltwise_copy_elem_slice((Fp*)into, (const Fp*)from, from_rows, from_cols, from_offset, from_stride, into_offset,
                          into_stride);
}
extern "C" void orc_prefix_products(uint32_t* io, uint64_t n) { prefix_products((FpExt*)io, n); }
extern "C" void orc_combos_prepare(uint32_t* combos, const uint32_t* coeff_u, uint64_t combo_count, uint64_t cycles,
                                   const uint32_t* reg_sizes, const uint32_t* reg_combo_ids, uint64_t nregs,
                                   const uint32_t* mix) {
  combos_prepare((FpExt*)combos, (const FpExt*)coeff_u, combo_count, cycles, reg_sizes, reg_combo_ids, nregs,
                 ext_of(mix));
}
// chunk i of `cycles` FpExt is divided by pows[pow_begin[i] .. pow_begin[i+1]); returns nonzero-remainder as error
extern "C" const char* orc_combos_divide(uint32_t* combos, uint64_t nchunks, const uint32_t* pow_begin,
                                         const uint32_t* pows, uint64_t cycles) {
  ORC_TRY
  std::vector<std::vector<FpExt>> ch(nchunks);
  for (size_t i = 0; i < nchunks; i++)
    for (uint32_t k = pow_begin[i]; k < pow_begin[i + 1]; k++) ch[i].push_back(ext_of(pows + 4 * k));
  if (!combos_divide((FpExt*)combos, ch, cycles)) throw std::runtime_error("combos_divide: nonzero remainder");
  ORC_CATCH
}
// core/poly.rs:81-89 poly_divide: in-place synthetic division by (x - z); the remainder goes to rem_out (4 words)
extern "C" void orc_poly_divide(uint32_t* poly, uint64_t n, const uint32_t* z, uint32_t* rem_out) {
  FpExt rem = poly_divide((FpExt*)poly, n, ext_of(z));
  for (size_t k = 0; k < EXT_SIZE; k++) rem_out[k] = rem.e[k].v;
}
// map_pow(poly_mix, POLY_MIX_POWERS) as rv32im/src/prove/hal/cuda.rs:207-211 hands it to the FFI (458 x 4 words)
extern "C" uint32_t orc_rv32im_poly_mix_pows(const uint32_t* poly_mix, uint32_t* out) {
  FpExt pm = ext_of(poly_mix);
  for (size_t i = 0; i < RV32IM_NUM_POLY_MIX_POWERS; i++) {
    FpExt v = pm.pow(RV32IM_POLY_MIX_POWERS[i]);
    for (size_t k = 0; k < EXT_SIZE; k++) out[4 * i + k] = v.e[k].v;
  }
  return RV32IM_NUM_POLY_MIX_POWERS;
}
extern "C" void orc_poly_interpolate(uint32_t* out, const uint32_t* x, const uint32_t* fx, uint64_t size) {
  poly_interpolate((FpExt*)out, size, (const FpExt*)x, (const FpExt*)fx, size);
}

// ------------------------------------------------------------------ whole-segment prove / verify
struct SealOut {
  std::vector<uint32_t> seal;
  ProveTrace trace;
};

static const char* finish(SealOut* so, uint32_t* seal_out, uint64_t seal_cap, uint64_t* seal_len, uint32_t* roots_out,
                          uint64_t roots_cap, uint64_t* nroots, uint32_t* qpos_out) {
  *seal_len = so->seal.size();
  if (so->seal.size() > seal_cap) return strdup("seal buffer too small");
  memcpy(seal_out, so->seal.data(), so->seal.size() * 4);
  if (nroots) *nroots = so->trace.roots.size();
  if (roots_out) {
    if (so->trace.roots.size() > roots_cap) return strdup("roots buffer too small");
    memcpy(roots_out, so->trace.roots.data(), so->trace.roots.size() * 32);
  }
  if (qpos_out) memcpy(qpos_out, so->trace.query_pos.data(), so->trace.query_pos.size() * 4);
  return nullptr;
}

// Restates SegmentProverImpl::prove_core's "prove_inner" block (rv32im/src/prove/hal/mod.rs:171-222) for a GIVEN
// witness: code (1 x N), data (211 x N), accum (103 x N), global (90). The accum witness is an input (the synthetic
// workload of SURVEY §8d), so the 36 drawn mix values only feed eval_check.
// prove_core's prove_inner block (rv32im/src/prove/hal/mod.rs:171-222). mix_out (36 words, may be null) receives the
// accum mix drawn after the code and data commits (:213); with accum == nullptr the function stops there - that is
// what a caller sees between the two phases (the accum matrix is computed from that mix, :213-216).
typedef void (*accum_cb_t)(const uint32_t* mix, uint32_t* accum_out);
static const char* prove_rv32im_impl(int hash_kind, uint32_t po2, const uint32_t* code, const uint32_t* data,
                                     const uint32_t* accum, const uint32_t* global, uint32_t* mix_out, uint32_t* seal_out,
                                     uint64_t seal_cap, uint64_t* seal_len, uint32_t* roots_out, uint64_t roots_cap,
                                     uint64_t* nroots, uint32_t* qpos_out, accum_cb_t accum_cb);
static const char* prove_rv32im_impl(int hash_kind, uint32_t po2, const uint32_t* code, const uint32_t* data,
                                     const uint32_t* accum, const uint32_t* global, uint32_t* mix_out, uint32_t* seal_out,
                                     uint64_t seal_cap, uint64_t* seal_len, uint32_t* roots_out, uint64_t roots_cap,
                                     uint64_t* nroots, uint32_t* qpos_out) {
  return prove_rv32im_impl(hash_kind, po2, code, data, accum, global, mix_out, seal_out, seal_cap, seal_len, roots_out,
                           roots_cap, nroots, qpos_out, nullptr);
}
// two-phase form, the real protocol order (rv32im/src/prove/hal/mod.rs:209-217): the accum matrix comes from a
// callback that receives the mix drawn after the data commit
extern "C" const char* orc_prove_rv32im_cb(int hash_kind, uint32_t po2, const uint32_t* code, const uint32_t* data,
                                           const uint32_t* global, accum_cb_t accum_cb, uint32_t* seal_out,
                                           uint64_t seal_cap, uint64_t* seal_len, uint32_t* roots_out, uint64_t roots_cap,
                                           uint64_t* nroots, uint32_t* qpos_out) {
  return prove_rv32im_impl(hash_kind, po2, code, data, nullptr, global, nullptr, seal_out, seal_cap, seal_len, roots_out,
                           roots_cap, nroots, qpos_out, accum_cb);
}
static const char* prove_rv32im_impl(int hash_kind, uint32_t po2, const uint32_t* code, const uint32_t* data,
                                     const uint32_t* accum, const uint32_t* global, uint32_t* mix_out, uint32_t* seal_out,
                                     uint64_t seal_cap, uint64_t* seal_len, uint32_t* roots_out, uint64_t roots_cap,
                                     uint64_t* nroots, uint32_t* qpos_out, accum_cb_t accum_cb) {
  ORC_TRY
  HashSuite suite = suite_of(hash_kind);
  TapSet taps = rv32im_taps();
  size_t N = size_t(1) << po2;
  SealOut so;
  Prover prover(suite, taps);
  prover.trace = &so.trace;
  uint32_t version = RV32IM_SEAL_VERSION;
  prover.iop.write_u32(&version, 1);
  auto info16 = [&](const char* s) {
    Fp e[16];
    for (int i = 0; i < 16; i++) e[i] = Fp(uint32_t(uint8_t(s[i])));
    prover.iop.commit(suite.hash_elem_slice(e, 16));
  };
  info16("RISC0_STARK:v1__");
  info16(RV32IM_CIRCUIT_INFO);
  std::vector<Fp> header(RV32IM_OUTPUT_SIZE + 1);
  std::vector<Fp> glob(RV32IM_OUTPUT_SIZE);
  for (size_t i = 0; i < RV32IM_OUTPUT_SIZE; i++) header[i] = glob[i] = Fp::raw(global[i]).valid_or_zero();
  header[RV32IM_OUTPUT_SIZE] = Fp::raw(po2);
  prover.iop.commit(suite.hash_elem_slice(header.data(), header.size()));
  prover.iop.write_elems(header.data(), header.size());
  prover.set_po2(po2);
  prover.commit_group(1, (const Fp*)code, N * 1);
  prover.commit_group(2, (const Fp*)data, N * 211);
  std::vector<Fp> mix(RV32IM_MIX_SIZE);
  for (auto& m : mix) m = prover.iop.random_elem();
  if (mix_out)
    for (size_t i = 0; i < RV32IM_MIX_SIZE; i++) mix_out[i] = mix[i].v;
  std::vector<uint32_t> accum_from_cb;
  if (!accum && accum_cb) {
    accum_from_cb.resize(N * 103);
    std::vector<uint32_t> mix_words(RV32IM_MIX_SIZE);
    for (size_t i = 0; i < RV32IM_MIX_SIZE; i++) mix_words[i] = mix[i].v;
    accum_cb(mix_words.data(), accum_from_cb.data());
    accum = accum_from_cb.data();
  }
  if (!accum) return nullptr;
  prover.commit_group(0, (const Fp*)accum, N * 103);
  EvalCheckFn ec = [&](Fp* check, const std::vector<const Fp*>& g, const std::vector<const Fp*>& globals, FpExt pm,
                       size_t p2, size_t steps) {
    rv32im_eval_check(check, g[0], g[2], globals[0], globals[1], pm, p2, steps, 0, steps * INV_RATE);
  };
  so.seal = prover.finalize({mix.data(), glob.data()}, ec);
  const char* e = finish(&so, seal_out, seal_cap, seal_len, roots_out, roots_cap, nroots, qpos_out);
  if (e) return e;
  ORC_CATCH
}

extern "C" const char* orc_prove_rv32im(int hash_kind, uint32_t po2, const uint32_t* code, const uint32_t* data,
                                        const uint32_t* accum, const uint32_t* global, uint32_t* seal_out,
                                        uint64_t seal_cap, uint64_t* seal_len, uint32_t* roots_out, uint64_t roots_cap,
                                        uint64_t* nroots, uint32_t* qpos_out) {
  return prove_rv32im_impl(hash_kind, po2, code, data, accum, global, nullptr, seal_out, seal_cap, seal_len, roots_out,
                           roots_cap, nroots, qpos_out);
}

extern "C" const char* orc_prove_rv32im_mix(int hash_kind, uint32_t po2, const uint32_t* code, const uint32_t* data,
                                            const uint32_t* global, uint32_t* mix_out) {
  return prove_rv32im_impl(hash_kind, po2, code, data, nullptr, global, mix_out, nullptr, 0, nullptr, nullptr, 0, nullptr,
                           nullptr);
}

// HelloCircuit in the verify_v3 protocol (no header, groups committed in tap order accum, code, data).
extern "C" const char* orc_prove_hello(int hash_kind, uint32_t po2, const uint32_t* accum, const uint32_t* code,
                                       const uint32_t* data, uint32_t* seal_out, uint64_t seal_cap, uint64_t* seal_len) {
  ORC_TRY
  HashSuite suite = suite_of(hash_kind);
  TapSet taps = hello_taps();
  size_t N = size_t(1) << po2;
  SealOut so;
  Prover prover(suite, taps);
  prover.set_po2(po2);
  prover.commit_group(0, (const Fp*)accum, N);
  prover.commit_group(1, (const Fp*)code, N);
  prover.commit_group(2, (const Fp*)data, N);
  EvalCheckFn ec = [&](Fp* check, const std::vector<const Fp*>& g, const std::vector<const Fp*>&, FpExt pm, size_t p2,
                       size_t steps) { hello_eval_check(check, g, pm, p2, steps); };
  so.seal = prover.finalize({}, ec);
  const char* e = finish(&so, seal_out, seal_cap, seal_len, nullptr, 0, nullptr, nullptr);
  if (e) return e;
  ORC_CATCH
}

extern "C" const char* orc_verify_hello(int hash_kind, const uint32_t* seal, uint64_t n, uint32_t po2,
                                        uint32_t* roots_out, uint64_t* nroots) {
  ORC_TRY
  HashSuite suite = suite_of(hash_kind);
  TapSet taps = hello_taps();
  Verifier* v = nullptr;
  verify_v3_simple(taps, suite, seal, n, po2, hello_poly_ext, &v);
  if (nroots) *nroots = v->roots.size();
  if (roots_out) memcpy(roots_out, v->roots.data(), v->roots.size() * 32);
  ORC_CATCH
}

// rv32im seal: Merkle / FRI / DEEP consistency (constraint evaluation needs the missing poly_ext.rs blob).
// CircuitCoreDef::poly_ext supplied by the caller (verify/mod.rs:555-556 -> adapter.rs): evaluates the circuit's
// constraint polynomial at the extension-field tap evaluations. tests/poly_ext_ir.py implements it from the committed
// circuit IR (rv32im's own poly_ext.rs is a missing blob in the reference snapshot; the IR is the same polynomial,
// parsed from the reference's poly_fp and pinned against its compiled form point by point).
typedef void (*poly_ext_cb_t)(const uint32_t* poly_mix, const uint32_t* eval_u, uint64_t ntaps, const uint32_t* out,
                              uint64_t nout, const uint32_t* mix, uint64_t nmix, uint32_t* result);
static PolyExtFn wrap_poly_ext(poly_ext_cb_t cb) {
  if (!cb) return PolyExtFn();
  return [cb](FpExt poly_mix, const std::vector<FpExt>& eval_u, const std::vector<std::vector<Fp>>& args) {
    uint32_t res[4] = {0, 0, 0, 0};
    cb((const uint32_t*)&poly_mix, (const uint32_t*)eval_u.data(), eval_u.size(), (const uint32_t*)args[0].data(),
       args[0].size(), (const uint32_t*)args[1].data(), args[1].size(), res);
    return FpExt(Fp::raw(res[0]), Fp::raw(res[1]), Fp::raw(res[2]), Fp::raw(res[3]));
  };
}

extern "C" const char* orc_verify_rv32im_ext(int hash_kind, const uint32_t* seal, uint64_t n, uint32_t* roots_out,
                                             uint64_t* nroots, poly_ext_cb_t poly_ext, int* validity_checked) {
  ORC_TRY
  HashSuite suite = suite_of(hash_kind);
  TapSet taps = rv32im_taps();
  if (n == 0 || seal[0] != RV32IM_SEAL_VERSION) throw VerifyError("bad seal version word");
  Verifier* v = nullptr;
  verify_standard(taps, suite, seal, n, RV32IM_CIRCUIT_INFO, RV32IM_OUTPUT_SIZE, RV32IM_MIX_SIZE, 1,
                  wrap_poly_ext(poly_ext), &v);
  if (nroots) *nroots = v->roots.size();
  if (roots_out) memcpy(roots_out, v->roots.data(), v->roots.size() * 32);
  if (validity_checked) *validity_checked = v->validity_checked ? 1 : 0;
  ORC_CATCH
}
extern "C" const char* orc_verify_rv32im(int hash_kind, const uint32_t* seal, uint64_t n, uint32_t* roots_out,
                                         uint64_t* nroots) {
  return orc_verify_rv32im_ext(hash_kind, seal, n, roots_out, nroots, nullptr, nullptr);
}

// ------------------------------------------------------------------ recursion circuit
// reference-compiled poly_fp through oracle/ref_shim_recursion.cpp
static rv32im_poly_fp_t g_recursion_poly_fp = nullptr;
extern "C" const char* orc_load_ref_recursion(const char* path) {
  ORC_TRY
  void* h = dlopen(path, RTLD_NOW | RTLD_LOCAL);
  if (!h) throw std::runtime_error(std::string("dlopen failed: ") + dlerror());
  g_recursion_poly_fp = (rv32im_poly_fp_t)dlsym(h, "r0ref_recursion_poly_fp");
  if (!g_recursion_poly_fp) throw std::runtime_error("symbol r0ref_recursion_poly_fp not found");
  ORC_CATCH
}
extern "C" int orc_ref_recursion_loaded() { return g_recursion_poly_fp != nullptr; }

static TapSet recursion_taps() {
  return TapSet::from_tables(RECURSION_TAPS, RECURSION_NUM_TAPS, RECURSION_COMBO_TAPS, RECURSION_TOT_COMBO_BACKS,
                             RECURSION_COMBO_BEGIN, RECURSION_NUM_COMBOS, RECURSION_GROUP_BEGIN, 3, RECURSION_NUM_REGS);
}

// recursion CpuCircuitHal::eval_check (recursion/src/prove/hal/cpu.rs:106-150 ->
// risc0_circuit_recursion_cpu_eval_check, recursion-sys/kernels/cxx/ffi.cpp:219-246): args = ctrl, global, data, mix, accum
static void recursion_eval_check(Fp* check, const Fp* ctrl, const Fp* data, const Fp* accum, const Fp* mix,
                                 const Fp* global, FpExt poly_mix, size_t po2, size_t steps, size_t begin, size_t end) {
  if (!g_recursion_poly_fp) throw std::runtime_error("oracle/_ref recursion poly_fp not loaded");
  size_t domain = steps * INV_RATE;
  std::vector<FpExt> pows(RECURSION_NUM_POLY_MIX_POWERS);
  for (size_t i = 0; i < pows.size(); i++) pows[i] = poly_mix.pow(RECURSION_POLY_MIX_POWERS[i]);
  Fp rou = rou_fwd(unsigned(po2 + 2));
  std::string err;
#pragma omp parallel for schedule(dynamic, 64)
  for (size_t cycle = begin; cycle < end; cycle++) {
    Fp* args[5] = {const_cast<Fp*>(ctrl), const_cast<Fp*>(global), const_cast<Fp*>(data), const_cast<Fp*>(mix),
                   const_cast<Fp*>(accum)};
    FpExt tot;
    const char* e = g_recursion_poly_fp(cycle, domain, pows.data(), args, &tot);
    if (e) {
#pragma omp critical
      err = e;
      continue;
    }
    Fp x = rou.pow(cycle);
    Fp y = (Fp(3) * x).pow(size_t(1) << po2);
    FpExt ret = tot * (y - Fp(1)).inv();
    for (size_t k = 0; k < EXT_SIZE; k++) check[k * domain + cycle] = ret.e[k];
  }
  if (!err.empty()) throw std::runtime_error(err);
}
extern "C" const char* orc_recursion_eval_check(uint32_t* check, const uint32_t* ctrl, const uint32_t* data,
                                                const uint32_t* accum, const uint32_t* mix, const uint32_t* global,
                                                const uint32_t* poly_mix, uint32_t po2, uint64_t begin, uint64_t end) {
  ORC_TRY
  FpExt pm(Fp::raw(poly_mix[0]), Fp::raw(poly_mix[1]), Fp::raw(poly_mix[2]), Fp::raw(poly_mix[3]));
  recursion_eval_check((Fp*)check, (const Fp*)ctrl, (const Fp*)data, (const Fp*)accum, (const Fp*)mix, (const Fp*)global,
                       pm, po2, size_t(1) << po2, begin, end);
  ORC_CATCH
}

// RecursionProverImpl::prove's prove block (recursion/src/prove/mod.rs:179-224) for a GIVEN witness: ctrl (23 x N),
// data (128 x N), accum (12 x N), global (32). No version word (the rv32im crate adds one, recursion does not).
extern "C" const char* orc_prove_recursion(int hash_kind, uint32_t po2, const uint32_t* ctrl, const uint32_t* data,
                                           const uint32_t* accum, const uint32_t* global, uint32_t* seal_out,
                                           uint64_t seal_cap, uint64_t* seal_len, uint32_t* roots_out, uint64_t roots_cap,
                                           uint64_t* nroots, uint32_t* qpos_out) {
  ORC_TRY
  HashSuite suite = suite_of(hash_kind);
  TapSet taps = recursion_taps();
  size_t N = size_t(1) << po2;
  SealOut so;
  Prover prover(suite, taps);
  prover.trace = &so.trace;
  auto info16 = [&](const char* s) {
    Fp e[16];
    for (int i = 0; i < 16; i++) e[i] = Fp(uint32_t(uint8_t(s[i])));
    prover.iop.commit(suite.hash_elem_slice(e, 16));
  };
  info16("RISC0_STARK:v1__");
  info16(RECURSION_CIRCUIT_INFO);
  std::vector<Fp> header(RECURSION_OUTPUT_SIZE + 1);
  std::vector<Fp> glob(RECURSION_OUTPUT_SIZE);
  for (size_t i = 0; i < RECURSION_OUTPUT_SIZE; i++) header[i] = glob[i] = Fp::raw(global[i]).valid_or_zero();
  header[RECURSION_OUTPUT_SIZE] = Fp::raw(po2);
  prover.iop.commit(suite.hash_elem_slice(header.data(), header.size()));
  prover.iop.write_elems(header.data(), header.size());
  prover.set_po2(po2);
  prover.commit_group(1, (const Fp*)ctrl, N * RECURSION_GROUP_SIZES[1]);
  prover.commit_group(2, (const Fp*)data, N * RECURSION_GROUP_SIZES[2]);
  std::vector<Fp> mix(RECURSION_MIX_SIZE);
  for (auto& m : mix) m = prover.iop.random_elem();
  prover.commit_group(0, (const Fp*)accum, N * RECURSION_GROUP_SIZES[0]);
  EvalCheckFn ec = [&](Fp* check, const std::vector<const Fp*>& g, const std::vector<const Fp*>& globals, FpExt pm,
                       size_t p2, size_t steps) {
    recursion_eval_check(check, g[1], g[2], g[0], globals[0], globals[1], pm, p2, steps, 0, steps * INV_RATE);
  };
  so.seal = prover.finalize({mix.data(), glob.data()}, ec);
  const char* e = finish(&so, seal_out, seal_cap, seal_len, roots_out, roots_cap, nroots, qpos_out);
  if (e) return e;
  ORC_CATCH
}

// recursion seal: Merkle / FRI / DEEP consistency (poly_ext.rs is present in the reference but not restated here)
extern "C" const char* orc_verify_recursion_ext(int hash_kind, const uint32_t* seal, uint64_t n, uint32_t* roots_out,
                                                uint64_t* nroots, poly_ext_cb_t poly_ext, int* validity_checked) {
  ORC_TRY
  HashSuite suite = suite_of(hash_kind);
  TapSet taps = recursion_taps();
  Verifier* v = nullptr;
  verify_standard(taps, suite, seal, n, RECURSION_CIRCUIT_INFO, RECURSION_OUTPUT_SIZE, RECURSION_MIX_SIZE, 0,
                  wrap_poly_ext(poly_ext), &v);
  if (nroots) *nroots = v->roots.size();
  if (roots_out) memcpy(roots_out, v->roots.data(), v->roots.size() * 32);
  if (validity_checked) *validity_checked = v->validity_checked ? 1 : 0;
  ORC_CATCH
}
extern "C" const char* orc_verify_recursion(int hash_kind, const uint32_t* seal, uint64_t n, uint32_t* roots_out,
                                            uint64_t* nroots) {
  return orc_verify_recursion_ext(hash_kind, seal, n, roots_out, nroots, nullptr, nullptr);
}

extern "C" void orc_free_str(char* s) { free(s); }
extern "C" int orc_num_threads() {
  int n = 1;
#pragma omp parallel
  {
#pragma omp single
    n = omp_get_num_threads();
  }
  return n;
}
