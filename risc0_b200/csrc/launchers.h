// Host-side launcher prototypes (one per kernel family). Internal to libr0b200.so.
#pragma once
#include "../../include/r0b200.h"
#include "ctx.h"

void r0_ntt_init_tables(r0::Ctx* c);
void r0_ntt_free_tables(r0::Ctx* c);
void r0_ntt_interpolate(r0::Ctx* c, uint32_t* io, size_t count, int k, bool zk, size_t cols_per_launch,
                        const uint32_t* src = nullptr);
void r0_ntt_expand_evaluate(r0::Ctx* c, uint32_t* out, const uint32_t* in, size_t count, int k, int eb,
                            size_t cols_per_launch);
void r0_bit_reverse(r0::Ctx* c, uint32_t* io, size_t count, int k);

void r0_poseidon2_init(r0::Ctx* c);
void r0_p2_hash_rows(r0::Ctx* c, uint32_t* out, const uint32_t* matrix, size_t rows, size_t cols);
void r0_p2_hash_fold(r0::Ctx* c, uint32_t* io, size_t in_size, size_t out_size);
void r0_p2_merkle_fold_all(r0::Ctx* c, uint32_t* nodes, size_t leaves);
void r0_p2_fold_pairs(r0::Ctx* c, uint32_t* out, const uint32_t* in, size_t n);
void r0_sha_fold_pairs(r0::Ctx* c, uint32_t* out, const uint32_t* in, size_t n);
void r0_sha_hash_rows(r0::Ctx* c, uint32_t* out, const uint32_t* matrix, size_t rows, size_t cols);
void r0_sha_hash_fold(r0::Ctx* c, uint32_t* io, size_t in_size, size_t out_size);

void r0_eltwise_add(r0::Ctx* c, uint32_t* out, const uint32_t* a, const uint32_t* b, size_t n);
void r0_eltwise_copy(r0::Ctx* c, uint32_t* out, const uint32_t* in, size_t n);
void r0_eltwise_zeroize(r0::Ctx* c, uint32_t* io, size_t n);
void r0_fill(r0::Ctx* c, uint32_t* io, uint32_t v, size_t n);
void r0_eltwise_sum_ext(r0::Ctx* c, uint32_t* out, const uint32_t* in, size_t count, size_t to_add);
void r0_zk_shift(r0::Ctx* c, uint32_t* io, size_t count, int bits);
void r0_fri_fold(r0::Ctx* c, uint32_t* out, const uint32_t* in, size_t count, const r0::FpExt& mix);
void r0_mix_poly_coeffs(r0::Ctx* c, uint32_t* out, const r0::FpExt& mix_start, const r0::FpExt& mix, const uint32_t* in,
                        const uint32_t* combos_host, size_t input_size, size_t count);
void r0_batch_evaluate_any(r0::Ctx* c, const uint32_t* coeffs, size_t n, const uint32_t* which_dev,
                           const uint32_t* xs_dev, uint32_t* out_dev, size_t eval_count, size_t distinct_polys = 0, const uint32_t* which_host = nullptr);
void r0_gather_sample(r0::Ctx* c, uint32_t* dst, const uint32_t* src, size_t idx, size_t size, size_t stride);
void r0_scatter(r0::Ctx* c, uint32_t* into, const uint32_t* index_host, size_t index_len, const uint32_t* offsets_host,
                const uint32_t* values_host);
void r0_scatter_dev(r0::Ctx* c, uint32_t* into, const uint32_t* index_dev, size_t count, const uint32_t* offsets_dev,
                    const uint32_t* values_dev);
void r0_copy_region_dev(r0::Ctx* c, uint32_t* into, const uint32_t* from_dev, size_t from_rows, size_t from_cols,
                        size_t from_offset, size_t from_stride, size_t into_offset, size_t into_stride);
void r0_expand_zero_interleave(r0::Ctx* c, uint32_t* out, const uint32_t* in, size_t n_in, int bits);
void r0_eltwise_mul_factor(r0::Ctx* c, uint32_t* io, uint32_t factor, size_t n);
void r0_copy_elem_slice(r0::Ctx* c, uint32_t* into, const uint32_t* from_host, size_t from_rows, size_t from_cols,
                        size_t from_offset, size_t from_stride, size_t into_offset, size_t into_stride);
void r0_prefix_products(r0::Ctx* c, uint32_t* io, size_t n);
void r0_combos_prepare(r0::Ctx* c, uint32_t* combos, const r0::FpExt* coeff_u_host, size_t coeff_u_len,
                       uint32_t combo_count, size_t cycles, const uint32_t* reg_sizes, const uint32_t* reg_combo_ids,
                       uint32_t nregs, const r0::FpExt& mix, uint32_t check_size);
void r0_poly_divide(r0::Ctx* c, uint32_t* poly, size_t n, const r0::FpExt& z, uint32_t* remainder_dev);
void r0_poly_divide_batch(r0::Ctx* c, uint32_t* const* polys, const r0::FpExt* zs, uint32_t* const* remainders_dev,
                          size_t njobs, size_t n);
// job j: dst[dst_off + g] = src[g * stride + idx] for g < size   (one MerkleTreeProver::prove column read)
struct GatherJob {
  const uint32_t* src;
  uint64_t idx, stride;
  uint32_t size, dst_off;
};
// dst[8*j .. 8*j+8) = nodes[8*idx ..]
struct DigestJob {
  const uint32_t* nodes;
  uint64_t idx;
};
void r0_gather_batched(r0::Ctx* c, uint32_t* dst, const GatherJob* jobs_host, size_t njobs);
void r0_gather_digests(r0::Ctx* c, uint32_t* dst, const DigestJob* jobs_host, size_t njobs);

// generated circuit kernels (csrc/gen/, tools/gen_eval_check.py)
void r0_eval_check_rv32im(r0::Ctx* c, uint32_t* check, const uint32_t* accum, const uint32_t* code, const uint32_t* data,
                          const uint32_t* global_host, const uint32_t* mix_host, const r0::FpExt& poly_mix, uint32_t po2);
void r0_eval_check_recursion(r0::Ctx* c, uint32_t* check, const uint32_t* accum, const uint32_t* code,
                             const uint32_t* data, const uint32_t* global_host, const uint32_t* mix_host,
                             const r0::FpExt& poly_mix, uint32_t po2);

// rv32im witness generation / accumulation on the device (witgen.cu)
r0b200_trace* r0_trace_upload(r0::Ctx* c, const r0b200_preflight_trace* trace_host, uint32_t cycles, cudaStream_t stream);
void r0_trace_free(r0b200_trace* t);
// checked: bit k = Buffer::checked of buffer k (0 data, 1 accum, 2 global, 3 mix); the reference's CPU path checks all
void r0_witgen_rv32im(r0::Ctx* c, r0b200_trace* t, uint32_t* global, uint32_t* data, bool sync_check, uint32_t checked = 0xf);
void r0_accum_rv32im(r0::Ctx* c, r0b200_trace* t, uint32_t* data, uint32_t* accum, uint32_t* global, uint32_t* mix,
                     bool sync_check, uint32_t checked = 0xf);
