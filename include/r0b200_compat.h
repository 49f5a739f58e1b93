/* r0b200_compat — the reference's OWN FFI symbol table, exported by libr0b200.so.
 *
 * With these symbols the unmodified Rust of risc0/zkp/src/hal/cuda.rs (CudaHal) and
 * risc0/circuit/rv32im/src/prove/hal/cuda.rs (CudaCircuitHal::eval_check) links against libr0b200.so instead of
 * risc0-sys's own kernels + sppark: same names, same argument lists, same return conventions.
 *
 *   risc0_zkp_cuda_*            risc0/sys/kernels/zkp/cuda/ffi.cu:25-145 (Rust decls inline in zkp/src/hal/cuda.rs):
 *                               return `const char*`, NULL = ok, else a strdup'd message the caller frees
 *                               (risc0/sys/src/lib.rs:53-75); counts are u32; scalars such as `mix` arrive as
 *                               1-element DEVICE buffers (hal/cuda.rs:722-723,943).
 *   sppark_* / supra_poly_divide risc0/sys/src/cuda.rs:19-80 (bodies: risc0/sys/kernels/zkp/cuda/supra/{api,ntt}.cu):
 *                               return `sppark::Error {code: i32, message: *mut c_char}` BY VALUE; code 0 = ok; the
 *                               Rust Drop of sppark::Error frees `message` with libc free.
 *   risc0_circuit_rv32im_cuda_eval_check
 *                               risc0/circuit/rv32im-sys/src/lib.rs:121-132, kernels/cuda/ffi_supra.cu:50-79.
 *
 * All pointers are device pointers unless stated. Like the reference's entry points every call blocks until the GPU
 * has finished (cuda.h:77-100 creates a stream and synchronises per launch; the sppark wrappers cudaDeviceSynchronize).
 * They run on one process-wide default context on the CURRENT CUDA device (the reference hard-codes device 0,
 * hal/cuda.rs:406). New code should use the stream-ordered r0b200_* entry points of r0b200.h instead; this table exists
 * so that the reference's Rust can be pointed at this library without touching hal/cuda.rs.
 * sppark_poseidon254_{fold,rows} (BN254 Poseidon, identity_p254 only) are out of this backend's scope: they return a
 * non-zero code.
 */
#ifndef R0B200_COMPAT_H
#define R0B200_COMPAT_H

#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>

#include "r0b200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
  int32_t code;
  char* message;
} r0b200_sppark_error; /* layout of sppark::Error (crate sppark 0.1.12, #[repr(C)]) */

/* ---- risc0_zkp_cuda_* (ffi.cu:25-145) ---- */
const char* risc0_zkp_cuda_eltwise_add_fp(uint32_t* out, const uint32_t* x, const uint32_t* y, uint32_t count);
const char* risc0_zkp_cuda_eltwise_mul_factor_fp(uint32_t* io, uint32_t factor, uint32_t count);
const char* risc0_zkp_cuda_eltwise_copy_fp(uint32_t* out, const uint32_t* in, uint32_t count);
const char* risc0_zkp_cuda_eltwise_copy_fp_region(uint32_t* into, const uint32_t* from, uint32_t from_rows,
                                                  uint32_t from_cols, uint32_t from_offset, uint32_t from_stride,
                                                  uint32_t into_offset, uint32_t into_stride);
const char* risc0_zkp_cuda_eltwise_sum_fpext(uint32_t* out, const uint32_t* in, uint32_t to_add, uint32_t count);
const char* risc0_zkp_cuda_eltwise_zeroize_fp(uint32_t* elems, uint32_t count);
const char* risc0_zkp_cuda_eltwise_zeroize_fpext(uint32_t* elems, uint32_t count);
const char* risc0_zkp_cuda_fri_fold(uint32_t* out, const uint32_t* in, const uint32_t* mix, uint32_t count);
const char* risc0_zkp_cuda_mix_poly_coeffs(uint32_t* out, const uint32_t* in, const uint32_t* combos,
                                           const uint32_t* mix_start, const uint32_t* mix, uint32_t input_size,
                                           uint32_t count);
const char* risc0_zkp_cuda_batch_bit_reverse(uint32_t* io, uint32_t n_bits, uint32_t count);
const char* risc0_zkp_cuda_batch_evaluate_any(uint32_t* out, const uint32_t* coeffs, const uint32_t* which,
                                              const uint32_t* xs, uint32_t shared_size, uint32_t count, uint32_t deg);
const char* risc0_zkp_cuda_gather_sample(uint32_t* dst, const uint32_t* src, uint32_t idx, uint32_t size,
                                         uint32_t stride);
const char* risc0_zkp_cuda_scatter(uint32_t* into, const uint32_t* index, const uint32_t* offsets,
                                   const uint32_t* values, uint32_t count);
const char* risc0_zkp_cuda_sha_rows(uint32_t* output, const uint32_t* matrix, uint32_t row_size, uint32_t col_size);
const char* risc0_zkp_cuda_sha_fold(uint32_t* output, const uint32_t* input, uint32_t count);
const char* risc0_zkp_cuda_combos_prepare(uint32_t* combos, const uint32_t* coeff_u, uint32_t combo_count,
                                          uint32_t cycles, uint32_t regs_count, const uint32_t* reg_sizes,
                                          const uint32_t* reg_combo_ids, uint32_t check_size, const uint32_t* mix);

/* ---- sppark_* / supra_* (risc0/sys/src/cuda.rs:19-80) ---- */
r0b200_sppark_error sppark_init(void);
r0b200_sppark_error sppark_batch_expand(uint32_t* d_out, const uint32_t* d_in, uint32_t lg_domain_size,
                                        uint32_t lg_blowup, uint32_t poly_count);
r0b200_sppark_error sppark_batch_NTT(uint32_t* d_inout, uint32_t lg_domain_size, uint32_t poly_count);
r0b200_sppark_error sppark_batch_iNTT(uint32_t* d_inout, uint32_t lg_domain_size, uint32_t poly_count);
r0b200_sppark_error sppark_batch_zk_shift(uint32_t* d_inout, uint32_t lg_domain_size, uint32_t poly_count);
r0b200_sppark_error sppark_poseidon2_fold(uint32_t* d_out, const uint32_t* d_in, size_t num_hashes);
r0b200_sppark_error sppark_poseidon2_rows(uint32_t* d_out, const uint32_t* d_in, uint32_t count, uint32_t col_size);
r0b200_sppark_error sppark_poseidon254_fold(void* d_out, const void* d_in, size_t num_hashes);
r0b200_sppark_error sppark_poseidon254_rows(void* d_out, const void* d_in, size_t count, uint32_t col_size);
/* remainder (4 words) and pow (4 words) are HOST pointers (hal/cuda.rs:423-446) */
r0b200_sppark_error supra_poly_divide(uint32_t* polynomial, size_t poly_size, uint32_t* remainder, const uint32_t* pow);

/* ---- circuit (rv32im-sys/src/lib.rs:121-132): rou (1 word) and poly_mix_pows (458 x 4 words) are HOST pointers ---- */
const char* risc0_circuit_rv32im_cuda_eval_check(uint32_t* check, const uint32_t* ctrl, const uint32_t* data,
                                                 const uint32_t* accum, const uint32_t* mix, const uint32_t* out,
                                                 const uint32_t* rou, uint32_t po2, uint32_t domain,
                                                 const uint32_t* poly_mix_pows);

/* rv32im-sys/src/lib.rs:21-119: RawBuffer / RawExecBuffers / RawAccumBuffers (buf = DEVICE pointer, as CudaCircuitHal
 * passes them, rv32im/src/prove/hal/cuda.rs:60-157) and RawPreflightTrace (= r0b200_preflight_trace, HOST pointers) */
typedef struct {
  const uint32_t* buf;
  size_t rows;
  size_t cols;
  bool checked;
} r0b200_raw_buffer;
typedef struct {
  r0b200_raw_buffer global;
  r0b200_raw_buffer data;
} r0b200_raw_exec_buffers;
typedef struct {
  r0b200_raw_buffer data;
  r0b200_raw_buffer accum;
  r0b200_raw_buffer global;
  r0b200_raw_buffer mix;
} r0b200_raw_accum_buffers;
const char* risc0_circuit_rv32im_cuda_witgen(uint32_t mode, const r0b200_raw_exec_buffers* buffers,
                                             const r0b200_preflight_trace* preflight, uint32_t cycles);
const char* risc0_circuit_rv32im_cuda_accum(const r0b200_raw_accum_buffers* buffers, const r0b200_preflight_trace* preflight,
                                            uint32_t cycles);

/* recursion-sys/src/lib.rs:95-107, kernels/cuda/ffi_supra.cu:54-77: same shape, 158 poly-mix powers */
const char* risc0_circuit_recursion_cuda_eval_check(uint32_t* check, const uint32_t* ctrl, const uint32_t* data,
                                                    const uint32_t* accum, const uint32_t* mix, const uint32_t* out,
                                                    const uint32_t* rou, uint32_t po2, uint32_t domain,
                                                    const uint32_t* poly_mix_pows);

#ifdef __cplusplus
}
#endif
#endif /* R0B200_COMPAT_H */
