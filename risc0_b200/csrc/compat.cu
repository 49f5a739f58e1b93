// The reference's own FFI symbol table (include/r0b200_compat.h) as thin shims over the r0b200 launchers: exact names,
// argument lists and return conventions of risc0/sys/kernels/zkp/cuda/ffi.cu:25-145, risc0/sys/src/cuda.rs:19-80 and
// risc0/circuit/{rv32im,recursion}-sys (eval_check), so the unmodified hal/cuda.rs links against libr0b200.so.
// Blocking like the reference's entry points; one default context on the current device.
#include "../../include/r0b200_compat.h"

#include <mutex>

#include "../../include/r0b200.h"
#include "ctx.h"
#include "launchers.h"
#include "tables/field_tables.h"

using namespace r0;

namespace {

std::mutex g_mu;
r0b200_ctx* g_ctx[64] = {nullptr};

// default context of the current device (created on first use; hal/cuda.rs:397-421 does the same work in
// CudaHal::new via sppark_init + cust::init)
r0b200_ctx* default_ctx() {
  int dev = 0;
  R0_CUDA(cudaGetDevice(&dev));
  R0_CHECK(dev >= 0 && dev < 64, "compat: device ordinal out of range");
  std::lock_guard<std::mutex> lock(g_mu);
  if (!g_ctx[dev]) {
    const char* e = r0b200_create(dev, &g_ctx[dev]);
    if (e) {
      std::string msg(e);
      r0b200_free_error(e);
      throw CudaError(msg);
    }
  }
  return g_ctx[dev];
}

void finish(r0b200_ctx* c) { R0_CUDA(cudaStreamSynchronize(c->stream)); }

template <typename T>
std::vector<T> fetch(r0b200_ctx* c, const void* dev, size_t n) {
  std::vector<T> host(n);
  if (n) R0_CUDA(cudaMemcpyAsync(host.data(), dev, n * sizeof(T), cudaMemcpyDeviceToHost, c->stream));
  R0_CUDA(cudaStreamSynchronize(c->stream));
  return host;
}

FpExt fetch_ext(r0b200_ctx* c, const uint32_t* dev) {
  std::vector<uint32_t> w = fetch<uint32_t>(c, dev, 4);
  return FpExt{{w[0], w[1], w[2], w[3]}};
}

int lg_of(size_t n, const char* what) {
  int k = 0;
  while ((size_t(1) << k) < n) k++;
  R0_CHECK((size_t(1) << k) == n, what);
  return k;
}

template <typename F>
const char* wrap(F f) {
  try {
    r0b200_ctx* c = default_ctx();
    f(c);
    finish(c);
  } catch (const std::exception& e) {
    return strdup(e.what());
  } catch (...) {
    return strdup("unknown C++ exception");
  }
  return nullptr;
}

template <typename F>
r0b200_sppark_error swrap(F f) {
  const char* e = wrap(f);
  return r0b200_sppark_error{e ? 1 : 0, const_cast<char*>(e)};
}

void eval_check_compat(r0b200_ctx* c, int circuit, uint32_t* check, const uint32_t* ctrl, const uint32_t* data,
                       const uint32_t* accum, const uint32_t* mix, const uint32_t* out, const uint32_t* rou, uint32_t po2,
                       uint32_t domain, const uint32_t* poly_mix_pows) {
  R0_CHECK(po2 + 2 < 28 && domain == (4u << po2), "eval_check: domain must be 4 * 2^po2");
  R0_CHECK(rou != nullptr && *rou == R0_ROU_FWD_MONT[po2 + 2], "eval_check: rou is not ROU_FWD[po2 + 2]");
  R0_CHECK(poly_mix_pows != nullptr, "eval_check: null poly_mix_pows");
  // POLY_MIX_POWERS starts 0, 1, ... for both circuits (rv32im/src/zirgen/info.rs:36, recursion/src/info.rs), so the
  // second table entry is poly_mix itself; the power table is re-derived inside from the circuit's own exponent list
  const FpExt poly_mix{{poly_mix_pows[4], poly_mix_pows[5], poly_mix_pows[6], poly_mix_pows[7]}};
  const size_t nmix = circuit == R0B200_CIRCUIT_RV32IM ? 36 : 20, nout = circuit == R0B200_CIRCUIT_RV32IM ? 90 : 32;
  std::vector<uint32_t> h_mix = fetch<uint32_t>(c, mix, nmix), h_out = fetch<uint32_t>(c, out, nout);
  if (circuit == R0B200_CIRCUIT_RV32IM)
    r0_eval_check_rv32im(c, check, accum, ctrl, data, h_out.data(), h_mix.data(), poly_mix, po2);
  else
    r0_eval_check_recursion(c, check, accum, ctrl, data, h_out.data(), h_mix.data(), poly_mix, po2);
}

}  // namespace

extern "C" {

const char* risc0_zkp_cuda_eltwise_add_fp(uint32_t* out, const uint32_t* x, const uint32_t* y, uint32_t count) {
  return wrap([&](r0b200_ctx* c) { r0_eltwise_add(c, out, x, y, count); });
}
const char* risc0_zkp_cuda_eltwise_mul_factor_fp(uint32_t* io, uint32_t factor, uint32_t count) {
  return wrap([&](r0b200_ctx* c) { r0_eltwise_mul_factor(c, io, factor, count); });
}
const char* risc0_zkp_cuda_eltwise_copy_fp(uint32_t* out, const uint32_t* in, uint32_t count) {
  return wrap([&](r0b200_ctx* c) { r0_eltwise_copy(c, out, in, count); });
}
const char* risc0_zkp_cuda_eltwise_copy_fp_region(uint32_t* into, const uint32_t* from, uint32_t from_rows,
                                                  uint32_t from_cols, uint32_t from_offset, uint32_t from_stride,
                                                  uint32_t into_offset, uint32_t into_stride) {
  return wrap([&](r0b200_ctx* c) {
    r0_copy_region_dev(c, into, from, from_rows, from_cols, from_offset, from_stride, into_offset, into_stride);
  });
}
const char* risc0_zkp_cuda_eltwise_sum_fpext(uint32_t* out, const uint32_t* in, uint32_t to_add, uint32_t count) {
  return wrap([&](r0b200_ctx* c) { r0_eltwise_sum_ext(c, out, in, count, to_add); });
}
const char* risc0_zkp_cuda_eltwise_zeroize_fp(uint32_t* elems, uint32_t count) {
  return wrap([&](r0b200_ctx* c) { r0_eltwise_zeroize(c, elems, count); });
}
const char* risc0_zkp_cuda_eltwise_zeroize_fpext(uint32_t* elems, uint32_t count) {
  return wrap([&](r0b200_ctx* c) { r0_eltwise_zeroize(c, elems, (size_t)count * 4); });
}
const char* risc0_zkp_cuda_fri_fold(uint32_t* out, const uint32_t* in, const uint32_t* mix, uint32_t count) {
  return wrap([&](r0b200_ctx* c) { r0_fri_fold(c, out, in, count, fetch_ext(c, mix)); });
}
const char* risc0_zkp_cuda_mix_poly_coeffs(uint32_t* out, const uint32_t* in, const uint32_t* combos,
                                           const uint32_t* mix_start, const uint32_t* mix, uint32_t input_size,
                                           uint32_t count) {
  return wrap([&](r0b200_ctx* c) {
    std::vector<uint32_t> ids = fetch<uint32_t>(c, combos, input_size);
    r0_mix_poly_coeffs(c, out, fetch_ext(c, mix_start), fetch_ext(c, mix), in, ids.data(), input_size, count);
  });
}
// hal/cuda.rs:594-613 passes the TOTAL element count; rows = count >> n_bits
const char* risc0_zkp_cuda_batch_bit_reverse(uint32_t* io, uint32_t n_bits, uint32_t count) {
  return wrap([&](r0b200_ctx* c) {
    R0_CHECK(n_bits <= (uint32_t)MAX_LG && (count & ((1u << n_bits) - 1)) == 0, "batch_bit_reverse: bad shape");
    r0_bit_reverse(c, io, count >> n_bits, (int)n_bits);
  });
}
// hal/cuda.rs:615-660: shared_size = threads_per_block * 16 bytes, count = evaluations * threads_per_block, deg = 2^po2
const char* risc0_zkp_cuda_batch_evaluate_any(uint32_t* out, const uint32_t* coeffs, const uint32_t* which,
                                              const uint32_t* xs, uint32_t shared_size, uint32_t count, uint32_t deg) {
  return wrap([&](r0b200_ctx* c) {
    R0_CHECK(shared_size >= 16 && count % (shared_size / 16) == 0, "batch_evaluate_any: bad launch geometry");
    lg_of(deg, "batch_evaluate_any: degree is not a power of two");
    r0_batch_evaluate_any(c, coeffs, deg, which, xs, out, count / (shared_size / 16));
  });
}
const char* risc0_zkp_cuda_gather_sample(uint32_t* dst, const uint32_t* src, uint32_t idx, uint32_t size,
                                         uint32_t stride) {
  return wrap([&](r0b200_ctx* c) { r0_gather_sample(c, dst, src, idx, size, stride); });
}
const char* risc0_zkp_cuda_scatter(uint32_t* into, const uint32_t* index, const uint32_t* offsets,
                                   const uint32_t* values, uint32_t count) {
  return wrap([&](r0b200_ctx* c) { r0_scatter_dev(c, into, index, count, offsets, values); });
}
const char* risc0_zkp_cuda_sha_rows(uint32_t* output, const uint32_t* matrix, uint32_t row_size, uint32_t col_size) {
  return wrap([&](r0b200_ctx* c) { r0_sha_hash_rows(c, output, matrix, row_size, col_size); });
}
const char* risc0_zkp_cuda_sha_fold(uint32_t* output, const uint32_t* input, uint32_t count) {
  return wrap([&](r0b200_ctx* c) { r0_sha_fold_pairs(c, output, input, count); });
}
const char* risc0_zkp_cuda_combos_prepare(uint32_t* combos, const uint32_t* coeff_u, uint32_t combo_count,
                                          uint32_t cycles, uint32_t regs_count, const uint32_t* reg_sizes,
                                          const uint32_t* reg_combo_ids, uint32_t check_size, const uint32_t* mix) {
  return wrap([&](r0b200_ctx* c) {
    std::vector<uint32_t> sizes = fetch<uint32_t>(c, reg_sizes, regs_count), ids = fetch<uint32_t>(c, reg_combo_ids, regs_count);
    size_t ulen = check_size;
    for (uint32_t s : sizes) ulen += s;
    std::vector<FpExt> u = fetch<FpExt>(c, coeff_u, ulen);
    r0_combos_prepare(c, combos, u.data(), ulen, combo_count, cycles, sizes.data(), ids.data(), regs_count, fetch_ext(c, mix),
                      check_size);
  });
}

r0b200_sppark_error sppark_init(void) {
  return swrap([&](r0b200_ctx*) {});   // creating the default context builds the twiddle tables (supra/ntt.cu:4-32)
}
// hal/cuda.rs:523-573 spells batch_expand_into_evaluate_ntt as sppark_batch_expand + sppark_batch_NTT. The expand's
// output is the zero-padded coefficient vector in bit-reversed order (coefficient j < n sits at brev_4n(j) =
// brev_n(j) << 2, i.e. out[i << lg_blowup] = in[i]); a full-size forward NTT (bit-reversed in, natural out) of it
// gives the evaluations. Both calls are stateless here; the fused one-pass form is r0b200_batch_expand_into_evaluate_ntt.
r0b200_sppark_error sppark_batch_expand(uint32_t* d_out, const uint32_t* d_in, uint32_t lg_domain_size, uint32_t lg_blowup,
                                        uint32_t poly_count) {
  return swrap([&](r0b200_ctx* c) {
    R0_CHECK(lg_domain_size + lg_blowup <= (uint32_t)MAX_LG, "batch_expand: size out of range");
    if (lg_domain_size == 0) return;   // supra/ntt.cu:35-36
    r0_expand_zero_interleave(c, d_out, d_in, (size_t)poly_count << lg_domain_size, (int)lg_blowup);
  });
}
r0b200_sppark_error sppark_batch_NTT(uint32_t* d_inout, uint32_t lg_domain_size, uint32_t poly_count) {
  return swrap([&](r0b200_ctx* c) {
    R0_CHECK(lg_domain_size <= (uint32_t)MAX_LG, "batch_NTT: size out of range");
    if (lg_domain_size == 0) return;
    const size_t words = (size_t)poly_count << lg_domain_size;
    uint32_t* tmp = nullptr;
    R0_CUDA(r0_malloc_async(c, &tmp, words ? words * 4 : 16, c->stream));
    r0_eltwise_copy(c, tmp, d_inout, words);
    r0_ntt_expand_evaluate(c, d_inout, tmp, poly_count, (int)lg_domain_size, 0, 0);
    R0_CUDA(cudaFreeAsync(tmp, c->stream));
  });
}
r0b200_sppark_error sppark_batch_iNTT(uint32_t* d_inout, uint32_t lg_domain_size, uint32_t poly_count) {
  return swrap([&](r0b200_ctx* c) {
    R0_CHECK(lg_domain_size <= (uint32_t)MAX_LG, "batch_iNTT: size out of range");
    r0_ntt_interpolate(c, d_inout, poly_count, (int)lg_domain_size, false, 0);
  });
}
r0b200_sppark_error sppark_batch_zk_shift(uint32_t* d_inout, uint32_t lg_domain_size, uint32_t poly_count) {
  return swrap([&](r0b200_ctx* c) {
    R0_CHECK(lg_domain_size <= (uint32_t)MAX_LG, "batch_zk_shift: size out of range");
    r0_zk_shift(c, d_inout, poly_count, (int)lg_domain_size);
  });
}
r0b200_sppark_error sppark_poseidon2_fold(uint32_t* d_out, const uint32_t* d_in, size_t num_hashes) {
  return swrap([&](r0b200_ctx* c) { r0_p2_fold_pairs(c, d_out, d_in, num_hashes); });
}
r0b200_sppark_error sppark_poseidon2_rows(uint32_t* d_out, const uint32_t* d_in, uint32_t count, uint32_t col_size) {
  return swrap([&](r0b200_ctx* c) { r0_p2_hash_rows(c, d_out, d_in, count, col_size); });
}
r0b200_sppark_error sppark_poseidon254_fold(void*, const void*, size_t) {
  return r0b200_sppark_error{1, strdup("poseidon254 (BN254, identity_p254 only) is outside the r0b200 backend's scope")};
}
r0b200_sppark_error sppark_poseidon254_rows(void*, const void*, size_t, uint32_t) {
  return r0b200_sppark_error{1, strdup("poseidon254 (BN254, identity_p254 only) is outside the r0b200 backend's scope")};
}
r0b200_sppark_error supra_poly_divide(uint32_t* polynomial, size_t poly_size, uint32_t* remainder, const uint32_t* pow) {
  return swrap([&](r0b200_ctx* c) {
    R0_CHECK(remainder != nullptr && pow != nullptr, "poly_divide: null host pointer");
    uint32_t* rem_dev = nullptr;
    R0_CUDA(r0_malloc_async(c, &rem_dev, 16, c->stream));
    r0_poly_divide(c, polynomial, poly_size, FpExt{{pow[0], pow[1], pow[2], pow[3]}}, rem_dev);
    R0_CUDA(cudaMemcpyAsync(remainder, rem_dev, 16, cudaMemcpyDeviceToHost, c->stream));
    R0_CUDA(cudaStreamSynchronize(c->stream));
    R0_CUDA(cudaFreeAsync(rem_dev, c->stream));
  });
}

// risc0_circuit_rv32im_cuda_witgen / _cuda_accum (rv32im-sys/src/lib.rs:105-119; kernels/cuda/ffi.cu:431-512): buffer
// pointers are DEVICE pointers (rv32im/src/prove/hal/cuda.rs:60-157), the trace arrays are HOST pointers that are
// copied to the device for the call
const char* risc0_circuit_rv32im_cuda_witgen(uint32_t mode, const r0b200_raw_exec_buffers* buffers,
                                             const r0b200_preflight_trace* preflight, uint32_t cycles) {
  return wrap([&](r0b200_ctx* c) {
    R0_CHECK(buffers != nullptr && preflight != nullptr && mode <= 2, "witgen: bad argument");
    R0_CHECK(buffers->data.rows == cycles && buffers->data.cols == 211 && buffers->global.cols == 90, "witgen: buffer shape");
    struct Guard {
      r0b200_trace* t;
      ~Guard() { r0_trace_free(t); }
    } g{r0_trace_upload(c, preflight, cycles, c->stream)};
    const uint32_t checked = (buffers->data.checked ? 1u : 0u) | (buffers->global.checked ? 4u : 0u);
    r0_witgen_rv32im(c, g.t, (uint32_t*)buffers->global.buf, (uint32_t*)buffers->data.buf, true, checked);
  });
}
const char* risc0_circuit_rv32im_cuda_accum(const r0b200_raw_accum_buffers* buffers, const r0b200_preflight_trace* preflight,
                                            uint32_t cycles) {
  return wrap([&](r0b200_ctx* c) {
    R0_CHECK(buffers != nullptr && preflight != nullptr, "accum: bad argument");
    R0_CHECK(buffers->data.rows == cycles && buffers->accum.rows == cycles && buffers->accum.cols == 103, "accum: buffer shape");
    struct Guard {
      r0b200_trace* t;
      ~Guard() { r0_trace_free(t); }
    } g{r0_trace_upload(c, preflight, cycles, c->stream)};
    const uint32_t checked = (buffers->data.checked ? 1u : 0u) | (buffers->accum.checked ? 2u : 0u) |
                             (buffers->global.checked ? 4u : 0u) | (buffers->mix.checked ? 8u : 0u);
    r0_accum_rv32im(c, g.t, (uint32_t*)buffers->data.buf, (uint32_t*)buffers->accum.buf, (uint32_t*)buffers->global.buf,
                    (uint32_t*)buffers->mix.buf, true, checked);
  });
}

const char* risc0_circuit_rv32im_cuda_eval_check(uint32_t* check, const uint32_t* ctrl, const uint32_t* data,
                                                 const uint32_t* accum, const uint32_t* mix, const uint32_t* out,
                                                 const uint32_t* rou, uint32_t po2, uint32_t domain,
                                                 const uint32_t* poly_mix_pows) {
  return wrap([&](r0b200_ctx* c) {
    eval_check_compat(c, R0B200_CIRCUIT_RV32IM, check, ctrl, data, accum, mix, out, rou, po2, domain, poly_mix_pows);
  });
}
const char* risc0_circuit_recursion_cuda_eval_check(uint32_t* check, const uint32_t* ctrl, const uint32_t* data,
                                                    const uint32_t* accum, const uint32_t* mix, const uint32_t* out,
                                                    const uint32_t* rou, uint32_t po2, uint32_t domain,
                                                    const uint32_t* poly_mix_pows) {
  return wrap([&](r0b200_ctx* c) {
    eval_check_compat(c, R0B200_CIRCUIT_RECURSION, check, ctrl, data, accum, mix, out, rou, po2, domain, poly_mix_pows);
  });
}

}  // extern "C"
