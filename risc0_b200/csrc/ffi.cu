// C ABI of libr0b200.so (declared in include/r0b200.h). Thin: argument checks + the launchers in the kernel files.
#include "../../include/r0b200.h"

#include <vector>

#include "ctx.h"
#include "launchers.h"

using namespace r0;

// the Hal passes mix / out as device buffers; the kernels take them in the parameter block (constant bank)
template <size_t MIX, size_t OUT, typename F>
static void eval_check_with_device_globals(r0b200_ctx* ctx, const uint32_t* mix, const uint32_t* out, F launch) {
  uint32_t host[MIX + OUT];
  R0_CUDA(cudaMemcpyAsync(host, mix, MIX * 4, cudaMemcpyDeviceToHost, ctx->stream));
  R0_CUDA(cudaMemcpyAsync(host + MIX, out, OUT * 4, cudaMemcpyDeviceToHost, ctx->stream));
  R0_CUDA(cudaStreamSynchronize(ctx->stream));
  launch(host + MIX, host);
}

extern "C" {

r0b200_err r0b200_create(int device, r0b200_ctx** out) {
  R0_API_BEGIN
  R0_CHECK(out != nullptr, "r0b200_create: null out pointer");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    throw CudaError(std::string("no CUDA device available (this backend has no CPU fallback): ") + cudaGetErrorString(e));
  R0_CHECK(device >= 0 && device < ndev, "r0b200_create: device ordinal out of range");
  R0_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  R0_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major < 10) throw CudaError("r0b200 is built for sm_100a only; found sm_" + std::to_string(prop.major * 10 + prop.minor));
  r0b200_ctx* c = new r0b200_ctx();
  c->device = device;
  c->sm_count = prop.multiProcessorCount;
  R0_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  R0_CUDA(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
  // keep freed blocks in the stream-ordered pool instead of returning them to the OS between proofs
  cudaMemPoolProps props = {};
  props.allocType = cudaMemAllocationTypePinned;
  props.handleTypes = cudaMemHandleTypeNone;
  props.location.type = cudaMemLocationTypeDevice;
  props.location.id = device;
  R0_CUDA(cudaMemPoolCreate(&c->pool, &props));
  uint64_t threshold = UINT64_MAX;
  R0_CUDA(cudaMemPoolSetAttribute(c->pool, cudaMemPoolAttrReleaseThreshold, &threshold));
  r0_ntt_init_tables(c);
  r0_poseidon2_init(c);
  *out = c;
  R0_API_END
}

void r0b200_destroy(r0b200_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  cudaStreamSynchronize(c->copy_stream);
  r0_ntt_free_tables(c);
  for (auto& r : c->phase_log) {
    cudaEventDestroy(r.a);
    cudaEventDestroy(r.b);
  }
  for (auto& e : c->event_pool) cudaEventDestroy(e);
  for (auto& e : c->aux_events) cudaEventDestroy(e);
  for (auto& st : c->aux_streams) cudaStreamDestroy(st);
  if (c->ev_start) cudaEventDestroy(c->ev_start);
  if (c->ev_stop) cudaEventDestroy(c->ev_stop);
  cudaStreamDestroy(c->copy_stream);
  cudaStreamDestroy(c->stream);
  if (c->pool) cudaMemPoolDestroy(c->pool);
  delete c;
}

void r0b200_free_error(const char* err) { free((void*)err); }

#define CTX_BEGIN                                       \
  R0_API_BEGIN                                          \
  R0_CHECK(ctx != nullptr, "null r0b200 context");      \
  R0_CUDA(cudaSetDevice(ctx->device));

r0b200_err r0b200_sync(r0b200_ctx* ctx) {
  CTX_BEGIN
  R0_CUDA(cudaStreamSynchronize(ctx->stream));
  R0_API_END
}
void* r0b200_stream(r0b200_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }
uint64_t r0b200_launch_count(r0b200_ctx* ctx) { return ctx ? ctx->launches : 0; }
uint64_t r0b200_bytes_peak(r0b200_ctx* ctx) { return ctx ? ctx->bytes_peak : 0; }

r0b200_err r0b200_timer_start(r0b200_ctx* ctx) {
  CTX_BEGIN
  if (!ctx->ev_start) {
    R0_CUDA(cudaEventCreate(&ctx->ev_start));
    R0_CUDA(cudaEventCreate(&ctx->ev_stop));
  }
  R0_CUDA(cudaEventRecord(ctx->ev_start, ctx->stream));
  R0_API_END
}
r0b200_err r0b200_timer_stop(r0b200_ctx* ctx, float* ms) {
  CTX_BEGIN
  R0_CHECK(ctx->ev_start != nullptr && ms != nullptr, "timer_stop without timer_start");
  R0_CUDA(cudaEventRecord(ctx->ev_stop, ctx->stream));
  R0_CUDA(cudaEventSynchronize(ctx->ev_stop));
  R0_CUDA(cudaEventElapsedTime(ms, ctx->ev_start, ctx->ev_stop));
  R0_API_END
}

r0b200_err r0b200_profile_begin(r0b200_ctx* ctx) {
  CTX_BEGIN
  for (auto& r : ctx->phase_log) {
    ctx->event_pool.push_back(r.a);
    ctx->event_pool.push_back(r.b);
  }
  ctx->phase_log.clear();
  ctx->profiling = true;
  R0_API_END
}
// Writes a JSON object {"<phase>": {"ms": total device ms, "n": launches of the phase, "bytes": algorithmic bytes}, ...}
r0b200_err r0b200_profile_end(r0b200_ctx* ctx, char* json_out, size_t cap) {
  CTX_BEGIN
  ctx->profiling = false;
  R0_CUDA(cudaStreamSynchronize(ctx->stream));
  struct Tot {
    double ms = 0, bytes = 0;
    size_t n = 0;
  };
  std::map<std::string, Tot> tot;
  for (auto& r : ctx->phase_log) {
    float ms = 0;
    R0_CUDA(cudaEventElapsedTime(&ms, r.a, r.b));
    Tot& t = tot[r.name];
    t.ms += ms;
    t.bytes += r.bytes;
    t.n++;
  }
  std::string js = "{";
  bool first = true;
  for (auto& kv : tot) {
    char buf[256];
    snprintf(buf, sizeof(buf), "%s\"%s\": {\"ms\": %.6f, \"n\": %zu, \"bytes\": %.0f}", first ? "" : ", ", kv.first.c_str(),
             kv.second.ms, kv.second.n, kv.second.bytes);
    js += buf;
    first = false;
  }
  js += "}";
  R0_CHECK(json_out != nullptr && js.size() + 1 <= cap, "profile_end: output buffer too small");
  memcpy(json_out, js.c_str(), js.size() + 1);
  R0_API_END
}

r0b200_err r0b200_alloc(r0b200_ctx* ctx, size_t bytes, void** dptr) {
  CTX_BEGIN
  R0_CHECK(dptr != nullptr, "r0b200_alloc: null out pointer");
  R0_CUDA(r0_malloc_async(ctx, dptr, bytes ? bytes : 16, ctx->stream));
  ctx->alloc_sizes[*dptr] = bytes;
  ctx->bytes_allocated += bytes;
  if (ctx->bytes_allocated > ctx->bytes_peak) ctx->bytes_peak = ctx->bytes_allocated;
  R0_API_END
}
r0b200_err r0b200_free(r0b200_ctx* ctx, void* dptr) {
  CTX_BEGIN
  if (dptr) {
    auto it = ctx->alloc_sizes.find(dptr);
    if (it != ctx->alloc_sizes.end()) {
      ctx->bytes_allocated -= it->second;
      ctx->alloc_sizes.erase(it);
    }
    R0_CUDA(cudaFreeAsync(dptr, ctx->stream));
  }
  R0_API_END
}
r0b200_err r0b200_copy_h2d(r0b200_ctx* ctx, void* dst, const void* src_host, size_t bytes) {
  CTX_BEGIN
  if (bytes) R0_CUDA(cudaMemcpyAsync(dst, src_host, bytes, cudaMemcpyHostToDevice, ctx->stream));
  R0_API_END
}
r0b200_err r0b200_copy_d2h(r0b200_ctx* ctx, void* dst_host, const void* src, size_t bytes) {
  CTX_BEGIN
  if (bytes) R0_CUDA(cudaMemcpyAsync(dst_host, src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  R0_CUDA(cudaStreamSynchronize(ctx->stream));
  R0_API_END
}
r0b200_err r0b200_fill_u32(r0b200_ctx* ctx, uint32_t* dst, uint32_t word, size_t count) {
  CTX_BEGIN
  r0_fill(ctx, dst, word, count);
  R0_API_END
}

r0b200_err r0b200_batch_interpolate_ntt(r0b200_ctx* ctx, uint32_t* io, size_t count, uint32_t lg_n) {
  CTX_BEGIN
  r0_ntt_interpolate(ctx, io, count, (int)lg_n, false, 0);
  R0_API_END
}
r0b200_err r0b200_batch_interpolate_ntt_zk(r0b200_ctx* ctx, uint32_t* io, size_t count, uint32_t lg_n) {
  CTX_BEGIN
  r0_ntt_interpolate(ctx, io, count, (int)lg_n, true, 0);
  R0_API_END
}
r0b200_err r0b200_zk_shift(r0b200_ctx* ctx, uint32_t* io, size_t count, uint32_t lg_n) {
  CTX_BEGIN
  R0_CHECK(lg_n <= MAX_LG, "zk_shift: size out of range");
  r0_zk_shift(ctx, io, count, (int)lg_n);
  R0_API_END
}
r0b200_err r0b200_batch_expand_into_evaluate_ntt(r0b200_ctx* ctx, uint32_t* out, const uint32_t* in, size_t count,
                                                 uint32_t lg_in, uint32_t expand_bits) {
  CTX_BEGIN
  r0_ntt_expand_evaluate(ctx, out, in, count, (int)(lg_in + expand_bits), (int)expand_bits, 0);
  R0_API_END
}
r0b200_err r0b200_batch_bit_reverse(r0b200_ctx* ctx, uint32_t* io, size_t count, uint32_t lg_n) {
  CTX_BEGIN
  r0_bit_reverse(ctx, io, count, (int)lg_n);
  R0_API_END
}

r0b200_err r0b200_hash_rows(r0b200_ctx* ctx, int hash, uint32_t* out, const uint32_t* matrix, size_t rows, size_t cols) {
  CTX_BEGIN
  if (hash == R0B200_HASH_POSEIDON2) {
    r0_p2_hash_rows(ctx, out, matrix, rows, cols);
  } else if (hash == R0B200_HASH_SHA256) {
    r0_sha_hash_rows(ctx, out, matrix, rows, cols);
  } else {
    throw std::invalid_argument("hash_rows: unknown hash suite");
  }
  R0_API_END
}
r0b200_err r0b200_hash_fold(r0b200_ctx* ctx, int hash, uint32_t* io, size_t input_size, size_t output_size) {
  CTX_BEGIN
  if (hash == R0B200_HASH_POSEIDON2) {
    r0_p2_hash_fold(ctx, io, input_size, output_size);
  } else if (hash == R0B200_HASH_SHA256) {
    r0_sha_hash_fold(ctx, io, input_size, output_size);
  } else {
    throw std::invalid_argument("hash_fold: unknown hash suite");
  }
  R0_API_END
}
r0b200_err r0b200_merkle_build(r0b200_ctx* ctx, int hash, uint32_t* nodes, const uint32_t* matrix, size_t rows,
                               size_t cols) {
  CTX_BEGIN
  R0_CHECK(rows > 0 && (rows & (rows - 1)) == 0, "merkle_build: rows must be a power of two");
  if (hash == R0B200_HASH_POSEIDON2) {
    r0_p2_hash_rows(ctx, nodes + rows * 8, matrix, rows, cols);
    r0_p2_merkle_fold_all(ctx, nodes, rows);
  } else if (hash == R0B200_HASH_SHA256) {
    r0_sha_hash_rows(ctx, nodes + rows * 8, matrix, rows, cols);
    for (size_t s = rows; s > 1; s >>= 1) r0_sha_hash_fold(ctx, nodes, s, s / 2);
  } else {
    throw std::invalid_argument("merkle_build: unknown hash suite");
  }
  R0_API_END
}

r0b200_err r0b200_eltwise_add_elem(r0b200_ctx* ctx, uint32_t* out, const uint32_t* a, const uint32_t* b, size_t n) {
  CTX_BEGIN
  r0_eltwise_add(ctx, out, a, b, n);
  R0_API_END
}
r0b200_err r0b200_eltwise_copy_elem(r0b200_ctx* ctx, uint32_t* out, const uint32_t* in, size_t n) {
  CTX_BEGIN
  r0_eltwise_copy(ctx, out, in, n);
  R0_API_END
}
r0b200_err r0b200_eltwise_zeroize_elem(r0b200_ctx* ctx, uint32_t* io, size_t n) {
  CTX_BEGIN
  r0_eltwise_zeroize(ctx, io, n);
  R0_API_END
}
r0b200_err r0b200_eltwise_sum_extelem(r0b200_ctx* ctx, uint32_t* out, const uint32_t* in, size_t count, size_t to_add) {
  CTX_BEGIN
  r0_eltwise_sum_ext(ctx, out, in, count, to_add);
  R0_API_END
}
r0b200_err r0b200_eltwise_copy_elem_slice(r0b200_ctx* ctx, uint32_t* into, const uint32_t* from_host, size_t from_rows,
                                          size_t from_cols, size_t from_offset, size_t from_stride, size_t into_offset,
                                          size_t into_stride) {
  CTX_BEGIN
  r0_copy_elem_slice(ctx, into, from_host, from_rows, from_cols, from_offset, from_stride, into_offset, into_stride);
  R0_API_END
}
static FpExt ext_from_host(const uint32_t* p) { return FpExt{{p[0], p[1], p[2], p[3]}}; }

r0b200_err r0b200_fri_fold(r0b200_ctx* ctx, uint32_t* out, const uint32_t* in, size_t count, const uint32_t* mix_host) {
  CTX_BEGIN
  r0_fri_fold(ctx, out, in, count, ext_from_host(mix_host));
  R0_API_END
}
r0b200_err r0b200_mix_poly_coeffs(r0b200_ctx* ctx, uint32_t* out, const uint32_t* mix_start_host,
                                  const uint32_t* mix_host, const uint32_t* in, const uint32_t* combos_host,
                                  size_t input_size, size_t count) {
  CTX_BEGIN
  r0_mix_poly_coeffs(ctx, out, ext_from_host(mix_start_host), ext_from_host(mix_host), in, combos_host, input_size, count);
  R0_API_END
}
r0b200_err r0b200_batch_evaluate_any(r0b200_ctx* ctx, const uint32_t* coeffs, size_t poly_count, uint32_t lg_n,
                                     const uint32_t* which, const uint32_t* xs, uint32_t* out, size_t eval_count) {
  CTX_BEGIN
  r0_batch_evaluate_any(ctx, coeffs, size_t(1) << lg_n, which, xs, out, eval_count, poly_count);
  R0_API_END
}
r0b200_err r0b200_gather_sample(r0b200_ctx* ctx, uint32_t* dst, const uint32_t* src, size_t idx, size_t size,
                                size_t stride) {
  CTX_BEGIN
  r0_gather_sample(ctx, dst, src, idx, size, stride);
  R0_API_END
}
r0b200_err r0b200_scatter(r0b200_ctx* ctx, uint32_t* into, const uint32_t* index_host, size_t index_len,
                          const uint32_t* offsets_host, const uint32_t* values_host) {
  CTX_BEGIN
  r0_scatter(ctx, into, index_host, index_len, offsets_host, values_host);
  R0_API_END
}
r0b200_err r0b200_prefix_products(r0b200_ctx* ctx, uint32_t* io_ext, size_t n) {
  CTX_BEGIN
  r0_prefix_products(ctx, io_ext, n);
  R0_API_END
}
r0b200_err r0b200_combos_prepare(r0b200_ctx* ctx, uint32_t* combos, const uint32_t* coeff_u_host, size_t coeff_u_len,
                                 uint32_t combo_count, size_t cycles, const uint32_t* reg_sizes_host,
                                 const uint32_t* reg_combo_ids_host, uint32_t nregs, const uint32_t* mix_host) {
  CTX_BEGIN
  r0_combos_prepare(ctx, combos, (const FpExt*)coeff_u_host, coeff_u_len, combo_count, cycles, reg_sizes_host,
                    reg_combo_ids_host, nregs, ext_from_host(mix_host), 16);
  R0_API_END
}
r0b200_err r0b200_combos_divide(r0b200_ctx* ctx, uint32_t* combos, size_t nchunks, const uint32_t* pow_begin_host,
                                const uint32_t* pows_host, size_t cycles) {
  CTX_BEGIN
  size_t ndiv = pow_begin_host[nchunks];
  if (ndiv == 0) return nullptr;
  uint32_t* rem_dev = nullptr;
  R0_CUDA(r0_malloc_async(ctx, &rem_dev, ndiv * 16, ctx->stream));
  // round r divides every combo that still has an r-th point: the combos are independent polynomials, so one round
  // is three launches whatever the number of combos (6 rounds instead of 15 divisions for rv32im)
  for (uint32_t r = 0;; r++) {
    std::vector<uint32_t*> polys, rems;
    std::vector<FpExt> zs;
    for (size_t i = 0; i < nchunks; i++) {
      const uint32_t k = pow_begin_host[i] + r;
      if (k < pow_begin_host[i + 1]) {
        polys.push_back(combos + i * cycles * 4);
        zs.push_back(ext_from_host(pows_host + 4 * k));
        rems.push_back(rem_dev + 4 * (k - pow_begin_host[0]));
      }
    }
    if (polys.empty()) break;
    r0_poly_divide_batch(ctx, polys.data(), zs.data(), rems.data(), polys.size(), cycles);
  }
  std::vector<uint32_t> rem(ndiv * 4);
  R0_CUDA(cudaMemcpyAsync(rem.data(), rem_dev, ndiv * 16, cudaMemcpyDeviceToHost, ctx->stream));
  R0_CUDA(cudaStreamSynchronize(ctx->stream));
  R0_CUDA(cudaFreeAsync(rem_dev, ctx->stream));
  for (size_t i = 0; i < rem.size(); i++)
    if (rem[i] != 0) throw std::runtime_error("combos_divide: non-zero remainder in division " + std::to_string(i / 4));
  R0_API_END
}

r0b200_err r0b200_eval_check_rv32im(r0b200_ctx* ctx, uint32_t* check, const uint32_t* ctrl, const uint32_t* data,
                                    const uint32_t* accum, const uint32_t* mix, const uint32_t* out,
                                    const uint32_t* poly_mix_host, uint32_t po2) {
  CTX_BEGIN
  eval_check_with_device_globals<36, 90>(ctx, mix, out, [&](const uint32_t* out_host, const uint32_t* mix_host) {
    r0_eval_check_rv32im(ctx, check, accum, ctrl, data, out_host, mix_host, ext_from_host(poly_mix_host), po2);
  });
  R0_API_END
}

r0b200_err r0b200_eval_check_recursion(r0b200_ctx* ctx, uint32_t* check, const uint32_t* ctrl, const uint32_t* data,
                                       const uint32_t* accum, const uint32_t* mix, const uint32_t* out,
                                       const uint32_t* poly_mix_host, uint32_t po2) {
  CTX_BEGIN
  eval_check_with_device_globals<20, 32>(ctx, mix, out, [&](const uint32_t* out_host, const uint32_t* mix_host) {
    r0_eval_check_recursion(ctx, check, accum, ctrl, data, out_host, mix_host, ext_from_host(poly_mix_host), po2);
  });
  R0_API_END
}

}  // extern "C"
