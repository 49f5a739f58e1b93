// rv32im witness generation and accumulation on the device: CircuitWitnessGenerator::generate_witness and
// CircuitAccumulator::step_accum (risc0/circuit/rv32im/src/prove/hal/mod.rs:82-102), replacing
// risc0_circuit_rv32im_cuda_witgen / _cuda_accum (rv32im-sys/kernels/cuda/ffi.cu:362-512; CPU spec kernels/cxx/ffi.cpp:265-365).
//
// The per-cycle step functions are generated from the circuit IR (tools/gen_witgen.py -> gen/witgen_rv32im.inc) against
// witgen_rt.cuh. What this file adds is the B200 scheduling around them:
//   * cycles are NOT processed in trace order. A cycle's work is one arm of the circuit's major/minor mux (13 majors x 8
//     minors: ALU, mem, ecall, Poseidon2 paging, table cycles ...), so 32 consecutive cycles of a real trace diverge
//     into many arms. The step functions are order-independent inside a phase (the reference runs them in parallel and
//     tests forward == reverse, prove/witgen/tests.rs:64-135), so each phase is first bucket-sorted on the device by
//     (major, minor) and the step kernel walks the sorted order: warps are convergent except at bucket boundaries, and
//     column stores of neighbouring lanes stay near each other for the dominant buckets.
//   * the 2^8 + 2^16 lookup counters live in global memory with warp-aggregated increments (witgen_rt.cuh).
//   * accum: step kernel in the same sorted order, then the 4-column BabyBear prefix sum as a three-kernel blocked
//     scan (block totals -> scan of totals -> apply), then the back-propagation of the running totals
//     (ffi.cpp:330-358), all stream-ordered; errors come back through one 4-word block instead of device asserts.
#include <algorithm>
#include <memory>
#include <string>
#include <vector>

#include "../../include/r0b200.h"
#include "ctx.h"
#include "launchers.h"
#include "witgen_rt.cuh"

namespace r0wg {
#define R0_WG_TABLES_ONLY
#include "gen/witgen_rv32im.inc"
static const uint16_t kLayoutHost[] = R0_WG_LAYOUT_DATA;
static_assert(sizeof(kLayoutHost) / sizeof(uint16_t) == R0_WG_LAYOUT_WORDS, "layout table size");

constexpr int NKEYS = 128;   // major * 8 + minor, major < 16

__global__ void k_key_histogram(const PreflightCycle* cycles, uint32_t begin, uint32_t count, uint32_t* hist) {
  __shared__ uint32_t sh[NKEYS];
  for (int i = threadIdx.x; i < NKEYS; i += blockDim.x) sh[i] = 0;
  __syncthreads();
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x) {
    const PreflightCycle& c = cycles[begin + i];
    atomicAdd(&sh[(c.major * 8u + c.minor) & (NKEYS - 1)], 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < NKEYS; i += blockDim.x)
    if (sh[i]) atomicAdd(&hist[i], sh[i]);
}
// exclusive scan of the 128 bucket sizes -> bucket cursors
__global__ void k_key_offsets(uint32_t* hist) {
  if (threadIdx.x == 0) {
    uint32_t run = 0;
    for (int i = 0; i < NKEYS; i++) {
      const uint32_t n = hist[i];
      hist[i] = run;
      run += n;
    }
  }
}
// order[begin + slot] = cycle; a block reserves one range per bucket, lanes take slots inside it
__global__ void k_key_scatter(const PreflightCycle* cycles, uint32_t begin, uint32_t count, uint32_t* cursor,
                              uint32_t* order) {
  __shared__ uint32_t sh_count[NKEYS], sh_base[NKEYS];
  const uint32_t per_block = (count + gridDim.x - 1) / gridDim.x;
  const uint32_t b = blockIdx.x * per_block, e = min(count, b + per_block);
  for (int i = threadIdx.x; i < NKEYS; i += blockDim.x) sh_count[i] = 0;
  __syncthreads();
  for (uint32_t i = b + threadIdx.x; i < e; i += blockDim.x) {
    const PreflightCycle& c = cycles[begin + i];
    atomicAdd(&sh_count[(c.major * 8u + c.minor) & (NKEYS - 1)], 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < NKEYS; i += blockDim.x) {
    sh_base[i] = sh_count[i] ? atomicAdd(&cursor[i], sh_count[i]) : 0u;
    sh_count[i] = 0;
  }
  __syncthreads();
  for (uint32_t i = b + threadIdx.x; i < e; i += blockDim.x) {
    const PreflightCycle& c = cycles[begin + i];
    const uint32_t key = (c.major * 8u + c.minor) & (NKEYS - 1);
    order[begin + sh_base[key] + atomicAdd(&sh_count[key], 1u)] = begin + i;
  }
}

// the two kernels that instantiate the generated step functions live in their own translation units
// (witgen_step_exec.cu / witgen_step_accum.cu): ptxas needs minutes for each, so they build in parallel
void launch_step_exec(cudaStream_t stream, const WShared* s, const uint32_t* order, uint32_t begin, uint32_t count);
void launch_step_accum(cudaStream_t stream, const WShared* s, const uint32_t* order, uint32_t count);

// ---- inclusive BabyBear prefix sum over rows [0, n) of `ncols` columns (column stride = rows words) ----------------
constexpr int SCAN_T = 256, SCAN_E = 8, SCAN_SEG = SCAN_T * SCAN_E;   // elements per block

__device__ __forceinline__ uint32_t block_scan_excl(uint32_t v, uint32_t* sh, uint32_t* total) {
  // exclusive scan of one value per thread over the block (SCAN_T threads), field addition
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t x = v;
  for (int d = 1; d < 32; d <<= 1) {
    const uint32_t y = __shfl_up_sync(0xffffffffu, x, d);
    if (lane >= d) x = r0::fp_add(x, y);
  }
  if (lane == 31) sh[warp] = x;
  __syncthreads();
  if (warp == 0) {
    uint32_t w = lane < SCAN_T / 32 ? sh[lane] : 0u;
    for (int d = 1; d < SCAN_T / 32; d <<= 1) {
      const uint32_t y = __shfl_up_sync(0xffffffffu, w, d);
      if (lane >= d) w = r0::fp_add(w, y);
    }
    if (lane < SCAN_T / 32) sh[lane] = w;
  }
  __syncthreads();
  const uint32_t warp_off = warp ? sh[warp - 1] : 0u;
  *total = sh[SCAN_T / 32 - 1];
  const uint32_t incl = r0::fp_add(x, warp_off);
  return r0::fp_sub(incl, v);
}

// pass 1: per-block totals; pass 3 (apply = true): rescan with the scanned block offset added
template <bool APPLY>
__global__ void __launch_bounds__(SCAN_T) k_scan_cols(uint32_t* cols, size_t col_stride, uint32_t n, uint32_t* block_tot,
                                                     uint32_t nblocks) {
  __shared__ uint32_t sh[SCAN_T / 32];
  uint32_t* col = cols + (size_t)blockIdx.y * col_stride;
  const uint32_t base = blockIdx.x * SCAN_SEG + threadIdx.x * SCAN_E;
  uint32_t v[SCAN_E];
  uint32_t sum = 0;
#pragma unroll
  for (int k = 0; k < SCAN_E; k++) {
    v[k] = base + k < n ? col[base + k] : 0u;
    sum = r0::fp_add(sum, v[k]);
  }
  uint32_t total;
  uint32_t run = block_scan_excl(sum, sh, &total);
  if (!APPLY) {
    if (threadIdx.x == 0) block_tot[blockIdx.y * nblocks + blockIdx.x] = total;
    return;
  }
  if (blockIdx.x > 0) run = r0::fp_add(run, block_tot[blockIdx.y * nblocks + blockIdx.x - 1]);
#pragma unroll
  for (int k = 0; k < SCAN_E; k++) {
    run = r0::fp_add(run, v[k]);
    if (base + k < n) col[base + k] = run;
  }
}
// pass 2: inclusive scan of each column's block totals (nblocks <= 2^22 / 2048 = 2048: one block, serial chunks)
__global__ void __launch_bounds__(SCAN_T) k_scan_totals(uint32_t* block_tot, uint32_t nblocks) {
  __shared__ uint32_t sh[SCAN_T / 32];
  uint32_t* t = block_tot + blockIdx.x * nblocks;
  uint32_t carry = 0;
  for (uint32_t b = 0; b < nblocks; b += SCAN_T) {
    const uint32_t i = b + threadIdx.x;
    const uint32_t v = i < nblocks ? t[i] : 0u;
    uint32_t total;
    const uint32_t ex = block_scan_excl(v, sh, &total);
    if (i < nblocks) t[i] = r0::fp_add(carry, r0::fp_add(ex, v));
    carry = r0::fp_add(carry, total);
    __syncthreads();
  }
}
// finalizeAccum (ffi.cpp:344-358 / ffi.cu:412-429): every machine column group but the last gets the previous row's
// running totals added
__global__ void k_accum_finalize(uint32_t* accum, uint32_t rows, uint32_t cols, uint32_t split, uint32_t n) {
  const uint32_t row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= n) return;
  const uint32_t back1 = (row + n - 1) % n;
  uint32_t prev[4];
#pragma unroll
  for (int k = 0; k < 4; k++) prev[k] = accum[(size_t)(cols - 4 + k) * rows + back1];
  const uint32_t groups = (cols - split) / 4;
  for (uint32_t j = 0; j + 1 < groups; j++) {
#pragma unroll
    for (int k = 0; k < 4; k++) {
      uint32_t* p = accum + (size_t)(split + j * 4 + k) * rows + row;
      *p = r0::fp_add(*p, prev[k]);
    }
  }
}

}  // namespace r0wg

using namespace r0;
using namespace r0wg;

// A preflight trace resident on the device (cycles, txns, bigint bytes) + the per-phase sorted cycle order.
struct r0b200_trace {
  Ctx* c = nullptr;
  uint32_t cycles = 0, txns_len = 0, bigint_len = 0, split = 0;
  PreflightCycle* d_cycles = nullptr;
  MemoryTxn* d_txns = nullptr;
  uint8_t* d_bigint = nullptr;
  uint32_t* d_order = nullptr;   // [0, split) and [split, cycles) each sorted by (major, minor)
  uint32_t* d_order_all = nullptr;   // [0, cycles) sorted as one range (accum)
  uint32_t* d_tables = nullptr;  // 256 + 65536 lookup counters
  uint32_t* d_err = nullptr;     // 4 words
  WShared* d_shared = nullptr;
  uint16_t* d_layout = nullptr;
};

namespace {

const char* kErrNames[] = {"", "Inconsistent set", "Read of unset value", "eqz failure", "txn cycle mismatch",
                           "memory peek not in preflight", "Invalid lookup table", "u8/16 table error",
                           "Reached unreachable mux arm"};

void sort_range(Ctx* c, cudaStream_t stream, r0b200_trace* t, uint32_t begin, uint32_t count, uint32_t* order,
                uint32_t* scratch_hist) {
  if (count == 0) return;
  R0_CUDA(cudaMemsetAsync(scratch_hist, 0, NKEYS * 4, stream));
  const unsigned blocks = (unsigned)std::min<size_t>((count + 1023) / 1024, (size_t)c->sm_count * 4);
  k_key_histogram<<<blocks, 256, 0, stream>>>(t->d_cycles, begin, count, scratch_hist);
  k_key_offsets<<<1, 32, 0, stream>>>(scratch_hist);
  k_key_scatter<<<blocks, 256, 0, stream>>>(t->d_cycles, begin, count, scratch_hist, order);
  count_launch(c, 3);
  R0_CUDA(cudaGetLastError());
}

void check_errors(Ctx* c, r0b200_trace* t, const char* what) {
  uint32_t err[4];
  R0_CUDA(cudaMemcpyAsync(err, t->d_err, sizeof(err), cudaMemcpyDeviceToHost, c->stream));
  R0_CUDA(cudaStreamSynchronize(c->stream));
  if (err[0]) {
    const char* name = err[1] < sizeof(kErrNames) / sizeof(kErrNames[0]) ? kErrNames[err[1]] : "?";
    throw std::runtime_error(std::string(what) + ": " + std::to_string(err[0]) + " failure(s), first: " + name +
                             " at cycle " + std::to_string(err[2]) + " (detail " + std::to_string(err[3]) + ")");
  }
}

// checked: bit k = Buffer::checked of buffer k (BUF_DATA .. BUF_MIX)
void upload_shared(Ctx* c, r0b200_trace* t, uint32_t* data, uint32_t* accum, uint32_t* global, uint32_t* mix, uint32_t checked) {
  WShared s;
  memset(&s, 0, sizeof(s));
  const uint32_t rows = t->cycles;
  s.bufs[BUF_DATA] = WBuf{data, rows, R0_WG_KREGCOUNTDATA, (checked >> BUF_DATA) & 1u, 0};
  s.bufs[BUF_ACCUM] = WBuf{accum, rows, R0_WG_KREGCOUNTACCUM, (checked >> BUF_ACCUM) & 1u, R0_WG_USER_ACCUM_SPLIT};
  s.bufs[BUF_GLOBAL] = WBuf{global, 1, R0_WG_KREGCOUNTGLOBAL, (checked >> BUF_GLOBAL) & 1u, 0};
  s.bufs[BUF_MIX] = WBuf{mix, 1, R0_WG_KREGCOUNTMIX, (checked >> BUF_MIX) & 1u, 0};
  s.cycles = t->d_cycles;
  s.txns = t->d_txns;
  s.bigint_bytes = t->d_bigint;
  s.layout = t->d_layout;
  s.table_u8 = t->d_tables;
  s.table_u16 = t->d_tables + 256;
  s.err = t->d_err;
  s.txns_len = t->txns_len;
  s.bigint_len = t->bigint_len;
  // small launch-wide block: staged through a pageable copy so the caller may return before the GPU reads it
  R0_CUDA(cudaMemcpyAsync(t->d_shared, &s, sizeof(s), cudaMemcpyHostToDevice, c->stream));
  R0_CUDA(cudaMemsetAsync(t->d_err, 0, 16, c->stream));
}

}  // namespace

// `stream`: the stream the allocations, copies and the sort run on (the compute stream, or the copy stream when a
// segment is uploaded while the previous one is being proved); the trace is released on the compute stream
r0b200_trace* r0_trace_upload(Ctx* c, const r0b200_preflight_trace* h, uint32_t cycles, cudaStream_t stream) {
  R0_CHECK(h != nullptr && h->cycles != nullptr && (h->txns != nullptr || h->txns_len == 0), "trace_upload: null trace");
  R0_CHECK(cycles > 0 && (cycles & (cycles - 1)) == 0, "trace_upload: cycle count must be a power of two");
  R0_CHECK(h->table_split_cycle <= cycles, "trace_upload: table split beyond the last cycle");
  std::unique_ptr<r0b200_trace> t(new r0b200_trace());
  t->c = c;
  t->cycles = cycles;
  t->txns_len = h->txns_len;
  t->bigint_len = h->bigint_bytes_len;
  t->split = h->table_split_cycle;
  cudaStream_t s = stream ? stream : c->stream;
  R0_CUDA(r0_malloc_async(c, &t->d_cycles, (size_t)cycles * sizeof(PreflightCycle), s));
  R0_CUDA(r0_malloc_async(c, &t->d_txns, std::max<size_t>(1, h->txns_len) * sizeof(MemoryTxn), s));
  R0_CUDA(r0_malloc_async(c, &t->d_bigint, std::max<size_t>(16, h->bigint_bytes_len), s));
  R0_CUDA(r0_malloc_async(c, &t->d_order, (size_t)cycles * 4, s));
  R0_CUDA(r0_malloc_async(c, &t->d_order_all, (size_t)cycles * 4, s));
  R0_CUDA(r0_malloc_async(c, &t->d_tables, (256 + 65536 + NKEYS) * 4, s));
  R0_CUDA(r0_malloc_async(c, &t->d_err, 16, s));
  R0_CUDA(r0_malloc_async(c, &t->d_shared, sizeof(WShared), s));
  R0_CUDA(r0_malloc_async(c, &t->d_layout, sizeof(kLayoutHost), s));
  R0_CUDA(cudaMemcpyAsync(t->d_cycles, h->cycles, (size_t)cycles * sizeof(PreflightCycle), cudaMemcpyHostToDevice, s));
  if (h->txns_len)
    R0_CUDA(cudaMemcpyAsync(t->d_txns, h->txns, (size_t)h->txns_len * sizeof(MemoryTxn), cudaMemcpyHostToDevice, s));
  if (h->bigint_bytes_len)
    R0_CUDA(cudaMemcpyAsync(t->d_bigint, h->bigint_bytes, h->bigint_bytes_len, cudaMemcpyHostToDevice, s));
  R0_CUDA(cudaMemcpyAsync(t->d_layout, kLayoutHost, sizeof(kLayoutHost), cudaMemcpyHostToDevice, s));
  uint32_t* hist = t->d_tables + 256 + 65536;
  sort_range(c, s, t.get(), 0, t->split, t->d_order, hist);
  sort_range(c, s, t.get(), t->split, cycles - t->split, t->d_order, hist);
  sort_range(c, s, t.get(), 0, cycles, t->d_order_all, hist);
  return t.release();
}

void r0_trace_free(r0b200_trace* t) {
  if (!t) return;
  cudaStream_t s = t->c->stream;
  void* ptrs[] = {t->d_cycles, t->d_txns, t->d_bigint, t->d_order, t->d_order_all, t->d_tables, t->d_err, t->d_shared, t->d_layout};
  for (void* p : ptrs)
    if (p) cudaFreeAsync(p, s);
  delete t;
}

// generate_witness: phase 1 = cycles [0, split) (they also count the u8 / u16 lookups), phase 2 = [split, cycles)
// (the table cycles read the counters), ffi.cpp:284-296
void r0_witgen_rv32im(Ctx* c, r0b200_trace* t, uint32_t* global, uint32_t* data, bool sync_check, uint32_t checked) {
  PhaseScope ph(c, "witgen", 4.0 * R0_WG_KREGCOUNTDATA * (double)t->cycles + 36.0 * t->cycles + 20.0 * t->txns_len);
  upload_shared(c, t, data, nullptr, global, nullptr, checked);
  R0_CUDA(cudaMemsetAsync(t->d_tables, 0, (256 + 65536) * 4, c->stream));
  const uint32_t n1 = t->split, n2 = t->cycles - t->split;
  if (n1) launch_step_exec(c->stream, t->d_shared, t->d_order, 0, n1);
  if (n2) launch_step_exec(c->stream, t->d_shared, t->d_order, n1, n2);
  count_launch(c, (n1 ? 1 : 0) + (n2 ? 1 : 0));
  R0_CUDA(cudaGetLastError());
  if (sync_check) check_errors(c, t, "witgen");
}

// step_accum + prefix sums + back-propagation (ffi.cpp:316-365)
void r0_accum_rv32im(Ctx* c, r0b200_trace* t, uint32_t* data, uint32_t* accum, uint32_t* global, uint32_t* mix,
                     bool sync_check, uint32_t checked) {
  PhaseScope ph(c, "accum", 4.0 * (R0_WG_KREGCOUNTDATA + 2.0 * R0_WG_KREGCOUNTACCUM) * (double)t->cycles);
  upload_shared(c, t, data, accum, global, mix, checked);
  const uint32_t n = t->cycles, rows = t->cycles, cols = R0_WG_KREGCOUNTACCUM;
  launch_step_accum(c->stream, t->d_shared, t->d_order_all, n);
  const uint32_t nblocks = (n + SCAN_SEG - 1) / SCAN_SEG;
  uint32_t* tot = nullptr;
  R0_CUDA(r0_malloc_async(c, &tot, (size_t)4 * nblocks * 4, c->stream));
  uint32_t* last4 = accum + (size_t)(cols - 4) * rows;
  k_scan_cols<false><<<dim3(nblocks, 4), SCAN_T, 0, c->stream>>>(last4, rows, n, tot, nblocks);
  k_scan_totals<<<4, SCAN_T, 0, c->stream>>>(tot, nblocks);
  k_scan_cols<true><<<dim3(nblocks, 4), SCAN_T, 0, c->stream>>>(last4, rows, n, tot, nblocks);
  k_accum_finalize<<<(n + 255) / 256, 256, 0, c->stream>>>(accum, rows, cols, R0_WG_USER_ACCUM_SPLIT, n);
  R0_CUDA(cudaFreeAsync(tot, c->stream));
  count_launch(c, 5);
  R0_CUDA(cudaGetLastError());
  if (sync_check) check_errors(c, t, "accum");
}

extern "C" {

r0b200_err r0b200_trace_upload(r0b200_ctx* ctx, const r0b200_preflight_trace* trace_host, uint32_t cycles,
                               r0b200_trace** out) {
  R0_API_BEGIN
  R0_CHECK(ctx != nullptr && out != nullptr, "trace_upload: null argument");
  R0_CUDA(cudaSetDevice(ctx->device));
  *out = r0_trace_upload(ctx, trace_host, cycles, ctx->stream);
  R0_API_END
}
void r0b200_trace_free(r0b200_trace* trace) {
  if (!trace) return;
  cudaSetDevice(trace->c->device);
  r0_trace_free(trace);
}
r0b200_err r0b200_witgen_rv32im(r0b200_ctx* ctx, uint32_t mode, r0b200_trace* trace, uint32_t* global, uint32_t* data) {
  R0_API_BEGIN
  R0_CHECK(ctx != nullptr && trace != nullptr && trace->c == ctx, "witgen: trace belongs to another context");
  R0_CHECK(mode <= 2, "witgen: unknown step mode");
  R0_CUDA(cudaSetDevice(ctx->device));
  r0_witgen_rv32im(ctx, trace, global, data, true);
  R0_API_END
}
r0b200_err r0b200_accum_rv32im(r0b200_ctx* ctx, r0b200_trace* trace, uint32_t* data, uint32_t* accum, uint32_t* global,
                               uint32_t* mix) {
  R0_API_BEGIN
  R0_CHECK(ctx != nullptr && trace != nullptr && trace->c == ctx, "accum: trace belongs to another context");
  R0_CUDA(cudaSetDevice(ctx->device));
  r0_accum_rv32im(ctx, trace, data, accum, global, mix, true);
  R0_API_END
}

}  // extern "C"
