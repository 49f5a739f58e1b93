// Internal context shared by the kernels' host launchers (not part of the public C ABI; see include/r0b200.h).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <map>
#include <stdexcept>
#include <string>
#include <vector>

#include <nvtx3/nvToolsExt.h>

#include "fp.cuh"

namespace r0 {

struct CudaError : std::runtime_error {
  using std::runtime_error::runtime_error;
};

#define R0_CUDA(expr)                                                                                     \
  do {                                                                                                    \
    cudaError_t _e = (expr);                                                                              \
    if (_e != cudaSuccess)                                                                                \
      throw r0::CudaError(std::string(#expr) + " failed: " + cudaGetErrorString(_e) + " (" __FILE__ ":" + \
                          std::to_string(__LINE__) + ")");                                                \
  } while (0)

#define R0_CHECK(cond, msg)                                 \
  do {                                                      \
    if (!(cond)) throw std::invalid_argument(std::string(msg)); \
  } while (0)

constexpr int MAX_LG = 24;  // largest supported NTT size 2^24 (MAX_CYCLES_PO2 = 22 -> 4N = 2^24, zkp/src/lib.rs:35-38)

struct Tables {
  // two-level twiddle tables for the order-2^24 roots: w^E = hi[E >> 12] * lo[E & 4095]; dir 0 = ROU_REV, 1 = ROU_FWD
  uint32_t* tw_lo[2];
  uint32_t* tw_hi[2];
  uint32_t* tw_row[2];  // inter-step twiddle rows (ntt.cu: row_off)
  // powers of three, same two-level split: 3^e = p3_hi[e >> 12] * p3_lo[e & 4095]
  uint32_t* p3_lo;
  uint32_t* p3_hi;
  // (2^k)^-1 * 3^(4096 j), lazily built per k ; and plain (2^k)^-1
  uint32_t* p3_hi_scaled[MAX_LG + 1];
  uint32_t ninv[MAX_LG + 1];
};

struct Ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t copy_stream = nullptr;  // host->device witness uploads, overlapped with compute on `stream`
  // Every context allocates from its OWN stream-ordered pool: with the device's default pool, two contexts proving
  // on one GPU at the same time hand freed blocks to each other, and the dependencies the allocator inserts for that
  // serialise their streams (measured at po2 = 20: 567 ms per segment with two contexts instead of 123 with one).
  cudaMemPool_t pool = nullptr;
  int sm_count = 148;
  Tables tab{};
  uint64_t launches = 0;       // kernels launched through this context (bench.py's gpu_launches)
  size_t bytes_allocated = 0;  // MemoryTracker analogue (hal/mod.rs:292-317)
  size_t bytes_peak = 0;
  std::map<void*, size_t> alloc_sizes;  // r0b200_alloc'd blocks, so r0b200_free can decrement bytes_allocated
  cudaEvent_t ev_start = nullptr, ev_stop = nullptr;
  // optional per-phase device timing (r0b200_profile_begin / _end): event pairs recorded on `stream` around each op
  bool profiling = false;
  struct PhaseRec {
    const char* name;
    cudaEvent_t a, b;
    double bytes;  // algorithmic bytes of the op (SURVEY 8d formulas), 0 if not stated
  };
  std::vector<PhaseRec> phase_log;
  std::vector<cudaEvent_t> event_pool;
  // auxiliary streams / events of the concurrent-tiled eval_check mode (created on demand, gen/eval_check_*.cu)
  std::vector<cudaStream_t> aux_streams;
  std::vector<cudaEvent_t> aux_events;
};

inline cudaError_t r0_malloc_async(Ctx* c, void** p, size_t bytes, cudaStream_t stream) {
  return cudaMallocFromPoolAsync(p, bytes, c->pool, stream);
}
template <typename T>
inline cudaError_t r0_malloc_async(Ctx* c, T** p, size_t bytes, cudaStream_t stream) {
  return cudaMallocFromPoolAsync(reinterpret_cast<void**>(p), bytes, c->pool, stream);
}

// RAII phase marker used by the launchers; free when profiling is off.
struct PhaseScope {
  Ctx* c;
  cudaEvent_t b = nullptr;
  PhaseScope(Ctx* c_, const char* name, double bytes = 0) : c(c_) {
    if (!c->profiling || !name) return;
    cudaEvent_t ev[2];
    for (int i = 0; i < 2; i++) {
      if (c->event_pool.empty()) {
        cudaEventCreate(&ev[i]);
      } else {
        ev[i] = c->event_pool.back();
        c->event_pool.pop_back();
      }
    }
    cudaEventRecord(ev[0], c->stream);
    b = ev[1];
    c->phase_log.push_back({name, ev[0], ev[1], bytes});
  }
  ~PhaseScope() {
    if (b) cudaEventRecord(b, c->stream);
  }
};

// NVTX range with the reference's scope! names (risc0/core/src/perf.rs:22-73: scope!(name) pushes an NVTX range for
// the enclosing block), so nsys timelines of this backend line up with the reference's. Header-only NVTX3: a no-op
// unless a profiler injects itself.
struct NvtxRange {
  explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
  NvtxRange(const NvtxRange&) = delete;
  NvtxRange& operator=(const NvtxRange&) = delete;
};

inline void count_launch(Ctx* c, uint64_t n = 1) { c->launches += n; }

}  // namespace r0

// The opaque handle of the C ABI.
struct r0b200_ctx : r0::Ctx {};

// C ABI error convention (same as risc0/sys/src/lib.rs:53-75): NULL = success, otherwise a strdup'd message the
// caller releases with r0b200_free_error / libc free. C++ exceptions never cross the boundary.
#define R0_API_BEGIN try {
#define R0_API_END                       \
  }                                      \
  catch (const std::exception& e) {      \
    return strdup(e.what());             \
  }                                      \
  catch (...) {                          \
    return strdup("unknown C++ exception"); \
  }                                      \
  return nullptr;
