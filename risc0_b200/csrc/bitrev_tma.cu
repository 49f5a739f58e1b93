// Hal::batch_bit_reverse (risc0/zkp/src/hal/cpu.rs:352-360, core/ntt.rs:26-35) as a TMA kernel for sm_100a.
//
// The op is pure data movement (8 bytes per element), so it belongs to the copy engines, not to the SMs' load/store
// path: an index p = (a : 5 bits | mid | b : 5 bits) of a 2^k column goes to (brev b | brev mid | brev a). One CTA owns
// the pair of 32 x 32-word tiles {mid, brev(mid)}: it pulls both with one `cp.async.bulk.tensor.4d` each (TMA, the
// 128-byte rows of a tile are 2^(k-5) words apart; completion on an mbarrier), transposes / bit-reverses them inside
// shared memory (128-byte hardware swizzle on both sides instead of padding), and pushes them back to the swapped
// positions with two TMA tensor stores. No thread computes a global address, the in-place swap needs no second buffer,
// and ~16 CTAs per SM keep > 100 KB of loads in flight.
//
// The column is described to the TMA unit as a rank-4 tensor {b: 32 | mid: 2^(k-10) | a: 32 | column}; the same map
// serves loads and stores. k < 10 (or a misaligned slice) stays on the register-path kernel in ntt.cu.
#include <cuda.h>

#include "ctx.h"

namespace r0 {
namespace {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  } while (!done);
}
__device__ __forceinline__ void tma_load_4d(void* smem, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(
          smem_u32(smem)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* map, const void* smem, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(map)),
               "r"(smem_u32(smem)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}

// byte offset of word (row, col) of a 32 x 32-word tile under CU_TENSOR_MAP_SWIZZLE_128B: the 16-byte chunk index of
// a 128-byte row is XORed with (row mod 8)
__device__ __forceinline__ uint32_t swz(uint32_t row, uint32_t col) {
  return row * 128u + ((((col >> 2) ^ (row & 7u)) << 4) | ((col & 3u) << 2));
}

constexpr int TILE_BYTES = 32 * 32 * 4;

__global__ void __launch_bounds__(256) bit_reverse_tma_kernel(const __grid_constant__ CUtensorMap map, int midbits) {
  extern __shared__ uint8_t smem_raw[];
  // the 128-byte swizzle pattern repeats every 1024 bytes: tiles must sit on 1024-byte boundaries
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* in0 = base;                    // tile `mid`
  uint8_t* in1 = base + TILE_BYTES;       // tile `rmid`
  uint8_t* out0 = base + 2 * TILE_BYTES;  // permuted in0, goes to position rmid
  uint8_t* out1 = base + 3 * TILE_BYTES;  // permuted in1, goes to position mid
  __shared__ uint64_t bar;

  const uint32_t mid = blockIdx.x;
  const uint32_t rmid = midbits ? (__brev(mid) >> (32 - midbits)) : 0u;
  if (mid > rmid) return;
  const bool pair = mid != rmid;
  const int col = blockIdx.y;

  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(&bar, pair ? 2 * TILE_BYTES : TILE_BYTES);
    tma_load_4d(in0, &map, &bar, 0, (int)mid, 0, col);
    if (pair) tma_load_4d(in1, &map, &bar, 0, (int)rmid, 0, col);
  }
  mbar_wait(&bar, 0);

  // element (a, b) of a source tile lands at (brev b, brev a) of its destination tile. Lanes run over a (the source
  // ROW): the write side is then one 128-byte destination row per warp (conflict-free), the read side 4-way conflicted.
  for (uint32_t w = threadIdx.x; w < 1024; w += 256) {
    const uint32_t a = w & 31u, b = w >> 5;
    const uint32_t y = __brev(b) >> 27, x = __brev(a) >> 27;
    const uint32_t src = swz(a, b), dst = swz(y, x);
    *reinterpret_cast<uint32_t*>(out0 + dst) = *reinterpret_cast<const uint32_t*>(in0 + src);
    if (pair) *reinterpret_cast<uint32_t*>(out1 + dst) = *reinterpret_cast<const uint32_t*>(in1 + src);
  }
  // make the generic-proxy writes visible to the async proxy (TMA) before the stores read shared memory
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (threadIdx.x == 0) {
    tma_store_4d(&map, out0, 0, (int)rmid, 0, col);
    if (pair) tma_store_4d(&map, out1, 0, (int)mid, 0, col);
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // shared memory must outlive the reads of the stores
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled() {
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
      p = nullptr;
    return reinterpret_cast<EncodeTiledFn>(p);
  }();
  return fn;
}

}  // namespace

// true if the TMA path took the call; false = caller falls back to the register-path kernel
bool r0_bit_reverse_tma(Ctx* c, uint32_t* io, size_t count, int k) {
  static const bool off = getenv("R0B200_BITREV_TMA") && atoi(getenv("R0B200_BITREV_TMA")) == 0;
  if (off || k < 10 || (reinterpret_cast<uintptr_t>(io) & 15) != 0) return false;
  EncodeTiledFn enc = encode_tiled();
  if (!enc) return false;
  const int smem_bytes = 4 * TILE_BYTES + 1024;   // 17 KB: under the 48 KB that needs no opt-in
  const int midbits = k - 10;
  for (size_t c0 = 0; c0 < count; c0 += 65535) {
    const size_t nc = count - c0 < 65535 ? count - c0 : 65535;
    CUtensorMap map;
    const cuuint64_t dims[4] = {32, cuuint64_t(1) << midbits, 32, nc};
    const cuuint64_t strides[3] = {128, (cuuint64_t(1) << (k - 5)) * 4, (cuuint64_t(1) << k) * 4};   // bytes, dims 1..3
    const cuuint32_t box[4] = {32, 1, 32, 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    const CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_UINT32, 4, io + (c0 << k), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      R0_CHECK(c0 == 0, "bit_reverse: cuTensorMapEncodeTiled failed after the first column chunk");
      return false;
    }
    bit_reverse_tma_kernel<<<dim3(1u << midbits, (unsigned)nc), 256, smem_bytes, c->stream>>>(map, midbits);
    count_launch(c);
  }
  R0_CUDA(cudaGetLastError());
  return true;
}

}  // namespace r0
