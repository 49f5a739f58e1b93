// mma.sync m16n8k32 u8 x u8 -> s32 issue rate on sm_100a, alone and next to a stream of IMAD.WIDE (does the legacy
// integer tensor path share anything with the integer pipes?). Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ITERS = 2048;

__device__ __forceinline__ void mma_u8(int (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

template <int MODE>
__global__ void k(uint32_t* out, uint32_t seed) {
  uint32_t a[4], b[2];
  int c[8][4];
  for (int i = 0; i < 4; i++) a[i] = seed * (threadIdx.x + i + 1);
  for (int i = 0; i < 2; i++) b[i] = seed ^ (threadIdx.x * 31 + i);
  for (int j = 0; j < 8; j++)
    for (int i = 0; i < 4; i++) c[j][i] = 0;
  uint64_t w[8];
  for (int j = 0; j < 8; j++) w[j] = seed + j;
  uint32_t x = seed | 1;
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int j = 0; j < 8; j++) {
      if (MODE != 1) mma_u8(c[j], a, b);
      if (MODE >= 1) {
        uint32_t lo = (uint32_t)w[j], hi = (uint32_t)(w[j] >> 32);
        asm volatile("mad.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.u32 %1, %2, %3, %1;" : "+r"(lo), "+r"(hi) : "r"(x), "r"(a[0]));
        w[j] = ((uint64_t)hi << 32) | lo;
      }
    }
  }
  uint32_t s = 0;
  for (int j = 0; j < 8; j++) {
    for (int i = 0; i < 4; i++) s ^= (uint32_t)c[j][i];
    s ^= (uint32_t)w[j] ^ (uint32_t)(w[j] >> 32);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
float run(uint32_t* out, int sms) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  k<MODE><<<sms * 8, 256>>>(out, 12345u);
  cudaEventRecord(e0);
  k<MODE><<<sms * 8, 256>>>(out, 12345u);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  return ms;
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  int khz = 0;
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  const double ghz = khz / 1e6;
  uint32_t* out;
  cudaMalloc(&out, (size_t)p.multiProcessorCount * 8 * 256 * 4);
  const int sms = p.multiProcessorCount;
  const double warps = sms * 8 * 8.0, per_warp = ITERS * 8.0;
  const float t0 = run<0>(out, sms), t1 = run<1>(out, sms), t2 = run<2>(out, sms);
  auto cyc = [&](float ms) { return ms * 1e-3 * ghz * 1e9 * sms * 4 / (warps * per_warp); };   // SMSP cycles per warp-level op
  printf("%s: mma.m16n8k32.u8 alone %.3f ms = %.2f cycles per MMA per sub-partition (%.0f int8 MAC/clk/SM)\n", p.name, t0, cyc(t0),
         4096.0 * 4 / cyc(t0));
  printf("IMAD.WIDE accumulate alone %.3f ms = %.2f cycles each\n", t1, cyc(t1));
  printf("one MMA + one IMAD.WIDE per iteration %.3f ms = %.2f cycles per pair (sum of the two alone: %.2f)\n", t2, cyc(t2), cyc(t0) + cyc(t1));
  return 0;
}
