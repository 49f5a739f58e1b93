// Integer pipe micro-benchmark for sm_100a: issue rate (warp instructions / clk / SM) of the instructions the BabyBear
// kernels are made of, alone and in pairs. Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/int_pipes tools/ubench/int_pipes.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ITERS = 4096;
constexpr int CH = 8;  // independent chains per thread

template <int OP>
__global__ void k(uint32_t* out, uint32_t a0, uint32_t b0) {
  uint32_t x[CH], y[CH];
#pragma unroll
  for (int i = 0; i < CH; i++) {
    x[i] = a0 + threadIdx.x + i;
    y[i] = b0 ^ (threadIdx.x * 7 + i);
  }
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int i = 0; i < CH; i++) {
      if (OP == 0) {  // IMAD (32-bit lo)
        asm volatile("mad.lo.u32 %0, %0, %1, %1;" : "+r"(x[i]) : "r"(y[i]));
      } else if (OP == 1) {  // IMAD.WIDE (mul.wide + use both halves)
        uint64_t t;
        asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(t) : "r"(x[i]), "r"(y[i]));
        x[i] = (uint32_t)t ^ (uint32_t)(t >> 32);
      } else if (OP == 2) {  // IMAD.HI
        asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(x[i]) : "r"(y[i]));
      } else if (OP == 3) {  // IADD3
        asm volatile("add.u32 %0, %0, %1;" : "+r"(x[i]) : "r"(y[i]));
      } else if (OP == 4) {  // VIADDMNMX (add + min)
        uint32_t t;
        asm volatile("add.u32 %0, %1, 0x87ffffff;" : "=r"(t) : "r"(x[i]));
        asm volatile("min.u32 %0, %0, %1;" : "+r"(x[i]) : "r"(t));
      } else if (OP == 5) {  // IMAD + VIADDMNMX pair (fma + alu)
        asm volatile("mad.lo.u32 %0, %0, %1, %1;" : "+r"(x[i]) : "r"(y[i]));
        uint32_t t;
        asm volatile("add.u32 %0, %1, 0x87ffffff;" : "=r"(t) : "r"(x[i]));
        asm volatile("min.u32 %0, %0, %1;" : "+r"(x[i]) : "r"(t));
      } else if (OP == 6) {  // IADD3 + VIADDMNMX pair (alu + alu)
        asm volatile("add.u32 %0, %0, %1;" : "+r"(x[i]) : "r"(y[i]));
        uint32_t t;
        asm volatile("add.u32 %0, %1, 0x87ffffff;" : "=r"(t) : "r"(x[i]));
        asm volatile("min.u32 %0, %0, %1;" : "+r"(x[i]) : "r"(t));
      } else if (OP == 7) {  // IMAD.WIDE with 64-bit accumulate (carry-chained pair)
        asm volatile("mad.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.u32 %1, %2, %3, %1;" : "+r"(x[i]), "+r"(y[i]) : "r"(a0), "r"(b0));
      } else if (OP == 8) {  // full Montgomery product
        uint64_t t = (uint64_t)x[i] * y[i];
        uint32_t m = (uint32_t)t * 0x88000001u;
        uint32_t h = __umulhi(m, 0x78000001u);
        uint32_t r = (uint32_t)(t >> 32) - h;
        uint32_t r2 = r + 0x78000001u;
        x[i] = r < r2 ? r : r2;
      } else if (OP == 9) {  // Montgomery product, m = lo * P^-1 as shifts + 3-input add (P^-1 = 2^31 + 2^27 + 1)
        uint64_t t = (uint64_t)x[i] * y[i];
        uint32_t lo = (uint32_t)t, s27, s31, m;
        asm volatile("shl.b32 %0, %1, 27;" : "=r"(s27) : "r"(lo));
        asm volatile("shl.b32 %0, %1, 31;" : "=r"(s31) : "r"(lo));
        asm volatile("add.u32 %0, %1, %2;" : "=r"(m) : "r"(lo), "r"(s27));
        asm volatile("add.u32 %0, %0, %1;" : "+r"(m) : "r"(s31));
        uint32_t h = __umulhi(m, 0x78000001u);
        uint32_t r = (uint32_t)(t >> 32) - h;
        uint32_t r2 = r + 0x78000001u;
        x[i] = r < r2 ? r : r2;
      } else if (OP == 12) {  // Montgomery product, m = (lo + (lo << 27)) ^ (lo << 31): LEA + SHF + LOP3, no IMAD
        uint64_t t = (uint64_t)x[i] * y[i];
        uint32_t lo = (uint32_t)t;
        uint32_t m = (lo + (lo << 27)) ^ (lo << 31);
        uint32_t h = __umulhi(m, 0x78000001u);
        uint32_t r = (uint32_t)(t >> 32) - h;
        uint32_t r2 = r + 0x78000001u;
        x[i] = r < r2 ? r : r2;
      } else if (OP == 10) {  // SHF / shl alone
        asm volatile("shl.b32 %0, %0, 3;" : "+r"(x[i]));
        asm volatile("add.u32 %0, %0, %1;" : "+r"(x[i]) : "r"(y[i]));
      } else if (OP == 11) {  // LOP3
        asm volatile("xor.b32 %0, %0, %1;" : "+r"(x[i]) : "r"(y[i]));
        asm volatile("and.b32 %0, %0, 0x7fffffff;" : "+r"(x[i]));
      }
    }
  }
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < CH; i++) s ^= x[i] ^ y[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int OP>
void run(const char* name, int instr_per_iter, uint32_t* out, int sms, double clock_ghz) {
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  const int blocks = sms * 2, threads = 1024;
  k<OP><<<blocks, threads>>>(out, 12345u, 678u);
  cudaEventRecord(a);
  k<OP><<<blocks, threads>>>(out, 12345u, 678u);
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms;
  cudaEventElapsedTime(&ms, a, b);
  const double warp_instr = (double)blocks * threads / 32 * ITERS * CH * instr_per_iter;
  const double per_clk_sm = warp_instr / (ms * 1e-3 * clock_ghz * 1e9) / sms;
  printf("%-44s %8.3f ms  %6.2f warp-instr/clk/SM (%.2f per SMSP)\n", name, ms, per_clk_sm, per_clk_sm / 4);
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  int khz = 0;
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  const double ghz = khz / 1e6;
  uint32_t* out;
  cudaMalloc(&out, (size_t)p.multiProcessorCount * 2 * 1024 * 4);
  printf("%s, %d SMs, max clock %.3f GHz (rates assume the max clock)\n", p.name, p.multiProcessorCount, ghz);
  run<0>("IMAD (mad.lo)", 1, out, p.multiProcessorCount, ghz);
  run<1>("IMAD.WIDE (+1 LOP3)", 2, out, p.multiProcessorCount, ghz);
  run<2>("IMAD.HI", 1, out, p.multiProcessorCount, ghz);
  run<3>("IADD3", 1, out, p.multiProcessorCount, ghz);
  run<4>("VIADDMNMX (add+min fused)", 1, out, p.multiProcessorCount, ghz);
  run<5>("IMAD + VIADDMNMX", 2, out, p.multiProcessorCount, ghz);
  run<6>("IADD3 + VIADDMNMX", 2, out, p.multiProcessorCount, ghz);
  run<7>("IMAD.WIDE accumulate (cc pair)", 1, out, p.multiProcessorCount, ghz);
  run<8>("Montgomery product (5 instr)", 1, out, p.multiProcessorCount, ghz);
  run<9>("Montgomery product, m by shift-add", 1, out, p.multiProcessorCount, ghz);
  run<12>("Montgomery product, m by LEA+SHF+LOP3", 1, out, p.multiProcessorCount, ghz);
  run<10>("shl + add", 2, out, p.multiProcessorCount, ghz);
  run<11>("xor + and (LOP3)", 2, out, p.multiProcessorCount, ghz);
  return 0;
}
