// Integer pipe micro-benchmark for sm_100a: issue rate (warp instructions / clk / SM) of the instructions the BabyBear
// kernels are made of, alone and in pairs. Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/int_pipes tools/ubench/int_pipes.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ITERS = 4096;
constexpr int CH = 8;  // independent chains per thread

template <int OP>
__global__ void k(uint32_t* out, uint32_t a0, uint32_t b0) {
  uint32_t x[CH], y[CH], z[CH], w[CH], v[CH];
#pragma unroll
  for (int i = 0; i < CH; i++) {
    x[i] = a0 + threadIdx.x + i;
    y[i] = b0 ^ (threadIdx.x * 7 + i);
    z[i] = x[i] * 3 + 1;
    w[i] = y[i] * 5 + 2;
    v[i] = y[i] * 9 + 4;
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
      } else if (OP == 13) {  // IMAD + IADD3
        asm volatile("mad.lo.u32 %0, %0, %1, %1;" : "+r"(x[i]) : "r"(y[i]));
        asm volatile("add.u32 %0, %0, %1;" : "+r"(y[i]) : "r"(a0));
      } else if (OP == 14) {  // IMAD + LEA
        asm volatile("mad.lo.u32 %0, %0, %1, %1;" : "+r"(x[i]) : "r"(y[i]));
        y[i] = (y[i] << 5) + a0;
      } else if (OP == 15) {  // IMAD + IADD3 + VIADDMNMX
        asm volatile("mad.lo.u32 %0, %0, %1, %1;" : "+r"(x[i]) : "r"(y[i]));
        asm volatile("add.u32 %0, %0, %1;" : "+r"(y[i]) : "r"(a0));
        uint32_t t;
        asm volatile("add.u32 %0, %1, 0x87ffffff;" : "=r"(t) : "r"(y[i]));
        asm volatile("min.u32 %0, %0, %1;" : "+r"(y[i]) : "r"(t));
      } else if (OP == 16) {  // IMAD + 2 IADD3
        asm volatile("mad.lo.u32 %0, %0, %1, %1;" : "+r"(x[i]) : "r"(y[i]));
        asm volatile("add.u32 %0, %0, %1;" : "+r"(y[i]) : "r"(a0));
        asm volatile("add.u32 %0, %0, %1;" : "+r"(z[i]) : "r"(b0));
      } else if (OP == 17) {  // IMAD.HI + 3 IADD3
        asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(x[i]) : "r"(b0));
        asm volatile("add.u32 %0, %0, %1;" : "+r"(y[i]) : "r"(a0));
        asm volatile("add.u32 %0, %0, %1;" : "+r"(z[i]) : "r"(b0));
        asm volatile("add.u32 %0, %0, %1;" : "+r"(w[i]) : "r"(b0));
      } else if (OP == 18) {  // Montgomery product, m = lo + ((17 lo) << 27) as two LEA
        uint64_t t = (uint64_t)x[i] * y[i];
        uint32_t lo = (uint32_t)t, t17, m;
        asm volatile("shl.b32 %0, %1, 4;\n\tadd.u32 %0, %0, %1;" : "=&r"(t17) : "r"(lo));
        asm volatile("shl.b32 %0, %1, 27;\n\tadd.u32 %0, %0, %2;" : "=&r"(m) : "r"(t17), "r"(lo));
        uint32_t h = __umulhi(m, 0x78000001u);
        uint32_t r = (uint32_t)(t >> 32) - h;
        uint32_t r2 = r + 0x78000001u;
        x[i] = r < r2 ? r : r2;
      } else if (OP == 19) {  // IMAD + SHF (funnel shift)
        asm volatile("mad.lo.u32 %0, %0, %1, %1;" : "+r"(x[i]) : "r"(y[i]));
        asm volatile("shf.l.wrap.b32 %0, %0, %1, 7;" : "+r"(z[i]) : "r"(a0));
      } else if (OP == 20) {  // IMAD + LOP3
        asm volatile("mad.lo.u32 %0, %0, %1, %1;" : "+r"(x[i]) : "r"(y[i]));
        asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(z[i]) : "r"(a0), "r"(w[i]));
      } else if (OP == 21) {  // LEA alone
        y[i] = (y[i] << 5) + a0;
        asm volatile("" : "+r"(y[i]));
      } else if (OP == 22) {  // SHF alone
        asm volatile("shf.l.wrap.b32 %0, %0, %1, 7;" : "+r"(z[i]) : "r"(a0));
      } else if (OP == 23) {  // LOP3 alone (3-input xor)
        asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(z[i]) : "r"(a0), "r"(w[i]));
      } else if (OP == 24) {  // VIMNMX alone
        asm volatile("min.u32 %0, %0, %1;" : "+r"(z[i]) : "r"(w[i]));
        asm volatile("max.u32 %0, %0, %1;" : "+r"(w[i]) : "r"(a0));
      } else if (OP == 25) {  // IMAD.WIDE + 2 VIADDMNMX
        uint64_t t;
        asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(t) : "r"(x[i]), "r"(y[i]));
        x[i] = (uint32_t)t ^ (uint32_t)(t >> 32);
        uint32_t u;
        asm volatile("add.u32 %0, %1, 0x87ffffff;" : "=r"(u) : "r"(z[i]));
        asm volatile("min.u32 %0, %0, %1;" : "+r"(z[i]) : "r"(u));
        asm volatile("add.u32 %0, %1, 0x87ffffff;" : "=r"(u) : "r"(w[i]));
        asm volatile("min.u32 %0, %0, %1;" : "+r"(w[i]) : "r"(u));
      } else if (OP == 26) {  // IMAD.HI + 2 VIADDMNMX + 2 IADD3 : 4 fma cycles, 4 alu-only cycles, 2 flexible
        asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(x[i]) : "r"(b0));
        uint32_t u;
        asm volatile("add.u32 %0, %1, 0x87ffffff;" : "=r"(u) : "r"(z[i]));
        asm volatile("min.u32 %0, %0, %1;" : "+r"(z[i]) : "r"(u));
        asm volatile("add.u32 %0, %1, 0x87ffffff;" : "=r"(u) : "r"(w[i]));
        asm volatile("min.u32 %0, %0, %1;" : "+r"(w[i]) : "r"(u));
        asm volatile("add.u32 %0, %0, %1;" : "+r"(y[i]) : "r"(a0));
        asm volatile("add.u32 %0, %0, %1;" : "+r"(v[i]) : "r"(b0));
      } else if (OP == 27) {  // IMAD.HI + 2 VIADDMNMX (4 fma cycles, 4 alu cycles)
        asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(x[i]) : "r"(b0));
        uint32_t u;
        asm volatile("add.u32 %0, %1, 0x87ffffff;" : "=r"(u) : "r"(z[i]));
        asm volatile("min.u32 %0, %0, %1;" : "+r"(z[i]) : "r"(u));
        asm volatile("add.u32 %0, %1, 0x87ffffff;" : "=r"(u) : "r"(w[i]));
        asm volatile("min.u32 %0, %0, %1;" : "+r"(w[i]) : "r"(u));
      } else if (OP == 28) {  // IMAD.HI + 4 IADD3
        asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(x[i]) : "r"(b0));
        asm volatile("add.u32 %0, %0, %1;" : "+r"(y[i]) : "r"(a0));
        asm volatile("add.u32 %0, %0, %1;" : "+r"(z[i]) : "r"(b0));
        asm volatile("add.u32 %0, %0, %1;" : "+r"(w[i]) : "r"(b0));
        asm volatile("add.u32 %0, %0, %1;" : "+r"(v[i]) : "r"(a0));
      } else if (OP == 29) {  // 64-bit add (IADD3 + IADD3.X carry pair)
        asm volatile("add.cc.u32 %0, %0, %2;\n\taddc.u32 %1, %1, %3;" : "+r"(x[i]), "+r"(y[i]) : "r"(a0), "r"(b0));
      } else if (OP == 30) {  // ISETP + SEL
        uint32_t u = z[i] + a0;
        z[i] = (u >= 0x78000001u) ? u - 0x78000001u : u;
        asm volatile("" : "+r"(z[i]));
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
  for (int i = 0; i < CH; i++) s ^= x[i] ^ y[i] ^ z[i] ^ w[i] ^ v[i];
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
  run<13>("IMAD + IADD3", 2, out, p.multiProcessorCount, ghz);
  run<14>("IMAD + LEA", 2, out, p.multiProcessorCount, ghz);
  run<15>("IMAD + IADD3 + VIADDMNMX", 3, out, p.multiProcessorCount, ghz);
  run<16>("IMAD + 2 IADD3", 3, out, p.multiProcessorCount, ghz);
  run<17>("IMAD.HI + 3 IADD3", 4, out, p.multiProcessorCount, ghz);
  run<28>("IMAD.HI + 4 IADD3", 5, out, p.multiProcessorCount, ghz);
  run<27>("IMAD.HI + 2 VIADDMNMX", 3, out, p.multiProcessorCount, ghz);
  run<26>("IMAD.HI + 2 VIADDMNMX + 2 IADD3", 5, out, p.multiProcessorCount, ghz);
  run<25>("IMAD.WIDE(+LOP3) + 2 VIADDMNMX", 4, out, p.multiProcessorCount, ghz);
  run<18>("Montgomery product, m by two LEA", 1, out, p.multiProcessorCount, ghz);
  run<19>("IMAD + SHF", 2, out, p.multiProcessorCount, ghz);
  run<20>("IMAD + LOP3", 2, out, p.multiProcessorCount, ghz);
  run<21>("LEA", 1, out, p.multiProcessorCount, ghz);
  run<22>("SHF", 1, out, p.multiProcessorCount, ghz);
  run<23>("LOP3 (3-input)", 1, out, p.multiProcessorCount, ghz);
  run<24>("VIMNMX (min + max)", 2, out, p.multiProcessorCount, ghz);
  run<29>("IADD3 + IADD3.X (64-bit add)", 2, out, p.multiProcessorCount, ghz);
  run<30>("add + ISETP + SEL-style reduce", 1, out, p.multiProcessorCount, ghz);
  run<10>("shl + add", 2, out, p.multiProcessorCount, ghz);
  run<11>("xor + and (LOP3)", 2, out, p.multiProcessorCount, ghz);
  return 0;
}
