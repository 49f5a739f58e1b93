// Batched BabyBear NTT / iNTT / x4 low-degree extension for sm_100a.
//
// Replaces: Hal::batch_interpolate_ntt (risc0/zkp/src/hal/cpu.rs:342-350 -> core/ntt.rs:232-282),
//           Hal::zk_shift (cpu.rs:395-408), Hal::batch_expand_into_evaluate_ntt (cpu.rs:305-340 ->
//           core/ntt.rs:284-343), Hal::batch_bit_reverse (cpu.rs:352-360); reference GPU path: per-column host loops
//           over sppark NTTs (risc0/sys/kernels/zkp/cuda/supra/ntt.cu:34-152).
//
// Design (modelled and checked against the oracle in tools/ntt_model.py):
//  * four-step split n = 2^k1 * 2^k2: one "strided" pass over tiles of 2^k1 rows x T adjacent columns and one
//    "contiguous" pass over 2^k2-element tiles; n <= 2^12 needs only the contiguous pass. All columns of a batch are
//    processed by one launch (grid.y = columns), not a host loop.
//  * inside a tile the 2^m-point transform runs in shared memory as radix-16 register steps (4 butterfly layers per
//    shared-memory round trip, constant twiddles from the constant bank, inter-step twiddles from a 16 KB table),
//    padded (i + i/16) so every step is bank-conflict free or 2-way at worst.
//  * everything that would be an extra HBM pass in the reference is fused into an epilogue: the 1/n scale and the
//    zk shift 3^brev(i) into the iNTT's last store, the expand-by-4 into the LDE's first load (the two skipped
//    butterfly layers become a template parameter), the inter-pass twiddle into the producing pass.
//  * algorithmic HBM bytes: iNTT 8 B/element, LDE 20 B/input element; two-pass sizes move 16 / 52 B unless the
//    caller launches column groups that fit the 126 MB L2 (`cols_per_launch`), which keeps the intermediate on chip.
#include "ctx.h"

namespace r0 {

__constant__ uint32_t c_w16[2][8];  // w_16^j, j < 8, Montgomery; [0] = ROU_REV[4], [1] = ROU_FWD[4]

__device__ __forceinline__ int pad16(int i) { return i + (i >> 4); }

// w_{2^m}^e for m <= 12, e < 2^m : a single lookup in the "hi" table (stride 2^(12-m))
__device__ __forceinline__ uint32_t tw_small(const uint32_t* __restrict__ hi, int m, uint32_t e) {
  return __ldg(hi + (e << (12 - m)));
}
// w_{2^k}^e for k <= 24
__device__ __forceinline__ uint32_t tw_big(const uint32_t* __restrict__ lo, const uint32_t* __restrict__ hi, int k,
                                           uint32_t e) {
  uint32_t E = e << (24 - k);
  uint32_t h = __ldg(hi + (E >> 12));
  uint32_t l = E & 4095u;
  return l ? fp_mul(h, __ldg(lo + l)) : h;
}

// ---- in-register radix-2^A transforms with compile-time twiddles --------------------------------------------
template <int A>
__device__ __forceinline__ void radix_dif(uint32_t (&v)[1 << A]) {  // natural in, bit-reversed out, ROU_REV
#pragma unroll
  for (int s = A; s >= 1; --s) {
    const int half = 1 << (s - 1);
#pragma unroll
    for (int b = 0; b < (1 << A); b += 2 * half) {
#pragma unroll
      for (int i = 0; i < half; ++i) {
        uint32_t x = v[b + i], y = v[b + i + half];
        v[b + i] = fp_add(x, y);
        uint32_t d = fp_sub(x, y);
        v[b + i + half] = (i == 0) ? d : fp_mul(d, c_w16[0][i << (4 - s)]);
      }
    }
  }
}
template <int A, int SKIP>
__device__ __forceinline__ void radix_dit(uint32_t (&v)[1 << A]) {  // bit-reversed in, natural out, ROU_FWD
#pragma unroll
  for (int s = SKIP + 1; s <= A; ++s) {
    const int half = 1 << (s - 1);
#pragma unroll
    for (int b = 0; b < (1 << A); b += 2 * half) {
#pragma unroll
      for (int i = 0; i < half; ++i) {
        uint32_t x = v[b + i];
        uint32_t y = (i == 0) ? v[b + i + half] : fp_mul(v[b + i + half], c_w16[1][i << (4 - s)]);
        v[b + i] = fp_add(x, y);
        v[b + i + half] = fp_sub(x, y);
      }
    }
  }
}

// ---- one radix step of the in-shared-memory transform -------------------------------------------------------
// Logical element (h, l), h < 2^m, l < T lives at s[pad16(h) * T + l].
template <int A>
__device__ __forceinline__ void step_dif(uint32_t* s, int m, int rem, int lgT, const uint32_t* __restrict__ tw_hi) {
  const int T = 1 << lgT;
  const int items = (1 << (m - A)) << lgT;
  const int lgstride = rem - A;
  for (int w = threadIdx.x; w < items; w += blockDim.x) {
    const int l = w & (T - 1);
    const int g = w >> lgT;
    const int r = g & ((1 << lgstride) - 1);
    const int base = ((g >> lgstride) << rem) + r;
    uint32_t v[1 << A];
#pragma unroll
    for (int j = 0; j < (1 << A); ++j) v[j] = s[pad16(base + (j << lgstride)) * T + l];
    radix_dif<A>(v);
    if (lgstride > 0) {
#pragma unroll
      for (int j = 1; j < (1 << A); ++j) v[j] = fp_mul(v[j], tw_small(tw_hi, rem, (uint32_t)r * (__brev(j) >> (32 - A))));
    }
#pragma unroll
    for (int j = 0; j < (1 << A); ++j) s[pad16(base + (j << lgstride)) * T + l] = v[j];
  }
}
template <int A, int SKIP>
__device__ __forceinline__ void step_dit(uint32_t* s, int m, int done, int lgT, const uint32_t* __restrict__ tw_hi) {
  const int T = 1 << lgT;
  const int items = (1 << (m - A)) << lgT;
  for (int w = threadIdx.x; w < items; w += blockDim.x) {
    const int l = w & (T - 1);
    const int g = w >> lgT;
    const int r = g & ((1 << done) - 1);
    const int base = ((g >> done) << (done + A)) + r;
    uint32_t v[1 << A];
#pragma unroll
    for (int j = 0; j < (1 << A); ++j) v[j] = s[pad16(base + (j << done)) * T + l];
    if (done > 0) {
#pragma unroll
      for (int j = 1; j < (1 << A); ++j)
        v[j] = fp_mul(v[j], tw_small(tw_hi, done + A, (uint32_t)r * (__brev(j) >> (32 - A))));
    }
    radix_dit<A, SKIP>(v);
#pragma unroll
    for (int j = 0; j < (1 << A); ++j) s[pad16(base + (j << done)) * T + l] = v[j];
  }
}

// natural in -> bit-reversed out (ROU_REV), no scaling. Caller syncs before (data loaded) ; ends synced.
__device__ void sm_dif(uint32_t* s, int m, int lgT, const uint32_t* __restrict__ tw_hi) {
  int rem = m;
  while (rem > 0) {
    const int a = rem < 4 ? rem : 4;
    switch (a) {
      case 4: step_dif<4>(s, m, rem, lgT, tw_hi); break;
      case 3: step_dif<3>(s, m, rem, lgT, tw_hi); break;
      case 2: step_dif<2>(s, m, rem, lgT, tw_hi); break;
      default: step_dif<1>(s, m, rem, lgT, tw_hi); break;
    }
    rem -= a;
    __syncthreads();
  }
}
// bit-reversed in -> natural out (ROU_FWD); the first `skip` (0 or 2) layers are skipped (replicated input).
__device__ void sm_dit(uint32_t* s, int m, int lgT, int skip, const uint32_t* __restrict__ tw_hi) {
  int done = 0;
  const int tail = m & 3;  // steps from the bottom: 4,4,...,tail  (tail last unless m < 4)
  while (done < m) {
    const int a = (m - done >= 4) ? 4 : (m - done);
    const bool first = done == 0;
    if (first && skip == 2) {
      switch (a) {
        case 4: step_dit<4, 2>(s, m, done, lgT, tw_hi); break;
        case 3: step_dit<3, 2>(s, m, done, lgT, tw_hi); break;
        default: step_dit<2, 2>(s, m, done, lgT, tw_hi); break;
      }
    } else {
      switch (a) {
        case 4: step_dit<4, 0>(s, m, done, lgT, tw_hi); break;
        case 3: step_dit<3, 0>(s, m, done, lgT, tw_hi); break;
        case 2: step_dit<2, 0>(s, m, done, lgT, tw_hi); break;
        default: step_dit<1, 0>(s, m, done, lgT, tw_hi); break;
      }
    }
    done += a;
    __syncthreads();
  }
  (void)tail;
}

struct NttArgs {
  const uint32_t* in;
  uint32_t* out;
  int k;         // log2 of the (output) transform size
  int k1, k2;    // pass split, k = k1 + k2 (k1 == 0: single pass)
  int eb;        // expand bits (forward only): input rows have 2^(k-eb) elements
  int mode;      // inverse epilogue: 0 none, 1 * n^-1, 2 * n^-1 * 3^brev_k(i)
  uint32_t ninv;
  const uint32_t* tw_lo;
  const uint32_t* tw_hi;
  const uint32_t* p3_lo;
  const uint32_t* p3_hi_scaled;
};

// Contiguous pass. grid = (2^k1 tiles, columns).
//  DIR 0: standalone 2^k2 DIF of in[col][tile*2^k2 ..] + scale epilogue.
//  DIR 1: load 2^(k2-eb) inputs, replicate, standalone DIT with `eb` skipped layers; if k1 > 0 the inter-pass twiddle
//         w_{2^k}^(brev_k1(tile) * i) is applied on store.
template <int DIR>
__global__ void __launch_bounds__(256) ntt_contig_kernel(NttArgs a) {
  extern __shared__ uint32_t s[];
  const int m = a.k2;
  const int tile = blockIdx.x;
  const size_t col = blockIdx.y;
  const int n_tile = 1 << m;
  uint32_t* out = a.out + (col << a.k) + ((size_t)tile << m);
  if (DIR == 0) {
    const uint32_t* in = a.in + (col << a.k) + ((size_t)tile << m);
    for (int i = threadIdx.x; i < n_tile; i += blockDim.x) s[pad16(i)] = in[i];
    __syncthreads();
    sm_dif(s, m, 0, a.tw_hi);
    for (int i = threadIdx.x; i < n_tile; i += blockDim.x) {
      uint32_t v = s[pad16(i)];
      if (a.mode == 1) {
        v = fp_mul(v, a.ninv);
      } else if (a.mode == 2) {
        uint32_t p = ((uint32_t)tile << m) + i;
        uint32_t e = __brev(p) >> (32 - a.k);
        uint32_t sc = __ldg(a.p3_hi_scaled + (e >> 12));
        uint32_t lo = e & 4095u;
        if (lo) sc = fp_mul(sc, __ldg(a.p3_lo + lo));
        v = fp_mul(v, sc);
      }
      out[i] = v;
    }
  } else {
    const int n_in = n_tile >> a.eb;
    const uint32_t* in = a.in + (col << (a.k - a.eb)) + ((size_t)tile << (m - a.eb));
    for (int i = threadIdx.x; i < n_in; i += blockDim.x) {
      uint32_t v = in[i];
      for (int r = 0; r < (1 << a.eb); ++r) s[pad16((i << a.eb) + r)] = v;
    }
    __syncthreads();
    sm_dit(s, m, 0, a.eb, a.tw_hi);
    const uint32_t bt = a.k1 ? (__brev((uint32_t)tile) >> (32 - a.k1)) : 0u;
    for (int i = threadIdx.x; i < n_tile; i += blockDim.x) {
      uint32_t v = s[pad16(i)];
      if (a.k1 && bt && i) v = fp_mul(v, tw_big(a.tw_lo, a.tw_hi, a.k, bt * (uint32_t)i));
      out[i] = v;
    }
  }
}

// Strided pass (in place on a.out). grid = (2^k2 / T, columns); tile = 2^k1 rows x T adjacent columns.
//  DIR 0: standalone 2^k1 DIF down the rows, then * w_{2^k}^(L * brev_k1(row)).
//  DIR 1: standalone 2^k1 DIT down the rows (inputs were pre-twiddled by the contiguous pass).
template <int DIR>
__global__ void __launch_bounds__(256) ntt_strided_kernel(NttArgs a, int lgT) {
  extern __shared__ uint32_t s[];
  const int T = 1 << lgT;
  const int m = a.k1;
  const size_t col = blockIdx.y;
  const uint32_t L0 = blockIdx.x << lgT;
  uint32_t* io = a.out + (col << a.k) + L0;
  const int total = (1 << m) << lgT;
  for (int w = threadIdx.x; w < total; w += blockDim.x) {
    const int l = w & (T - 1), h = w >> lgT;
    s[pad16(h) * T + l] = io[((size_t)h << a.k2) + l];
  }
  __syncthreads();
  if (DIR == 0) {
    sm_dif(s, m, lgT, a.tw_hi);
  } else {
    sm_dit(s, m, lgT, 0, a.tw_hi);
  }
  for (int w = threadIdx.x; w < total; w += blockDim.x) {
    const int l = w & (T - 1), h = w >> lgT;
    uint32_t v = s[pad16(h) * T + l];
    if (DIR == 0) {
      const uint32_t L = L0 + l;
      const uint32_t bh = __brev((uint32_t)h) >> (32 - m);
      if (L && bh) v = fp_mul(v, tw_big(a.tw_lo, a.tw_hi, a.k, L * bh));
    }
    io[((size_t)h << a.k2) + l] = v;
  }
}

// ---- bit reversal ------------------------------------------------------------------------------------------
// Index p = (a : t bits | mid | b : t bits) maps to (brev b | brev mid | brev a). A block swaps the 2^t x 2^t tile of
// `mid` with the tile of brev(mid) through shared memory, so both the reads and the writes are 2^t-element runs.
__global__ void __launch_bounds__(256) bit_reverse_kernel(uint32_t* io, int k, int t) {
  __shared__ uint32_t sa[32][33];
  __shared__ uint32_t sb[32][33];
  const int midbits = k - 2 * t;
  const uint32_t mid = blockIdx.x;
  const uint32_t rmid = midbits ? (__brev(mid) >> (32 - midbits)) : 0u;
  if (mid > rmid) return;
  uint32_t* col = io + ((size_t)blockIdx.y << k);
  const int side = 1 << t;
  for (int w = threadIdx.x; w < side * side; w += blockDim.x) {
    const int b = w & (side - 1), a = w >> t;
    sa[a][b] = col[((size_t)a << (k - t)) + ((size_t)mid << t) + b];
    if (mid != rmid) sb[a][b] = col[((size_t)a << (k - t)) + ((size_t)rmid << t) + b];
  }
  __syncthreads();
  for (int w = threadIdx.x; w < side * side; w += blockDim.x) {
    const int x = w & (side - 1), y = w >> t;  // destination: row y (top bits), column x (low bits)
    const int ra = __brev((uint32_t)x) >> (32 - t);  // source a = brev(x)
    const int rb = __brev((uint32_t)y) >> (32 - t);  // source b = brev(y)
    // element (a, mid, b) goes to (brev b, rmid, brev a): destination (y, rmid, x) takes source (ra, mid, rb)
    col[((size_t)y << (k - t)) + ((size_t)rmid << t) + x] = sa[ra][rb];
    if (mid != rmid) col[((size_t)y << (k - t)) + ((size_t)mid << t) + x] = sb[ra][rb];
  }
}
// tiny rows (k < 2): nothing to do. k in [2, 3]: t = 1 works with the tiled kernel (side 2).

// ---- table construction ------------------------------------------------------------------------------------
__global__ void pow_table_kernel(uint32_t* out, uint32_t base, int n) {  // out[j] = base^j
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < n) out[j] = fp_pow(base, (uint64_t)j);
}
__global__ void scale_table_kernel(uint32_t* out, const uint32_t* in, uint32_t s, int n) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < n) out[j] = fp_mul(in[j], s);
}

}  // namespace r0

using namespace r0;

#include "tables/field_tables.h"

void r0_ntt_init_tables(Ctx* c) {
  Tables& t = c->tab;
  for (int d = 0; d < 2; d++) {
    uint32_t w24 = d == 0 ? R0_ROU_REV_MONT[24] : R0_ROU_FWD_MONT[24];
    uint32_t w24_4096 = fp_pow(w24, 4096);
    R0_CUDA(cudaMalloc(&t.tw_lo[d], 4096 * 4));
    R0_CUDA(cudaMalloc(&t.tw_hi[d], 4096 * 4));
    pow_table_kernel<<<16, 256, 0, c->stream>>>(t.tw_lo[d], w24, 4096);
    pow_table_kernel<<<16, 256, 0, c->stream>>>(t.tw_hi[d], w24_4096, 4096);
  }
  R0_CUDA(cudaMalloc(&t.p3_lo, 4096 * 4));
  R0_CUDA(cudaMalloc(&t.p3_hi, 4096 * 4));
  pow_table_kernel<<<16, 256, 0, c->stream>>>(t.p3_lo, FP_THREE, 4096);
  pow_table_kernel<<<16, 256, 0, c->stream>>>(t.p3_hi, fp_pow(FP_THREE, 4096), 4096);
  for (int k = 0; k <= MAX_LG; k++) {
    t.p3_hi_scaled[k] = nullptr;
    t.ninv[k] = fp_inv(fp_encode(1u << k));
  }
  uint32_t w16[2][8];
  for (int j = 0; j < 8; j++) {
    w16[0][j] = fp_pow(R0_ROU_REV_MONT[4], j);
    w16[1][j] = fp_pow(R0_ROU_FWD_MONT[4], j);
  }
  R0_CUDA(cudaMemcpyToSymbolAsync(c_w16, w16, sizeof(w16), 0, cudaMemcpyHostToDevice, c->stream));
  R0_CUDA(cudaStreamSynchronize(c->stream));
  R0_CUDA(cudaGetLastError());
  count_launch(c, 6);
}

void r0_ntt_free_tables(Ctx* c) {
  Tables& t = c->tab;
  for (int d = 0; d < 2; d++) {
    cudaFree(t.tw_lo[d]);
    cudaFree(t.tw_hi[d]);
  }
  cudaFree(t.p3_lo);
  cudaFree(t.p3_hi);
  for (int k = 0; k <= MAX_LG; k++)
    if (t.p3_hi_scaled[k]) cudaFree(t.p3_hi_scaled[k]);
}

static const uint32_t* scaled_p3(Ctx* c, int k) {
  Tables& t = c->tab;
  if (!t.p3_hi_scaled[k]) {
    R0_CUDA(cudaMalloc(&t.p3_hi_scaled[k], 4096 * 4));
    scale_table_kernel<<<16, 256, 0, c->stream>>>(t.p3_hi_scaled[k], t.p3_hi, t.ninv[k], 4096);
    count_launch(c);
  }
  return t.p3_hi_scaled[k];
}

static void split(int k, int& k1, int& k2) {
  if (k <= 12) {
    k1 = 0;
    k2 = k;
  } else {
    k2 = (k + 1) / 2;
    k1 = k - k2;
  }
}
static int strided_lgT(int k1) { return k1 <= 11 ? 4 : 3; }
static size_t contig_smem(int k2) { return ((size_t(1) << k2) + (size_t(1) << k2) / 16 + 1) * 4; }
static size_t strided_smem(int k1, int lgT) { return (((size_t(1) << k1) + (size_t(1) << k1) / 16 + 1) << lgT) * 4; }

template <typename K>
static void set_smem(K kernel, size_t bytes) {
  R0_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
}

// cols_per_launch: how many columns one pass-pair covers before moving on (keeps the two-pass intermediate of a
// group inside L2). 0 = pick automatically from the working-set size.
static size_t auto_group(int k, size_t count, size_t bytes_per_elem_resident) {
  const size_t budget = 48u << 20;  // conservative share of the 126 MB L2
  size_t per_col = (size_t(1) << k) * bytes_per_elem_resident;
  size_t g = budget / per_col;
  if (g < 1) g = 1;
  return g > count ? count : g;
}

// io: count rows of 2^k, natural order in, bit-reversed coefficients out, scaled by 2^-k; zk: also * 3^brev(i)
void r0_ntt_interpolate(Ctx* c, uint32_t* io, size_t count, int k, bool zk, size_t cols_per_launch) {
  PhaseScope ph(c, zk ? "ntt_interpolate_zk" : "ntt_interpolate", 8.0 * (double)count * (double)(size_t(1) << k));
  R0_CHECK(k >= 0 && k <= MAX_LG, "batch_interpolate_ntt: size out of range");
  if (count == 0) return;
  NttArgs a{};
  a.k = k;
  split(k, a.k1, a.k2);
  a.eb = 0;
  a.mode = zk ? 2 : 1;
  a.ninv = c->tab.ninv[k];
  a.tw_lo = c->tab.tw_lo[0];
  a.tw_hi = c->tab.tw_hi[0];
  a.p3_lo = c->tab.p3_lo;
  a.p3_hi_scaled = zk ? scaled_p3(c, k) : nullptr;
  const size_t sm2 = contig_smem(a.k2);
  set_smem(ntt_contig_kernel<0>, sm2);
  size_t group = cols_per_launch ? cols_per_launch : (a.k1 ? auto_group(k, count, 4) : count);
  if (group > 65535) group = 65535;
  int lgT = 0;
  size_t sm1 = 0;
  if (a.k1) {
    lgT = strided_lgT(a.k1);
    sm1 = strided_smem(a.k1, lgT);
    set_smem(ntt_strided_kernel<0>, sm1);
  }
  for (size_t c0 = 0; c0 < count; c0 += group) {
    size_t nc = count - c0 < group ? count - c0 : group;
    a.in = io + (c0 << k);
    a.out = io + (c0 << k);
    if (a.k1) {
      ntt_strided_kernel<0><<<dim3(1u << (a.k2 - lgT), (unsigned)nc), 256, sm1, c->stream>>>(a, lgT);
      count_launch(c);
    }
    ntt_contig_kernel<0><<<dim3(1u << a.k1, (unsigned)nc), 256, sm2, c->stream>>>(a);
    count_launch(c);
  }
  R0_CUDA(cudaGetLastError());
}

// out: count rows of 2^k ; in: count rows of 2^(k-eb), bit-reversed coefficient order; natural-order evaluations out
void r0_ntt_expand_evaluate(Ctx* c, uint32_t* out, const uint32_t* in, size_t count, int k, int eb,
                            size_t cols_per_launch) {
  PhaseScope ph(c, "ntt_expand_evaluate", (eb ? 5.0 : 8.0) * 4.0 * (double)count * (double)(size_t(1) << (k - eb)));
  R0_CHECK(k >= eb && k <= MAX_LG, "batch_expand_into_evaluate_ntt: size out of range");
  R0_CHECK(eb == 0 || eb == 2, "batch_expand_into_evaluate_ntt: expand_bits must be 0 or 2");
  if (count == 0) return;
  NttArgs a{};
  a.k = k;
  split(k, a.k1, a.k2);
  a.eb = eb;
  a.tw_lo = c->tab.tw_lo[1];
  a.tw_hi = c->tab.tw_hi[1];
  const size_t sm2 = contig_smem(a.k2);
  set_smem(ntt_contig_kernel<1>, sm2);
  size_t group = cols_per_launch ? cols_per_launch : (a.k1 ? auto_group(k, count, 4) : count);
  if (group > 65535) group = 65535;
  int lgT = 0;
  size_t sm1 = 0;
  if (a.k1) {
    lgT = strided_lgT(a.k1);
    sm1 = strided_smem(a.k1, lgT);
    set_smem(ntt_strided_kernel<1>, sm1);
  }
  for (size_t c0 = 0; c0 < count; c0 += group) {
    size_t nc = count - c0 < group ? count - c0 : group;
    a.in = in + (c0 << (k - eb));
    a.out = out + (c0 << k);
    ntt_contig_kernel<1><<<dim3(1u << a.k1, (unsigned)nc), 256, sm2, c->stream>>>(a);
    count_launch(c);
    if (a.k1) {
      ntt_strided_kernel<1><<<dim3(1u << (a.k2 - lgT), (unsigned)nc), 256, sm1, c->stream>>>(a, lgT);
      count_launch(c);
    }
  }
  R0_CUDA(cudaGetLastError());
}

void r0_bit_reverse(Ctx* c, uint32_t* io, size_t count, int k) {
  PhaseScope ph(c, "bit_reverse", 8.0 * (double)count * (double)(size_t(1) << k));
  R0_CHECK(k >= 0 && k <= 30, "batch_bit_reverse: size out of range");
  if (count == 0 || k < 2) return;
  int t = k / 2 < 5 ? k / 2 : 5;
  for (size_t c0 = 0; c0 < count; c0 += 65535) {
    size_t nc = count - c0 < 65535 ? count - c0 : 65535;
    bit_reverse_kernel<<<dim3(1u << (k - 2 * t), (unsigned)nc), 256, 0, c->stream>>>(io + (c0 << k), k, t);
    count_launch(c);
  }
  R0_CUDA(cudaGetLastError());
}
