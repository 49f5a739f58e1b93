// Batched BabyBear NTT / iNTT / x4 low-degree extension for sm_100a.
//
// Replaces: Hal::batch_interpolate_ntt (risc0/zkp/src/hal/cpu.rs:342-350 -> core/ntt.rs:232-282),
//           Hal::zk_shift (cpu.rs:395-408), Hal::batch_expand_into_evaluate_ntt (cpu.rs:305-340 ->
//           core/ntt.rs:284-343), Hal::batch_bit_reverse (cpu.rs:352-360); reference GPU path: per-column host loops
//           over sppark NTTs (risc0/sys/kernels/zkp/cuda/supra/ntt.cu:34-152).
//
// Design (modelled and checked against the oracle in tools/ntt_model.py; tuning history in profiles/r1_ntt_sweep*.log):
//  * four-step split n = 2^k1 * 2^k2: one "contiguous" pass over 2^k2-element tiles (k2 = 12 whenever n allows: three
//    radix-16 steps) and one "strided" pass over tiles of 2^k1 rows x 8 adjacent columns; n <= 2^12 needs only the
//    contiguous pass. All columns of a batch are processed by one launch per pass (grid.y = columns), not a host loop.
//  * inside a tile the 2^m-point transform runs as radix-16 register steps (4 butterfly layers per shared-memory round
//    trip, sixteenth roots from the constant bank, inter-step twiddles as one contiguous table row per thread fetched
//    with 128-bit loads); shared memory is padded (i + i/16) so every step is bank-conflict free or 2-way at worst.
//  * sums / differences that only feed a twiddle product stay unreduced (lazy_add / lazy_sub).
//  * everything that would be an extra HBM pass in the reference is fused: the 1/n scale and the zk shift 3^brev(i)
//    into the iNTT's last store, the expand-by-4 into the LDE's first load (the two skipped butterfly layers become a
//    template parameter), the inter-pass twiddle into the producing pass (one lookup per thread x 16 per-tile factors
//    staged in shared memory).
//  * algorithmic HBM bytes: iNTT 8 B/element, LDE 20 B/input element; two-pass sizes move 16 / 52 B. The kernels are
//    bound by the integer pipes (ncu: fmaheavy 69-79 %, DRAM <= 37 %), so the intermediate is NOT kept in L2 by
//    splitting the batch into column groups any more - that only cost tail waves (R0_NTT_L2_MB re-enables it).
#include "ctx.h"

// Compile-time tuning knobs (tools/sweep_ntt.py builds and times variants; the defaults are the measured best).
#ifndef R0_NTT_THREADS
#define R0_NTT_THREADS 256
#endif
#ifndef R0_NTT_STRIDED_THREADS
#define R0_NTT_STRIDED_THREADS R0_NTT_THREADS   // strided-pass block size for k1 <= 10
#endif
#ifndef R0_NTT_BIG_STRIDED
#define R0_NTT_BIG_STRIDED 1   // k1 = 11, 12: 8-column tiles (32 B segments) in blocks of 512 / 1024 threads
#endif
#ifndef R0_NTT_ROWTW
#define R0_NTT_ROWTW 1   // inter-step twiddles from per-step row tables (vector loads) instead of 15 scalar gathers
#endif
#ifndef R0_NTT_LAZY
#define R0_NTT_LAZY 1    // operands of a twiddle product stay unreduced (x + y, x - y + P < 2^32): no VIADDMNMX
#endif
#ifndef R0_NTT_LGT_SMALL
#define R0_NTT_LGT_SMALL 3  // log2 of the adjacent columns per strided tile for k1 <= 10
#endif

#ifndef R0_NTT_ADDITIVE
#define R0_NTT_ADDITIVE 1   // twiddle products reduced in the additive Montgomery form (see ntt_mul): -2.6 % on iNTT and LDE (profiles/r2_ntt_sweep5.log)
#endif
#ifndef R0_NTT_SHOUP
#define R0_NTT_SHOUP 0   // table twiddles as (w, floor(w 2^32 / P)) pairs: IMAD.HI + 2 IMAD + VIADDMNMX per product
#endif
#ifndef R0_NTT_IMAD
#define R0_NTT_IMAD 0    // bit 0: butterfly sums, bit 1: butterfly differences issued as IMAD (fma pipe) instead of IADD3
#endif
#ifndef R0_NTT_IADD3
#define R0_NTT_IADD3 0   // bit 0: butterfly sums, bit 1: differences, bit 2: the Montgomery reduction's subtraction are
#endif                   // written as three-input adds (+ an opaque 0) so ptxas cannot turn them into IMAD.IADD

namespace r0 {

__constant__ uint32_t c_w16[2][8];  // w_16^j, j < 8, Montgomery; [0] = ROU_REV[4], [1] = ROU_FWD[4]
__constant__ uint2 c_w16s[2][8];    // the same twiddles as Shoup pairs (plain w, floor(w * 2^32 / P))

// Shoup product by a table constant: x any u32 (the data stay in Montgomery form: (xR) * w = (xw)R), w < P plain,
// wp = floor(w * 2^32 / P). q = hi(x * wp) is floor(x * w / P) or one less, so x * w - q * P lies in [0, 2P) and only the
// low words are needed: IMAD.HI + 2 IMAD = 8 fma-pipe cycles against the 10 of a Montgomery product (IMAD.WIDE, IMAD,
// IMAD.HI), and one alu instruction instead of two.
__device__ __forceinline__ uint32_t shoup_mul(uint32_t x, uint32_t w, uint32_t wp) {
  const uint32_t q = __umulhi(x, wp);
  const uint32_t r = x * w - q * P;
  return umin32(r, r - P);
}
__device__ __forceinline__ uint32_t ntt_mul(uint32_t a, uint32_t b);
template <int DIR>
__device__ __forceinline__ uint32_t tw16_mul(uint32_t x, int idx) {
#if R0_NTT_SHOUP
  const uint2 w = c_w16s[DIR][idx];
  return shoup_mul(x, w.x, w.y);
#else
  return ntt_mul(x, c_w16[DIR][idx]);
#endif
}

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
  return l ? ntt_mul(h, __ldg(lo + l)) : h;
}

// ---- in-register radix-2^A transforms with compile-time twiddles --------------------------------------------
// Every loop bound is a template constant so the butterflies unroll completely and v[] stays in registers.
// Butterfly add / sub. The alu pipe (IADD3, VIADDMNMX) is the busier one in these kernels (ncu: profiles/), so the plain
// 32-bit add can be issued on the fma pipe as x * 1 + y with a 1 the compiler cannot see through (R0_NTT_IMAD).
// Conversely ptxas itself rewrites two-input adds as IMAD.IADD (it models IMAD.WIDE / IMAD.HI cheaper than they are on
// sm_100a), which loads the fmaheavy pipe - the one ncu shows saturated; a three-input add with an opaque zero stays an
// IADD3 on the alu pipe (R0_NTT_IADD3).
__constant__ uint32_t c_ntt_one[2];  // {1, 0xffffffff}
__constant__ uint32_t c_ntt_zero;    // 0 (never written)
__device__ __forceinline__ uint32_t bf_add(uint32_t x, uint32_t y) {
  const uint32_t r = (R0_NTT_IMAD & 1) ? x * c_ntt_one[0] + y : (R0_NTT_IADD3 & 1) ? x + y + c_ntt_zero : x + y;
  return umin32(r, r - P);
}
__device__ __forceinline__ uint32_t bf_sub(uint32_t x, uint32_t y) {
  const uint32_t r = (R0_NTT_IMAD & 2) ? y * c_ntt_one[1] + x : (R0_NTT_IADD3 & 2) ? x - y + c_ntt_zero : x - y;
  return umin32(r, r + P);
}
// Montgomery product as fp_mul (fp.cuh), with the same control over the subtraction's pipe
__device__ __forceinline__ uint32_t ntt_mul(uint32_t a, uint32_t b) {
  const uint64_t t = (uint64_t)a * b;
#if R0_NTT_ADDITIVE
  // additive reduction: t + (lo(t) * -P^-1) * P has a zero low word; with ONE canonical factor t < 2^32 P, so the sum
  // stays below 2^33 P < 2^64 and its high word is the product in [0, 2P). The carry-chained pair becomes one IMAD.WIDE
  // with a 64-bit addend: IMAD.WIDE, IMAD, IMAD.WIDE, VIADDMNMX - one alu-pipe instruction less than the subtractive form.
  uint32_t lo = (uint32_t)t, hi = (uint32_t)(t >> 32);
  const uint32_t mm = lo * MONT_NINV;
  asm("mad.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.u32 %1, %2, %3, %1;" : "+r"(lo), "+r"(hi) : "r"(mm), "r"(P));
  return umin32(hi, hi - P);
#endif
  const uint32_t m = (uint32_t)t * MONT_PINV;
  const uint32_t h = __umulhi(m, P);
  const uint32_t r = (R0_NTT_IADD3 & 4) ? (uint32_t)(t >> 32) - h + c_ntt_zero : (uint32_t)(t >> 32) - h;
  return umin32(r, r + P);
}

// Lazy operands: fp_mul only needs a * b < 2^32 * P, i.e. ONE canonical factor; the other may be any u32. With x, y
// canonical, x + y < 2P and x - y + P in (0, 2P) both fit 32 bits, so a sum or difference that is only ever consumed
// by a twiddle product skips its conditional subtract (the half-rate VIADDMNMX).
__device__ __forceinline__ uint32_t lazy_add(uint32_t x, uint32_t y) { return R0_NTT_LAZY ? x + y : fp_add(x, y); }
__device__ __forceinline__ uint32_t lazy_sub(uint32_t x, uint32_t y) { return R0_NTT_LAZY ? x - y + P : fp_sub(x, y); }

// LAZY_OUT (last layer only): 0 = canonical outputs, 1 = outputs j >= 1 lazy (the caller multiplies them by a twiddle),
// 2 = all outputs lazy (the caller scales every element).
template <int A, int S, int LAZY_OUT>
__device__ __forceinline__ void dif_layer(uint32_t (&v)[1 << A]) {
  constexpr int half = 1 << (S - 1);
#pragma unroll
  for (int b = 0; b < (1 << A); b += 2 * half) {
#pragma unroll
    for (int i = 0; i < half; ++i) {
      const uint32_t x = v[b + i], y = v[b + i + half];
      const bool lazy_top = (S == 1) && (LAZY_OUT == 2 || (LAZY_OUT == 1 && b + i != 0));
      v[b + i] = lazy_top ? lazy_add(x, y) : bf_add(x, y);
      if (i == 0) {
        v[b + i + half] = (S == 1 && LAZY_OUT != 0) ? lazy_sub(x, y) : bf_sub(x, y);
      } else {
        v[b + i + half] = tw16_mul<0>(lazy_sub(x, y), i << (4 - S));
      }
    }
  }
}
template <int A, int LAZY_OUT = 0>
__device__ __forceinline__ void radix_dif(uint32_t (&v)[1 << A]) {  // natural in, bit-reversed out, ROU_REV
  if constexpr (A >= 4) dif_layer<A, 4, LAZY_OUT>(v);
  if constexpr (A >= 3) dif_layer<A, 3, LAZY_OUT>(v);
  if constexpr (A >= 2) dif_layer<A, 2, LAZY_OUT>(v);
  if constexpr (A >= 1) dif_layer<A, 1, LAZY_OUT>(v);
}
template <int A, int S>
__device__ __forceinline__ void dit_layer(uint32_t (&v)[1 << A]) {
  constexpr int half = 1 << (S - 1);
#pragma unroll
  for (int b = 0; b < (1 << A); b += 2 * half) {
#pragma unroll
    for (int i = 0; i < half; ++i) {
      const uint32_t x = v[b + i];
      const uint32_t y = (i == 0) ? v[b + i + half] : tw16_mul<1>(v[b + i + half], i << (4 - S));
      // an output is only an operand of the next layer's twiddle product when it is a lower input there (bit S set)
      // with a non-trivial twiddle (low S bits non-zero): leave exactly those unreduced
      const int p0 = b + i, p1 = b + i + half;
      const bool more = S < A;
      const bool lz0 = more && ((p0 >> S) & 1) && (p0 & ((1 << S) - 1));
      const bool lz1 = more && ((p1 >> S) & 1) && (p1 & ((1 << S) - 1));
      v[p0] = lz0 ? lazy_add(x, y) : bf_add(x, y);
      v[p1] = lz1 ? lazy_sub(x, y) : bf_sub(x, y);
    }
  }
}
template <int A, int SKIP>
__device__ __forceinline__ void radix_dit(uint32_t (&v)[1 << A]) {  // bit-reversed in, natural out, ROU_FWD
  if constexpr (A >= 1 && SKIP < 1) dit_layer<A, 1>(v);
  if constexpr (A >= 2 && SKIP < 2) dit_layer<A, 2>(v);
  if constexpr (A >= 3 && SKIP < 3) dit_layer<A, 3>(v);
  if constexpr (A >= 4 && SKIP < 4) dit_layer<A, 4>(v);
}

// ---- register-resident radix steps ---------------------------------------------------------------------------
// A tile is a 2^M-point transform over the row index h (x T adjacent columns l for the strided pass). It is computed
// as a sequence of radix-2^A steps (A = 4 wherever possible); in each step a thread holds the 2^A elements
// h = base + (j << S) in registers. Between two steps the elements are exchanged through shared memory, laid out as
// s[pad(h) * T + l] with pad(h) = h + (h >> 4); because S is always 0, 4 or 8 the padded index is linear in j
// (offset j * PSTRIDE), so every shared-memory address is base + immediate. The first step reads global memory
// directly and the last one writes it directly (fused with the expand-by-4 load, the inter-pass twiddle, the 1/n scale
// and the zk shift), so a 2^12-point tile makes only two trips through shared memory.
__constant__ uint32_t c_p3top[MAX_LG + 1][16];  // [k][j] = 3^(brev4(j) << (k-4)) for k >= 4 (zk-shift epilogue)

__device__ __forceinline__ constexpr int brev_small(int j, int bits) {
  int r = 0;
  for (int b = 0; b < bits; ++b) r |= ((j >> b) & 1) << (bits - 1 - b);
  return r;
}
__host__ __device__ constexpr int pstride(int s) { return s == 0 ? 1 : (s == 4 ? 17 : 272); }
__device__ __forceinline__ int padh(int h) { return h + (h >> 4); }
__host__ __device__ constexpr int pad_size(int m) { return (1 << m) + ((1 << m) >> 4) + 1; }

// Inter-step twiddle rows. For a step at stride 2^D (D = 4 or 8) of radix 2^A the thread with residue r needs
// w_{2^(D+A)}^(r * brev_A(j)), j < 2^A: stored as one contiguous row per r so a thread fetches them with 128-bit loads
// and consecutive threads (consecutive r) read consecutive rows. Layout: [D = 4: A = 1..4][D = 8: A = 1..4].
__host__ __device__ constexpr int row_off(int D, int A) {
  int off = 0;
  for (int d = 4; d <= 8; d += 4)
    for (int a = 1; a <= 4; ++a) {
      if (d == D && a == A) return off;
      off += (1 << d) << a;
    }
  return off;
}
constexpr int ROW_TABLE_WORDS = row_off(12, 0);

template <int A, int D>
__device__ __forceinline__ void load_tw_row(uint32_t (&t)[1 << A], const uint32_t* __restrict__ rows, int r) {
  const uint32_t* p = rows + row_off(D, A) + (r << A);
  if constexpr (A == 1) {
    const uint2 q = __ldg(reinterpret_cast<const uint2*>(p));
    t[0] = q.x; t[1] = q.y;
  } else {
#pragma unroll
    for (int q4 = 0; q4 < (1 << A) / 4; ++q4) {
      const uint4 q = __ldg(reinterpret_cast<const uint4*>(p) + q4);
      t[4 * q4] = q.x; t[4 * q4 + 1] = q.y; t[4 * q4 + 2] = q.z; t[4 * q4 + 3] = q.w;
    }
  }
}

// DIT: elements v[j] = x[base + (j << DONE)], r = base & (2^DONE - 1)
template <int A, int DONE, int SKIP>
__device__ __forceinline__ void dit_regs(uint32_t (&v)[1 << A], int r, const uint32_t* __restrict__ tw_hi,
                                         const uint32_t* __restrict__ tw_row) {
  if (DONE > 0) {
#if R0_NTT_ROWTW
    uint32_t t[1 << A];
    load_tw_row<A, DONE>(t, tw_row, r);
#if R0_NTT_SHOUP
    uint32_t tp[1 << A];
    load_tw_row<A, DONE>(tp, tw_row + ROW_TABLE_WORDS, r);
#pragma unroll
    for (int j = 1; j < (1 << A); ++j) v[j] = shoup_mul(v[j], t[j], tp[j]);
#else
#pragma unroll
    for (int j = 1; j < (1 << A); ++j) v[j] = ntt_mul(v[j], t[j]);
#endif
#else
#pragma unroll
    for (int j = 1; j < (1 << A); ++j) v[j] = ntt_mul(v[j], __ldg(tw_hi + ((r * brev_small(j, A)) << (12 - DONE - A))));
#endif
  }
  radix_dit<A, SKIP>(v);
}
// DIF: elements v[j] = x[base + (j << LGS)], r = base & (2^LGS - 1). SCALED: the caller multiplies every output.
template <int A, int LGS, bool SCALED = false>
__device__ __forceinline__ void dif_regs(uint32_t (&v)[1 << A], int r, const uint32_t* __restrict__ tw_hi,
                                         const uint32_t* __restrict__ tw_row) {
  if (LGS > 0) {
#if R0_NTT_ROWTW
    uint32_t t[1 << A];
    load_tw_row<A, LGS>(t, tw_row, r);   // issued before the butterflies: the latency hides behind them
#if R0_NTT_SHOUP
    uint32_t tp[1 << A];
    load_tw_row<A, LGS>(tp, tw_row + ROW_TABLE_WORDS, r);
    radix_dif<A, 1>(v);
#pragma unroll
    for (int j = 1; j < (1 << A); ++j) v[j] = shoup_mul(v[j], t[j], tp[j]);
#else
    radix_dif<A, 1>(v);
#pragma unroll
    for (int j = 1; j < (1 << A); ++j) v[j] = ntt_mul(v[j], t[j]);
#endif
#else
    radix_dif<A, 1>(v);
#pragma unroll
    for (int j = 1; j < (1 << A); ++j) v[j] = ntt_mul(v[j], __ldg(tw_hi + ((r * brev_small(j, A)) << (12 - LGS - A))));
#endif
  } else {
    radix_dif<A, SCALED ? 2 : 0>(v);
  }
}

struct NttArgs {
  const uint32_t* in;
  uint32_t* out;
  int k;         // log2 of the (output) transform size
  int k1, k2;    // pass split, k = k1 + k2 (k1 == 0: single pass)
  int eb;        // expand bits (forward only): input rows have 2^(k-eb) elements
  int mode;      // inverse epilogue: 0 none, 1 * n^-1, 2 * n^-1 * 3^brev_k(i)
  uint32_t ninv;
  size_t tiles_total;  // contiguous pass: number of 2^k2 tiles over all columns of this launch
  const uint32_t* tw_lo;
  const uint32_t* tw_hi;
  const uint32_t* tw_row;  // inter-step twiddle rows of this direction (row_off)
  const uint32_t* p3_lo;
  const uint32_t* p3_hi_scaled;
};

constexpr int NTT_THREADS = R0_NTT_THREADS;

// Geometry of one step inside a block that owns TILES tiles of 2^M rows x T columns.
//   item index w -> (tile t, group g, column l); the thread's elements are rows base + (j << S), j < 2^A.
template <int M, int A, int S, int LGT>
struct StepGeom {
  static constexpr int GROUPS = 1 << (M - A);          // per tile
  static constexpr int ITEMS_PER_TILE = GROUPS << LGT;
  __device__ static __forceinline__ void decode(int w, int& t, int& l, int& r, int& base) {
    l = w & ((1 << LGT) - 1);
    const int gg = w >> LGT;
    t = gg >> (M - A);
    const int g = gg & (GROUPS - 1);
    r = g & ((1 << S) - 1);
    base = ((g >> S) << (S + A)) + r;
  }
};

// ---- contiguous pass -----------------------------------------------------------------------------------------
// Block = BLOCK_ELEMS consecutive elements = TILES tiles of 2^M (M = k2). DIR 1 (forward): load 2^(M-EB) inputs per
// tile (replicated 2^EB times), DIT with EB skipped layers, inter-pass twiddle on store when k1 > 0.
// DIR 0 (inverse): DIF, then the scale / zk-shift epilogue.
template <int M>
struct ContigCfg {
  static constexpr int BLOCK_LG = M > 12 ? M : 12;
  static constexpr int TILES = 1 << (BLOCK_LG - M);
  static constexpr int TILE_PAD = pad_size(M);
  // last step of the forward transform (the one that applies the inter-pass twiddle): radix 2^LAST_A at stride 2^LAST_D
  static constexpr int LAST_A = (M % 4) ? (M % 4) : 4;
  static constexpr int LAST_D = M - LAST_A;
  static constexpr bool STEP_TW = M > 4;  // at least one barrier separates fwd_contig_prep from the last step
  static constexpr size_t SMEM = ((size_t)TILES * TILE_PAD + (STEP_TW ? (TILES << LAST_A) : 0)) * 4;
};

// s_tw[t][j] = w_{2^k}^(bt_t * (j << LAST_D)) for the tiles of this block
template <int M>
__device__ __forceinline__ void fwd_contig_prep(const NttArgs& a, uint32_t* s, size_t tile0) {
  using C = ContigCfg<M>;
  if constexpr (C::STEP_TW) {
    if (a.k1) {
      uint32_t* sx = s + C::TILES * C::TILE_PAD;
      for (int w = threadIdx.x; w < (C::TILES << C::LAST_A); w += R0_NTT_THREADS) {
        const size_t tile = tile0 + (w >> C::LAST_A);
        const uint32_t j = w & ((1 << C::LAST_A) - 1);
        const uint32_t bt = __brev((uint32_t)(tile & ((size_t(1) << a.k1) - 1))) >> (32 - a.k1);
        sx[w] = tw_big(a.tw_lo, a.tw_hi, a.k, bt * (j << C::LAST_D));
      }
    }
  }
}

template <int M, int A, int DONE, int SKIP, int EB, bool FROM_GLOBAL, bool TO_GLOBAL>
__device__ __forceinline__ void fwd_contig_step(const NttArgs& a, uint32_t* s, size_t tile0) {
  using G = StepGeom<M, A, DONE, 0>;
  constexpr int TILES = ContigCfg<M>::TILES;
  constexpr int ITEMS = G::ITEMS_PER_TILE * TILES;
  constexpr int PS = pstride(DONE);
#pragma unroll 1
  for (int w = threadIdx.x; w < ITEMS; w += NTT_THREADS) {
    int t, l, r, base;
    G::decode(w, t, l, r, base);
    const size_t tile = tile0 + t;
    if (tile >= a.tiles_total) continue;
    uint32_t v[1 << A];
    uint32_t* sp = s + t * ContigCfg<M>::TILE_PAD + padh(base);
    if (FROM_GLOBAL) {
      // DONE == 0: the thread's elements are 2^A consecutive outputs = 2^(A-EB) consecutive inputs
      const uint32_t* in = a.in + (tile << (M - EB)) + (base >> EB);
      if (A - EB == 2) {
        const uint4 q = *reinterpret_cast<const uint4*>(in);
        const uint32_t x[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int j = 0; j < (1 << A); ++j) v[j] = x[j >> EB];
      } else if (A - EB == 4) {
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) {
          const uint4 q = reinterpret_cast<const uint4*>(in)[q4];
          v[4 * q4] = q.x; v[4 * q4 + 1] = q.y; v[4 * q4 + 2] = q.z; v[4 * q4 + 3] = q.w;
        }
      } else {
#pragma unroll
        for (int j = 0; j < (1 << A); ++j) v[j] = in[j >> EB];
      }
    } else {
#pragma unroll
      for (int j = 0; j < (1 << A); ++j) v[j] = sp[j * PS];
    }
    dit_regs<A, DONE, SKIP>(v, r, a.tw_hi, a.tw_row);
    if (TO_GLOBAL) {
      uint32_t* out = a.out + (tile << M) + base;
      if (a.k1) {
        // inter-pass twiddle w_{2^k}^(bt * i), bt = brev_k1(column tile), i = base + (j << DONE) the index inside the tile
        const uint32_t bt = __brev((uint32_t)(tile & ((size_t(1) << a.k1) - 1))) >> (32 - a.k1);
        if (bt) {
          if constexpr (ContigCfg<M>::STEP_TW) {
            // = w^(bt * base) (one two-level lookup per thread) * w^(bt * (j << DONE)) (2^A values per tile, in shared
            // memory since fwd_contig_prep): same number of products as a lookup per element, 2 loads instead of 2^(A+1)
            const uint32_t* sx = s + TILES * ContigCfg<M>::TILE_PAD + (t << A);
            const uint32_t g0 = tw_big(a.tw_lo, a.tw_hi, a.k, bt * (uint32_t)base);
#pragma unroll
            for (int j = 0; j < (1 << A); ++j) v[j] = ntt_mul(v[j], j ? ntt_mul(g0, sx[j]) : g0);
          } else {
#pragma unroll
            for (int j = 0; j < (1 << A); ++j) {
              const uint32_t i = (uint32_t)(base + (j << DONE));
              if (i) v[j] = ntt_mul(v[j], tw_big(a.tw_lo, a.tw_hi, a.k, bt * i));
            }
          }
        }
      }
      if (DONE == 0 && A == 4) {
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4)
          reinterpret_cast<uint4*>(out)[q4] = make_uint4(v[4 * q4], v[4 * q4 + 1], v[4 * q4 + 2], v[4 * q4 + 3]);
      } else {
#pragma unroll
        for (int j = 0; j < (1 << A); ++j) out[j << DONE] = v[j];
      }
    } else {
#pragma unroll
      for (int j = 0; j < (1 << A); ++j) sp[j * PS] = v[j];
    }
  }
}

template <int M, int EB>
__global__ void __launch_bounds__(NTT_THREADS) ntt_fwd_contig_kernel(NttArgs a) {
  extern __shared__ uint32_t s[];
  constexpr int Q = M / 4, REM = M % 4;
  constexpr int TILES = ContigCfg<M>::TILES;
  const size_t tile0 = (size_t)blockIdx.x * TILES;
  fwd_contig_prep<M>(a, s, tile0);
  if constexpr (Q == 0) {
    fwd_contig_step<M, REM, 0, EB, EB, true, true>(a, s, tile0);   // M in 1..3 (M == 0 is handled by the host)
  } else {
    fwd_contig_step<M, 4, 0, EB, EB, true, (Q == 1 && REM == 0)>(a, s, tile0);
    if constexpr (Q >= 2) {
      __syncthreads();
      fwd_contig_step<M, 4, 4, 0, EB, false, (Q == 2 && REM == 0)>(a, s, tile0);
    }
    if constexpr (Q >= 3) {
      __syncthreads();
      fwd_contig_step<M, 4, 8, 0, EB, false, (Q == 3 && REM == 0)>(a, s, tile0);
    }
    if constexpr (REM != 0) {
      __syncthreads();
      fwd_contig_step<M, REM, 4 * Q, 0, EB, false, true>(a, s, tile0);
    }
  }
}

template <int M, int A, int LGS, bool FROM_GLOBAL, bool TO_GLOBAL>
__device__ __forceinline__ void inv_contig_step(const NttArgs& a, uint32_t* s, size_t tile0) {
  using G = StepGeom<M, A, LGS, 0>;
  constexpr int TILES = ContigCfg<M>::TILES;
  constexpr int ITEMS = G::ITEMS_PER_TILE * TILES;
  constexpr int PS = pstride(LGS);
#pragma unroll 1
  for (int w = threadIdx.x; w < ITEMS; w += NTT_THREADS) {
    int t, l, r, base;
    G::decode(w, t, l, r, base);
    const size_t tile = tile0 + t;
    if (tile >= a.tiles_total) continue;
    uint32_t v[1 << A];
    uint32_t* sp = s + t * ContigCfg<M>::TILE_PAD + padh(base);
    if (FROM_GLOBAL) {
      const uint32_t* in = a.in + (tile << M) + base;
      if (LGS == 0 && A == 4) {
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) {
          const uint4 q = reinterpret_cast<const uint4*>(in)[q4];
          v[4 * q4] = q.x; v[4 * q4 + 1] = q.y; v[4 * q4 + 2] = q.z; v[4 * q4 + 3] = q.w;
        }
      } else {
#pragma unroll
        for (int j = 0; j < (1 << A); ++j) v[j] = in[j << LGS];
      }
    } else {
#pragma unroll
      for (int j = 0; j < (1 << A); ++j) v[j] = sp[j * PS];
    }
    dif_regs<A, LGS, TO_GLOBAL>(v, r, a.tw_hi, a.tw_row);
    if (TO_GLOBAL) {
      // LGS == 0: the thread owns 2^A consecutive outputs p = tile * 2^M + base + j
      uint32_t* out = a.out + (tile << M) + base;
      if (a.mode == 1) {
#pragma unroll
        for (int j = 0; j < (1 << A); ++j) v[j] = ntt_mul(v[j], a.ninv);
      } else if (a.mode == 2) {
        // n^-1 * 3^brev_k(p); p = p0 + j with p0 a multiple of 2^A, so brev_k(p) = brev_k(p0) + (brev_A(j) << (k - A))
        const uint32_t p0 = (uint32_t)(((tile << M) + base) & ((size_t(1) << a.k) - 1));
        const uint32_t e0 = a.k ? (__brev(p0) >> (32 - a.k)) : 0u;
        uint32_t sc = __ldg(a.p3_hi_scaled + (e0 >> 12));
        const uint32_t lo = e0 & 4095u;
        if (lo) sc = ntt_mul(sc, __ldg(a.p3_lo + lo));
#pragma unroll
        for (int j = 0; j < (1 << A); ++j) {
          // c_p3top[k][x] = 3^(brev4(x) << (k-4)); brev_A(j) << (k-A) == brev4(j << (4-A)) << (k-4)
          const uint32_t f = (j == 0) ? sc : ntt_mul(sc, c_p3top[a.k][j << (4 - A)]);
          v[j] = ntt_mul(v[j], f);
        }
      } else {
#pragma unroll
        for (int j = 0; j < (1 << A); ++j) v[j] = umin32(v[j], v[j] - P);  // the last layer left its outputs unreduced
      }
      if (A == 4) {
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4)
          reinterpret_cast<uint4*>(out)[q4] = make_uint4(v[4 * q4], v[4 * q4 + 1], v[4 * q4 + 2], v[4 * q4 + 3]);
      } else {
#pragma unroll
        for (int j = 0; j < (1 << A); ++j) out[j] = v[j];
      }
    } else {
#pragma unroll
      for (int j = 0; j < (1 << A); ++j) sp[j * PS] = v[j];
    }
  }
}

template <int M>
__global__ void __launch_bounds__(NTT_THREADS) ntt_inv_contig_kernel(NttArgs a) {
  extern __shared__ uint32_t s[];
  constexpr int Q = M / 4, REM = M % 4;
  constexpr int TILES = ContigCfg<M>::TILES;
  const size_t tile0 = (size_t)blockIdx.x * TILES;
  if constexpr (Q == 0) {
    inv_contig_step<M, REM, 0, true, true>(a, s, tile0);
  } else {
    // remainder first (top layers), then radix-16 steps at strides 2^(4(Q-1)), ..., 16, 1
    if constexpr (REM != 0) {
      inv_contig_step<M, REM, 4 * Q, true, false>(a, s, tile0);
      __syncthreads();
    }
    if constexpr (Q >= 3) {
      inv_contig_step<M, 4, 8, (REM == 0), false>(a, s, tile0);
      __syncthreads();
    }
    if constexpr (Q >= 2) {
      inv_contig_step<M, 4, 4, (REM == 0 && Q == 2), false>(a, s, tile0);
      __syncthreads();
    }
    inv_contig_step<M, 4, 0, (REM == 0 && Q == 1), true>(a, s, tile0);
  }
}

// ---- strided pass (in place on a.out) ---------------------------------------------------------------------------
// Block = 2^M rows (M = k1, row stride 2^k2) x T = 2^LGT adjacent columns. grid = (2^k2 / T, columns).
//  DIR 0: standalone DIF down the rows, then * w_{2^k}^(L * brev_k1(row)) (L = column index inside the row).
//  DIR 1: standalone DIT down the rows (inputs were pre-twiddled by the contiguous pass).
template <int M>
struct StridedCfg {
  // 2^LGT adjacent columns per tile. A tile of 2^12 rows x 8 columns is 139 KB: one block per SM, so the block is made
  // large enough (1024 threads, 2 items per thread and step) to keep the SM busy on its own.
  static constexpr int LGT = M <= 10 ? R0_NTT_LGT_SMALL : (R0_NTT_BIG_STRIDED ? 3 : (M == 11 ? 3 : 2));
  static constexpr int THREADS =
      M <= 10 || !R0_NTT_BIG_STRIDED ? R0_NTT_STRIDED_THREADS : (M == 11 ? 512 : 1024);
  static constexpr int LAST_A = M >= 4 ? 4 : M;   // radix of the inverse transform's last step (stride 1)
  static constexpr bool STEP_TW = M > 4;          // a barrier separates strided_prep from the last step
  static constexpr size_t SMEM = (((size_t)pad_size(M) << LGT) + (STEP_TW ? (size_t(1) << (LGT + LAST_A)) : 0)) * 4;
};

// inverse direction: s_tw[l][j] = w_{2^k}^(L_l * (brev_A(j) << (M - A))), the per-column part of the inter-pass twiddle
template <int M>
__device__ __forceinline__ void strided_prep(const NttArgs& a, uint32_t* s, uint32_t L0) {
  using C = StridedCfg<M>;
  if constexpr (C::STEP_TW) {
    uint32_t* sx = s + (pad_size(M) << C::LGT);
    for (int w = threadIdx.x; w < (1 << (C::LGT + C::LAST_A)); w += C::THREADS) {
      const uint32_t L = L0 + (w >> C::LAST_A);
      const uint32_t j = w & ((1 << C::LAST_A) - 1);
      const uint32_t bj = __brev(j) >> (32 - C::LAST_A);
      sx[w] = tw_big(a.tw_lo, a.tw_hi, a.k, L * (bj << (M - C::LAST_A)));
    }
  }
}

template <int M, int A, int S, int DIR, bool FROM_GLOBAL, bool TO_GLOBAL>
__device__ __forceinline__ void strided_step(const NttArgs& a, uint32_t* s, uint32_t* io, uint32_t L0) {
  constexpr int LGT = StridedCfg<M>::LGT;
  constexpr int T = 1 << LGT;
  using G = StepGeom<M, A, S, LGT>;
  constexpr int ITEMS = G::ITEMS_PER_TILE;
  constexpr int PS = pstride(S) * T;
#pragma unroll 1
  for (int w = threadIdx.x; w < ITEMS; w += StridedCfg<M>::THREADS) {
    int t, l, r, base;
    G::decode(w, t, l, r, base);
    uint32_t v[1 << A];
    uint32_t* sp = s + padh(base) * T + l;
    uint32_t* gp = io + ((size_t)base << a.k2) + l;
    if (FROM_GLOBAL) {
      // the inverse transform may read its input from another buffer (a.in) and leave it untouched; io - a.out is the
      // element offset of this block's tile
      const uint32_t* rp = (DIR == 0) ? a.in + (gp - a.out) : gp;
#pragma unroll
      for (int j = 0; j < (1 << A); ++j) v[j] = rp[(size_t)(j << S) << a.k2];
    } else {
#pragma unroll
      for (int j = 0; j < (1 << A); ++j) v[j] = sp[j * PS];
    }
    if (DIR == 0) {
      dif_regs<A, S>(v, r, a.tw_hi, a.tw_row);
    } else {
      dit_regs<A, S, 0>(v, r, a.tw_hi, a.tw_row);
    }
    if (TO_GLOBAL) {
      if (DIR == 0) {
        // * w_{2^k}^(L * brev_M(row)); S == 0 here, so row = base + j and brev_M(row) = brev_M(base) + (brev_A(j) << (M - A))
        const uint32_t L = L0 + l;
        if (L) {
          if constexpr (StridedCfg<M>::STEP_TW) {
            const uint32_t* sx = s + (pad_size(M) << LGT) + (l << A);
            const uint32_t g0 = tw_big(a.tw_lo, a.tw_hi, a.k, L * (__brev((uint32_t)base) >> (32 - M)));
#pragma unroll
            for (int j = 0; j < (1 << A); ++j) v[j] = ntt_mul(v[j], j ? ntt_mul(g0, sx[j]) : g0);
          } else {
#pragma unroll
            for (int j = 0; j < (1 << A); ++j) {
              const uint32_t bh = __brev((uint32_t)(base + (j << S))) >> (32 - M);
              if (bh) v[j] = ntt_mul(v[j], tw_big(a.tw_lo, a.tw_hi, a.k, L * bh));
            }
          }
        }
      }
#pragma unroll
      for (int j = 0; j < (1 << A); ++j) gp[(size_t)(j << S) << a.k2] = v[j];
    } else {
#pragma unroll
      for (int j = 0; j < (1 << A); ++j) sp[j * PS] = v[j];
    }
  }
}

template <int M, int DIR>
__global__ void __launch_bounds__(StridedCfg<M>::THREADS) ntt_strided_kernel(NttArgs a) {
  extern __shared__ uint32_t s[];
  constexpr int Q = M / 4, REM = M % 4;
  constexpr int LGT = StridedCfg<M>::LGT;
  const uint32_t L0 = blockIdx.x << LGT;
  uint32_t* io = a.out + ((size_t)blockIdx.y << a.k) + L0;
  if constexpr (Q == 0) {
    strided_step<M, REM, 0, DIR, true, true>(a, s, io, L0);
  } else if constexpr (DIR == 1) {
    strided_step<M, 4, 0, 1, true, (Q == 1 && REM == 0)>(a, s, io, L0);
    if constexpr (Q >= 2) {
      __syncthreads();
      strided_step<M, 4, 4, 1, false, (Q == 2 && REM == 0)>(a, s, io, L0);
    }
    if constexpr (Q >= 3) {
      __syncthreads();
      strided_step<M, 4, 8, 1, false, (Q == 3 && REM == 0)>(a, s, io, L0);
    }
    if constexpr (REM != 0) {
      __syncthreads();
      strided_step<M, REM, 4 * Q, 1, false, true>(a, s, io, L0);
    }
  } else {
    strided_prep<M>(a, s, L0);
    if constexpr (REM != 0) {
      strided_step<M, REM, 4 * Q, 0, true, false>(a, s, io, L0);
      __syncthreads();
    }
    if constexpr (Q >= 3) {
      strided_step<M, 4, 8, 0, (REM == 0), false>(a, s, io, L0);
      __syncthreads();
    }
    if constexpr (Q >= 2) {
      strided_step<M, 4, 4, 0, (REM == 0 && Q == 2), false>(a, s, io, L0);
      __syncthreads();
    }
    strided_step<M, 4, 0, 0, (REM == 0 && Q == 1), true>(a, s, io, L0);
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
// (R0_NTT_SHOUP: plain values, followed by a second plane of floor(w * 2^32 / P))
// rows[row_off(D, A) + (r << A) + j] = hi[(r * brev_A(j)) << (12 - D - A)] = w_{2^(D+A)}^(r * brev_A(j))
__global__ void row_table_kernel(uint32_t* rows, const uint32_t* hi) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ROW_TABLE_WORDS) return;
  for (int D = 4; D <= 8; D += 4)
    for (int A = 1; A <= 4; ++A) {
      const int off = row_off(D, A), size = (1 << D) << A;
      if (i >= off && i < off + size) {
        const int r = (i - off) >> A, j = (i - off) & ((1 << A) - 1);
        const uint32_t w = hi[(r * (int)(__brev((uint32_t)j) >> (32 - A))) << (12 - D - A)];
#if R0_NTT_SHOUP
        const uint32_t plain = fp_decode(w);
        rows[i] = plain;
        rows[ROW_TABLE_WORDS + i] = (uint32_t)(((uint64_t)plain << 32) / P);
#else
        rows[i] = w;
#endif
      }
    }
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
    R0_CUDA(cudaMalloc(&t.tw_row[d], 2 * ROW_TABLE_WORDS * 4));
    row_table_kernel<<<(ROW_TABLE_WORDS + 255) / 256, 256, 0, c->stream>>>(t.tw_row[d], t.tw_hi[d]);
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
  static uint2 w16s[2][8];
  for (int d = 0; d < 2; d++)
    for (int j = 0; j < 8; j++) {
      const uint32_t plain = fp_decode(w16[d][j]);
      w16s[d][j] = make_uint2(plain, (uint32_t)(((uint64_t)plain << 32) / P));
    }
  R0_CUDA(cudaMemcpyToSymbolAsync(c_w16s, w16s, sizeof(w16s), 0, cudaMemcpyHostToDevice, c->stream));
  static const uint32_t ones[2] = {1u, 0xffffffffu};
  R0_CUDA(cudaMemcpyToSymbolAsync(c_ntt_one, ones, sizeof(ones), 0, cudaMemcpyHostToDevice, c->stream));
  // c_p3top[k][x] = 3^(brev4(x) * 2^(k-4)) (k >= 4) or 3^brev4(x) (k < 4, only x = j << (4-k) is used): zk-shift factors of the 16
  // consecutive outputs a thread owns in the inverse transform's last step
  static uint32_t p3top[MAX_LG + 1][16];
  for (int k = 0; k <= MAX_LG; k++)
    for (int x = 0; x < 16; x++) {
      uint64_t b4 = brev_bits((uint32_t)x, 4);
      uint64_t e = k >= 4 ? (b4 << (k - 4)) : b4;  // k < 4: x = j << (4-k), so brev4(x) = brev_k(j)
      p3top[k][x] = fp_pow(FP_THREE, e);
    }
  R0_CUDA(cudaMemcpyToSymbolAsync(c_p3top, p3top, sizeof(p3top), 0, cudaMemcpyHostToDevice, c->stream));
  R0_CUDA(cudaStreamSynchronize(c->stream));
  R0_CUDA(cudaGetLastError());
  count_launch(c, 8);
}

void r0_ntt_free_tables(Ctx* c) {
  Tables& t = c->tab;
  for (int d = 0; d < 2; d++) {
    cudaFree(t.tw_lo[d]);
    cudaFree(t.tw_hi[d]);
    cudaFree(t.tw_row[d]);
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

static int env_int(const char* name, int dflt) {
  const char* v = getenv(name);
  return v && *v ? atoi(v) : dflt;
}
// Experiment overrides (read once): R0_NTT_SPLIT = "k:k2,k:k2,..." sets the size of the contiguous pass for transform
// size 2^k, R0_NTT_L2_MB = L2 share of a column group (0: no grouping, one pass pair over all columns).
static const int g_env_l2_mb = env_int("R0_NTT_L2_MB", 0);
static int env_split(int k) {
  static int table[MAX_LG + 1];
  static bool parsed = false;
  if (!parsed) {
    parsed = true;
    const char* v = getenv("R0_NTT_SPLIT");
    while (v && *v) {
      int kk = 0, k2 = 0;
      if (sscanf(v, "%d:%d", &kk, &k2) == 2 && kk >= 0 && kk <= MAX_LG) table[kk] = k2;
      v = strchr(v, ',');
      if (v) ++v;
    }
  }
  return table[k];
}

static void split(int k, int& k1, int& k2) {
  if (k <= 12) {
    k1 = 0;
    k2 = k;
  } else {
    // Largest contiguous pass (three radix-16 steps at k2 = 12), the rest strided: fewest steps and shared-memory
    // round trips. Measured 2^20 x 64 (profiles/r1_ntt_sweep.log): iNTT+zk 0.54 -> 0.40 ms, expand+NTT 1.71 -> 1.52 ms
    // against the balanced split (k + 1) / 2.
    k2 = k - 4 < 12 ? k - 4 : 12;
    const int e = env_split(k);
    if (e >= 6 && e <= 12 && k - e >= 1 && k - e <= 12) k2 = e;
    k1 = k - k2;
  }
}

template <typename K>
static void set_smem(K kernel, size_t bytes) {
  if (bytes > 48 * 1024) R0_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
}

// cols_per_launch: how many columns one pass-pair covers before moving on (keeps the two-pass intermediate of a
// group inside L2). 0 = pick automatically from the working-set size.
static size_t auto_group(int k, size_t count, size_t bytes_per_elem_resident) {
  if (g_env_l2_mb <= 0) return count;
  const size_t budget = (size_t)g_env_l2_mb << 20;  // share of the 126 MB L2 left to the intermediate of one column group
  size_t per_col = (size_t(1) << k) * bytes_per_elem_resident;
  size_t g = budget / per_col;
  if (g < 1) g = 1;
  return g > count ? count : g;
}

#define R0_FOR_M(X) X(1) X(2) X(3) X(4) X(5) X(6) X(7) X(8) X(9) X(10) X(11) X(12)

template <int M>
static void launch_inv_contig(Ctx* c, const NttArgs& a) {
  constexpr int TILES = ContigCfg<M>::TILES;
  set_smem(ntt_inv_contig_kernel<M>, ContigCfg<M>::SMEM);
  const size_t blocks = (a.tiles_total + TILES - 1) / TILES;
  ntt_inv_contig_kernel<M><<<(unsigned)blocks, NTT_THREADS, ContigCfg<M>::SMEM, c->stream>>>(a);
}
template <int M>
static void launch_fwd_contig(Ctx* c, const NttArgs& a) {
  constexpr int TILES = ContigCfg<M>::TILES;
  const size_t blocks = (a.tiles_total + TILES - 1) / TILES;
  if (a.eb == 2) {
    if constexpr (M >= 2) {
      set_smem(ntt_fwd_contig_kernel<M, 2>, ContigCfg<M>::SMEM);
      ntt_fwd_contig_kernel<M, 2><<<(unsigned)blocks, NTT_THREADS, ContigCfg<M>::SMEM, c->stream>>>(a);
    }
  } else {
    set_smem(ntt_fwd_contig_kernel<M, 0>, ContigCfg<M>::SMEM);
    ntt_fwd_contig_kernel<M, 0><<<(unsigned)blocks, NTT_THREADS, ContigCfg<M>::SMEM, c->stream>>>(a);
  }
}
template <int M, int DIR>
static void launch_strided(Ctx* c, const NttArgs& a, size_t ncols) {
  set_smem(ntt_strided_kernel<M, DIR>, StridedCfg<M>::SMEM);
  dim3 grid(1u << (a.k2 - StridedCfg<M>::LGT), (unsigned)ncols);
  ntt_strided_kernel<M, DIR><<<grid, StridedCfg<M>::THREADS, StridedCfg<M>::SMEM, c->stream>>>(a);
}

static void dispatch_inv_contig(Ctx* c, const NttArgs& a) {
  switch (a.k2) {
#define X(M) case M: launch_inv_contig<M>(c, a); break;
    R0_FOR_M(X)
#undef X
    default: throw std::invalid_argument("ntt: unsupported tile size");
  }
  count_launch(c);
}
static void dispatch_fwd_contig(Ctx* c, const NttArgs& a) {
  switch (a.k2) {
#define X(M) case M: launch_fwd_contig<M>(c, a); break;
    R0_FOR_M(X)
#undef X
    default: throw std::invalid_argument("ntt: unsupported tile size");
  }
  count_launch(c);
}
template <int DIR>
static void dispatch_strided(Ctx* c, const NttArgs& a, size_t ncols) {
  switch (a.k1) {
#define X(M) case M: launch_strided<M, DIR>(c, a, ncols); break;
    R0_FOR_M(X)
#undef X
    default: throw std::invalid_argument("ntt: unsupported tile size");
  }
  count_launch(c);
}

// io: count rows of 2^k, natural order in, bit-reversed coefficients out, scaled by 2^-k; zk: also * 3^brev(i).
// src (optional): read the evaluations from there instead of io (out-of-place: saves the caller's copy)
void r0_ntt_interpolate(Ctx* c, uint32_t* io, size_t count, int k, bool zk, size_t cols_per_launch,
                        const uint32_t* src) {
  if (!src) src = io;
  R0_CHECK(k >= 0 && k <= MAX_LG, "batch_interpolate_ntt: size out of range");
  PhaseScope ph(c, zk ? "ntt_interpolate_zk" : "ntt_interpolate", 8.0 * (double)count * (double)(size_t(1) << k));
  if (count == 0) return;
  if (k == 0) {  // a 1-point transform is the identity (n^-1 = 3^0 = 1)
    if (src != io) R0_CUDA(cudaMemcpyAsync(io, src, count * 4, cudaMemcpyDeviceToDevice, c->stream));
    return;
  }
  NttArgs a{};
  a.k = k;
  split(k, a.k1, a.k2);
  a.eb = 0;
  a.mode = zk ? 2 : 1;
  a.ninv = c->tab.ninv[k];
  a.tw_lo = c->tab.tw_lo[0];
  a.tw_hi = c->tab.tw_hi[0];
  a.tw_row = c->tab.tw_row[0];
  a.p3_lo = c->tab.p3_lo;
  a.p3_hi_scaled = zk ? scaled_p3(c, k) : nullptr;
  size_t group = cols_per_launch ? cols_per_launch : (a.k1 ? auto_group(k, count, 4) : count);
  if (group > 65535) group = 65535;
  for (size_t c0 = 0; c0 < count; c0 += group) {
    size_t nc = count - c0 < group ? count - c0 : group;
    a.in = src + (c0 << k);   // read by the first pass only; everything after it works in place on `io`
    a.out = io + (c0 << k);
    a.tiles_total = nc << a.k1;
    if (a.k1) {
      dispatch_strided<0>(c, a, nc);
      a.in = a.out;
    }
    dispatch_inv_contig(c, a);
  }
  R0_CUDA(cudaGetLastError());
}

// out: count rows of 2^k ; in: count rows of 2^(k-eb), bit-reversed coefficient order; natural-order evaluations out
void r0_ntt_expand_evaluate(Ctx* c, uint32_t* out, const uint32_t* in, size_t count, int k, int eb,
                            size_t cols_per_launch) {
  R0_CHECK(k >= eb && k <= MAX_LG, "batch_expand_into_evaluate_ntt: size out of range");
  R0_CHECK(eb == 0 || eb == 2, "batch_expand_into_evaluate_ntt: expand_bits must be 0 or 2");
  PhaseScope ph(c, "ntt_expand_evaluate", (eb ? 5.0 : 8.0) * 4.0 * (double)count * (double)(size_t(1) << (k - eb)));
  if (count == 0) return;
  if (k == 0) {
    R0_CUDA(cudaMemcpyAsync(out, in, count * 4, cudaMemcpyDeviceToDevice, c->stream));
    return;
  }
  NttArgs a{};
  a.k = k;
  split(k, a.k1, a.k2);
  a.eb = eb;
  a.tw_lo = c->tab.tw_lo[1];
  a.tw_hi = c->tab.tw_hi[1];
  a.tw_row = c->tab.tw_row[1];
  size_t group = cols_per_launch ? cols_per_launch : (a.k1 ? auto_group(k, count, 4) : count);
  if (group > 65535) group = 65535;
  for (size_t c0 = 0; c0 < count; c0 += group) {
    size_t nc = count - c0 < group ? count - c0 : group;
    a.in = in + (c0 << (k - eb));
    a.out = out + (c0 << k);
    a.tiles_total = nc << a.k1;
    dispatch_fwd_contig(c, a);
    if (a.k1) dispatch_strided<1>(c, a, nc);
  }
  R0_CUDA(cudaGetLastError());
}

namespace r0 {
bool r0_bit_reverse_tma(Ctx* c, uint32_t* io, size_t count, int k);
}

void r0_bit_reverse(Ctx* c, uint32_t* io, size_t count, int k) {
  PhaseScope ph(c, "bit_reverse", 8.0 * (double)count * (double)(size_t(1) << k));
  R0_CHECK(k >= 0 && k <= 30, "batch_bit_reverse: size out of range");
  if (count == 0 || k < 2) return;
  if (r0_bit_reverse_tma(c, io, count, k)) return;   // k >= 10: TMA tensor loads / stores (bitrev_tma.cu)
  int t = k / 2 < 5 ? k / 2 : 5;
  for (size_t c0 = 0; c0 < count; c0 += 65535) {
    size_t nc = count - c0 < 65535 ? count - c0 : 65535;
    bit_reverse_kernel<<<dim3(1u << (k - 2 * t), (unsigned)nc), 256, 0, c->stream>>>(io + (c0 << k), k, t);
    count_launch(c);
  }
  R0_CUDA(cudaGetLastError());
}
