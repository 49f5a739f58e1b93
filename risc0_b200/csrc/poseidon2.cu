// Poseidon2 (BabyBear, t = 24, rate 16, capacity 8, x^7, 4 + 21 + 4 rounds) row hashing and Merkle folding, sm_100a.
//
// Replaces: Hal::hash_rows (risc0/zkp/src/hal/cpu.rs:555-567 -> poseidon2/mod.rs:221-244 unpadded_hash) and
//           Hal::hash_fold (cpu.rs:569-581 -> poseidon2/mod.rs:46-58 hash_pair); reference GPU kernels
//           risc0/sys/kernels/zkp/cuda/supra/poseidon2.cuh:112-163 (one launch + device sync per tree level).
//
// The whole 24-word state of a permutation lives in registers (one row / one node per thread). The kernels are
// INT32-pipe bound (1356 Montgomery products per permutation): HBM traffic is 4*cols + 32 bytes per row for
// hash_rows and 96 bytes per node for the fold, far under the memory roofline.
//  * hash_rows reads the column-major matrix with one coalesced 128 B line per warp per column.
//  * merkle_fold_tree folds up to 9 tree levels per launch: a block owns a 512-leaf subtree, keeps the shrinking
//    levels in shared memory and writes every level to the heap array, so a 2^22-leaf tree takes 3 launches
//    instead of 22 launch+sync pairs.
#include "ctx.h"
#include "tables/poseidon2_tables.h"

namespace r0 {

__constant__ uint32_t c_rc_full[8 * 24];
__constant__ uint32_t c_rc_partial[21];
__constant__ uint32_t c_diag[24];
__constant__ uint32_t c_diag_n[24];   // M_INT_DIAG_HZN in NORMAL form: x~ * d mod P keeps x~'s Montgomery form
__constant__ uint32_t c_diag_q[24];   // floor(d * 2^32 / P): Shoup's precomputed quotient for multiplying by the constant d
// ---- tensor-core partial rounds (hash_rows variant 6, see p2_hash_rows_tc_kernel) -------------------------------------
__constant__ uint32_t c_Dm[20];      // D_k = sum_{i>=1} d_i^k, Montgomery form
__constant__ uint32_t c_d21_n[24];   // d_i^21 in normal form, and its Shoup quotient
__constant__ uint32_t c_d21_q[24];
// B operands of the two constant matrices as mma.m16n8k32 fragments: [matrix][(ob * 7 + pos) * 3 + kt][lane] -> (b0, b1)
__device__ uint2 g_p2_bfrag[2][63 * 32];
__constant__ uint32_t c_one;  // = 1, opaque to the compiler: a * c_one + b is an IMAD, i.e. an add on the fma pipe
__constant__ uint32_t c_zero;  // = 0, opaque to the compiler: a + b + c_zero is a genuine three-input IADD3 (alu pipe)

// An addition pinned to the fma pipe: x * 1 + y with a 1 the compiler cannot see through is an IMAD. Which groups of
// linear-layer additions take this form is the MODE mask below; it was chosen by measurement.
__device__ __forceinline__ uint32_t fp_add_fma(uint32_t a, uint32_t b) {
  uint32_t r = a * c_one + b;
  return umin32(r, r - P);
}

// ptxas balances the two integer pipes on its own: it rewrites plain two-input adds as IMAD.IADD (fma pipe) or
// LEA / IADD3 / VIADD (alu pipe) whichever way they are written in C, with a machine model in which IMAD.WIDE / IMAD.HI
// hold the fma pipe 4 cycles and the alu pipe 2 (measured: profiles/r2_int_pipe_rates.log). A three-input add cannot
// become an IMAD, so this form pins an addition to the alu pipe (modes 128 - 372: never faster than ptxas' own split).
__device__ __forceinline__ uint32_t fp_add_alu(uint32_t a, uint32_t b) {
  uint32_t r;   // inline PTX so that the front end cannot share (a + c_zero) between sums and leave two-input adds behind
  asm("{ .reg .u32 t; add.u32 t, %1, %2; add.u32 %0, t, %3; }" : "=r"(r) : "r"(a), "r"(b), "r"(c_zero));
  return umin32(r, r - P);
}

// Signed Montgomery product: for |a|, |b| < P the result is in (-P, P) and congruent to a*b/2^32 - exact because the
// low words of t and m*P cancel. Inside the x^7 chain the intermediate powers never meet an addition, so they can stay
// in this non-canonical signed form and only x^7 pays the "conditional add P" (one VIADDMNMX instead of four on the
// alu pipe, which is the pipe that bounds the permutation).
template <bool ALU = false>
__device__ __forceinline__ int32_t mul_signed(int32_t a, int32_t b) {
  const int64_t t = (int64_t)a * b;
  const int32_t m = (int32_t)((uint32_t)t * MONT_PINV);
  if (ALU) {   // three inputs: stays an IADD3
    int32_t r;
    asm("{ .reg .s32 t; sub.s32 t, %1, %2; add.s32 %0, t, %3; }" : "=r"(r) : "r"((int32_t)(t >> 32)), "r"(__mulhi(m, (int32_t)P)), "r"((int32_t)c_zero));
    return r;
  }
  return (int32_t)(t >> 32) - __mulhi(m, (int32_t)P);
}
template <bool ALU = false>
__device__ __forceinline__ uint32_t sbox7(uint32_t x) {
  const int32_t x1 = (int32_t)x;
  const int32_t x2 = mul_signed<ALU>(x1, x1);
  const int32_t x4 = mul_signed<ALU>(x2, x2);
  const int32_t x6 = mul_signed<ALU>(x4, x2);
  const uint32_t r = (uint32_t)mul_signed<ALU>(x6, x1);
  return umin32(r, r + P);
}
// MODE selects, per group of linear-layer additions, which integer pipe they issue on (bit set = alu pipe as IADD3,
// bit clear = fma pipe as IMAD x*1+y), and whether the round-constant addition is folded into the S-box input:
//   bit 0: M_ext first stage (t0..t3 incl. doublings)   bit 1: M_ext second stage (t4..t7)
//   bit 2: M_ext column sums                            bit 3: M_ext final "+ column sum"
//   bit 4: partial-round sum tree                       bit 5: partial-round "+ sum"
//   bit 7: EVERY addition (and the Montgomery subtraction) pinned to the alu pipe as a three-input IADD3
//   bit 8: bits 0-5 choose between the two PINNED forms (set = three-input IADD3, clear = IMAD x*1+y); without it a
//          set bit only writes a plain add, which ptxas re-issues as IMAD.IADD / IADD3 to balance instruction counts
//   bit 9: the S-box products in the additive (unsigned) Montgomery form, see sbox7_additive
//   bit 6: round constants enter the S-box unreduced: x = c + rc - P in (-P, P) is a legal signed operand of the x^7
//          chain, so the modular add (IADD3 + VIADDMNMX) becomes one IADD3
// The best mask is measured (tools/bench_hash.py --mode, profiles/r2_poseidon2_modes.log); R0B200_P2_MODE overrides it.
template <int MODE, int BIT>
__device__ __forceinline__ uint32_t add_m(uint32_t a, uint32_t b) {
  if (MODE & 128) return fp_add_alu(a, b);
  if (MODE & 256) return (MODE & (1 << BIT)) ? fp_add_alu(a, b) : fp_add_fma(a, b);   // both choices pinned
  if (MODE & (1 << BIT)) return fp_add(a, b);
  return fp_add_fma(a, b);
}
// Unsigned Montgomery product in the additive form: t + (lo(t) * -P^-1) * P has a zero low word and its high word is
// a * b / 2^32 + (less than) P. IMAD.WIDE, IMAD, IMAD.WIDE with a 64-bit addend: no subtraction on the alu pipe.
__device__ __forceinline__ uint32_t mul_additive(uint32_t a, uint32_t b) {
  const uint64_t t = (uint64_t)a * b;
  uint32_t lo = (uint32_t)t, hi = (uint32_t)(t >> 32);
  const uint32_t m = lo * MONT_NINV;
  asm("mad.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.u32 %1, %2, %3, %1;" : "+r"(lo), "+r"(hi) : "r"(m), "r"(P));
  return hi;
}
// x^7 with the additive products (MODE bit 9). Bounds, with c = P / 2^32 = 0.469 and x < P: x2 < (c + 1) P = 1.47 P is
// made canonical; x4 = x2^2 < 1.47 P; x6 = x4 x2 < (1.47 c + 1) P = 1.69 P; x7 = x6 x < (1.69 c + 1) P = 1.79 P < 2^32.
__device__ __forceinline__ uint32_t sbox7_additive(uint32_t x) {
  uint32_t x2 = mul_additive(x, x);
  x2 = umin32(x2, x2 - P);
  const uint32_t x4 = mul_additive(x2, x2);
  const uint32_t x6 = mul_additive(x4, x2);
  const uint32_t r = mul_additive(x6, x);
  return umin32(r, r - P);
}

template <int MODE>
__device__ __forceinline__ uint32_t sbox7_rc(uint32_t x, uint32_t rc) {
  if (MODE & 512) return sbox7_additive(fp_add(x, rc));
  if (MODE & 64) {
    int32_t x1 = (int32_t)(x + rc - P);   // in (-P, P)
    if (MODE & 384) asm("{ .reg .u32 t; add.u32 t, %1, %2; add.u32 %0, t, %3; }" : "=r"(x1) : "r"(x), "r"(rc), "r"(c_zero - P));
    constexpr bool A = (MODE & 384) != 0;
    const int32_t x2 = mul_signed<A>(x1, x1);
    const int32_t x4 = mul_signed<A>(x2, x2);
    const int32_t x6 = mul_signed<A>(x4, x2);
    const uint32_t r = (uint32_t)mul_signed<A>(x6, x1);
    return umin32(r, r + P);
  }
  if (MODE & 384) return sbox7<true>(fp_add_alu(x, rc));
  return sbox7(fp_add(x, rc));
}

// M_ext = circ(2*M4, M4, ..., M4) with M4 the 4x4 matrix of poseidon2/mod.rs:139-151
template <int MODE>
__device__ __forceinline__ void m_ext(uint32_t (&c)[24]) {
  uint32_t s0 = 0, s1 = 0, s2 = 0, s3 = 0;
#pragma unroll
  for (int i = 0; i < 6; i++) {
    uint32_t x0 = c[4 * i], x1 = c[4 * i + 1], x2 = c[4 * i + 2], x3 = c[4 * i + 3];
    uint32_t t0 = add_m<MODE, 0>(x0, x1);
    uint32_t t1 = add_m<MODE, 0>(x2, x3);
    uint32_t t2 = add_m<MODE, 0>(add_m<MODE, 0>(x1, x1), t1);
    uint32_t t3 = add_m<MODE, 0>(add_m<MODE, 0>(x3, x3), t0);
    uint32_t t1d = add_m<MODE, 1>(t1, t1), t0d = add_m<MODE, 1>(t0, t0);
    uint32_t t4 = add_m<MODE, 1>(add_m<MODE, 1>(t1d, t1d), t3);
    uint32_t t5 = add_m<MODE, 1>(add_m<MODE, 1>(t0d, t0d), t2);
    uint32_t t6 = add_m<MODE, 1>(t3, t5);
    uint32_t t7 = add_m<MODE, 1>(t2, t4);
    c[4 * i] = t6;
    c[4 * i + 1] = t5;
    c[4 * i + 2] = t7;
    c[4 * i + 3] = t4;
    s0 = add_m<MODE, 2>(s0, t6);
    s1 = add_m<MODE, 2>(s1, t5);
    s2 = add_m<MODE, 2>(s2, t7);
    s3 = add_m<MODE, 2>(s3, t4);
  }
#pragma unroll
  for (int i = 0; i < 6; i++) {
    c[4 * i] = add_m<MODE, 3>(c[4 * i], s0);
    c[4 * i + 1] = add_m<MODE, 3>(c[4 * i + 1], s1);
    c[4 * i + 2] = add_m<MODE, 3>(c[4 * i + 2], s2);
    c[4 * i + 3] = add_m<MODE, 3>(c[4 * i + 3], s3);
  }
}

template <int MODE>
__device__ __forceinline__ void full_round(uint32_t (&c)[24], int r) {
#pragma unroll
  for (int i = 0; i < 24; i++) c[i] = sbox7_rc<MODE>(c[i], c_rc_full[r * 24 + i]);
  m_ext<MODE>(c);
}

template <int MODE>
__device__ __forceinline__ void partial_round(uint32_t (&c)[24], int r) {
  c[0] = sbox7_rc<MODE>(c[0], c_rc_partial[r]);
  // sum of 24 canonical values: pairwise tree keeps the dependency chain short
  uint32_t p[12];
#pragma unroll
  for (int i = 0; i < 12; i++) p[i] = add_m<MODE, 4>(c[2 * i], c[2 * i + 1]);
#pragma unroll
  for (int i = 0; i < 6; i++) p[i] = add_m<MODE, 4>(p[2 * i], p[2 * i + 1]);
  uint32_t sum = add_m<MODE, 4>(add_m<MODE, 4>(add_m<MODE, 4>(p[0], p[1]), add_m<MODE, 4>(p[2], p[3])), add_m<MODE, 4>(p[4], p[5]));
  // sum + diag_i * c_i. diag_i is a constant, so Shoup's method applies: q = hi(c * floor(d 2^32 / P)), r = c*d - q*P in
  // [0, 2P) using only the LOW 32 bits of both products: IMAD.HI + 2 IMAD = 8 fma-pipe cycles instead of the 10 of a
  // Montgomery product (IMAD.WIDE and IMAD.HI issue at quarter rate on sm_100a, profiles/r1_int_pipe_rates.log).
#pragma unroll
  for (int i = 0; i < 24; i++) {
    const uint32_t q = __umulhi(c[i], c_diag_q[i]);
    uint32_t r = c[i] * c_diag_n[i] - q * P;
    r = umin32(r, r - P);
    c[i] = add_m<MODE, 5>(r, sum);
  }
}

template <int MODE>
__device__ __forceinline__ void p2_permute_m(uint32_t (&c)[24]) {
  m_ext<MODE>(c);
#pragma unroll 1
  for (int r = 0; r < 4; r++) full_round<MODE>(c, r);
#pragma unroll 1
  for (int r = 0; r < 21; r++) partial_round<MODE>(c, r);
#pragma unroll 1
  for (int r = 4; r < 8; r++) full_round<MODE>(c, r);
}
#ifndef R0_P2_DEFAULT_MODE
// 563 = additive S-box products (bit 9) + plain adds (ptxas picks the pipe) in both M4 stages, the partial rounds' sum
// tree and "+ sum" (bits 0, 1, 4, 5), the rest as IMAD: 2.26 Gperm/s against 2.15 for round 1's schedule (32), measured
// over 40 masks (profiles/r2_poseidon2_modes*.log)
#define R0_P2_DEFAULT_MODE 563
#endif

// out[row] = sponge over matrix[j*rows + row], j < cols (overwrite mode, zero-filled tail, empty input = one permute)
template <int MODE>
__global__ void __launch_bounds__(256) p2_hash_rows_kernel(uint32_t* __restrict__ out, const uint32_t* __restrict__ matrix,
                                                         size_t rows, uint32_t cols) {
  size_t row = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= rows) return;
  uint32_t c[24];
#pragma unroll
  for (int i = 0; i < 24; i++) c[i] = 0;
  // single permutation call site (keeps the unrolled round code in the binary once); tail block is zero-filled
  uint32_t done = 0;
  do {
    if (done + 16 <= cols) {
#pragma unroll
      for (int i = 0; i < 16; i++) c[i] = matrix[(size_t)(done + i) * rows + row];
    } else {
#pragma unroll
      for (int i = 0; i < 16; i++) c[i] = (done + i < cols) ? matrix[(size_t)(done + i) * rows + row] : 0u;
    }
    p2_permute_m<MODE>(c);
    done += 16;
  } while (done < cols);
  uint4* o = reinterpret_cast<uint4*>(out + row * 8);
  o[0] = make_uint4(c[0], c[1], c[2], c[3]);
  o[1] = make_uint4(c[4], c[5], c[6], c[7]);
}


// ---- experiment variants of hash_rows (R0B200_P2_VARIANT, tools/bench_hash.py): occupancy against per-thread ILP ----
// V = 1: the same kernel under a 32-register cap (8 blocks of 256 = all 64 warp slots)
// V = 2: two rows per thread, the two states carried through every round together (twice the independent work per
//        warp at half the warps)
template <int MODE, int MINB>
__global__ void __launch_bounds__(256, MINB) p2_hash_rows_occ_kernel(uint32_t* __restrict__ out, const uint32_t* __restrict__ matrix,
                                                                   size_t rows, uint32_t cols) {
  size_t row = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= rows) return;
  uint32_t c[24];
#pragma unroll
  for (int i = 0; i < 24; i++) c[i] = 0;
  uint32_t done = 0;
  do {
#pragma unroll
    for (int i = 0; i < 16; i++) c[i] = (done + i < cols) ? matrix[(size_t)(done + i) * rows + row] : 0u;
    p2_permute_m<MODE>(c);
    done += 16;
  } while (done < cols);
  uint4* o = reinterpret_cast<uint4*>(out + row * 8);
  o[0] = make_uint4(c[0], c[1], c[2], c[3]);
  o[1] = make_uint4(c[4], c[5], c[6], c[7]);
}

template <int MODE>
__device__ __forceinline__ void p2_permute2_m(uint32_t (&a)[24], uint32_t (&b)[24]) {
  m_ext<MODE>(a);
  m_ext<MODE>(b);
#pragma unroll 1
  for (int r = 0; r < 4; r++) {
    full_round<MODE>(a, r);
    full_round<MODE>(b, r);
  }
#pragma unroll 1
  for (int r = 0; r < 21; r++) {
    partial_round<MODE>(a, r);
    partial_round<MODE>(b, r);
  }
#pragma unroll 1
  for (int r = 4; r < 8; r++) {
    full_round<MODE>(a, r);
    full_round<MODE>(b, r);
  }
}

template <int MODE, int MINB>
__global__ void __launch_bounds__(128, MINB) p2_hash_rows2_kernel(uint32_t* __restrict__ out, const uint32_t* __restrict__ matrix,
                                                                size_t rows, uint32_t cols) {
  // rows is even for every caller that reaches this kernel (host checks); thread t takes rows t and t + rows / 2
  const size_t half = rows >> 1;
  const size_t r0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r0 >= half) return;
  const size_t r1 = r0 + half;
  uint32_t a[24], b[24];
#pragma unroll
  for (int i = 0; i < 24; i++) a[i] = b[i] = 0;
  uint32_t done = 0;
  do {
#pragma unroll
    for (int i = 0; i < 16; i++) {
      const bool in = done + i < cols;
      a[i] = in ? matrix[(size_t)(done + i) * rows + r0] : 0u;
      b[i] = in ? matrix[(size_t)(done + i) * rows + r1] : 0u;
    }
    p2_permute2_m<MODE>(a, b);
    done += 16;
  } while (done < cols);
  uint4* o = reinterpret_cast<uint4*>(out + r0 * 8);
  o[0] = make_uint4(a[0], a[1], a[2], a[3]);
  o[1] = make_uint4(a[4], a[5], a[6], a[7]);
  o = reinterpret_cast<uint4*>(out + r1 * 8);
  o[0] = make_uint4(b[0], b[1], b[2], b[3]);
  o[1] = make_uint4(b[4], b[5], b[6], b[7]);
}

// ---- hash_rows with the 21 partial rounds' linear algebra on the INT8 tensor cores (R0B200_P2_VARIANT=6) -------------
// The partial rounds only touch element 0 non-linearly. Writing v_i (i >= 1) for the state at their start and s_r for
// the sum a round adds to every element:  x_i after the 21 rounds = d_i^21 v_i + sum_j d_i^(20-j) s_j, and
// s_r = z_r + sum_i d_i^r v_i + sum_{j<r} D_(r-1-j) s_j with z_r the S-box output of round r and D_k = sum_i d_i^k.
// So the 21 x 24 diagonal multiplications become two constant matrix-vector products per permutation,
//     A = M1 v  (21 x 23, before the rounds)     and     B = M2 s  (23 x 21, after them),
// plus a 21-step scalar recurrence. The products run as u8 x u8 -> s32 mma.sync (IMMA.16832) over a whole warp's 32
// permutations: a field word IS four u8 limbs in the K direction of the A fragment (k = 4 i + limb), the B operand holds
// the limbs of (constant * 2^32 mod P) shifted to the seven product positions, each accumulator stays below 2^23, and
// the seven position sums are folded back to one Montgomery word per output exactly
// (sum_p T_p 2^(8p) = H 2^32 + L  ->  H mod P + L 2^-32 mod P). Everything is integer and exact: same digests
// (the whole Poseidon2 test set passes with R0B200_P2_VARIANT=6).
// Measured (2^22 x 64, profiles/r2_poseidon2_tensor.log): 1.92 Gperm/s against 2.15 for the register kernel, so it stays
// opt-in. It removes 5.2 k of the 29.9 k integer pipe-cycles of a permutation, but as built its phases do not overlap
// well enough: removing in turn the 252 IMMAs (+ their B-fragment loads, 64 KB of L1 traffic per warp and
// permutation), the 21-step recurrence and the 48 recombinations brings 8.75 ms down to 6.46 / 6.94 / 7.38 ms, all
// three to 5.05 ms (the full rounds). The reason is in tools/ubench/mma_u8.cu (profiles/r2_mma_u8_rates.log):
// IMMA.16832 issues once per 8.4 cycles per sub-partition (1953 MAC/clk/SM) and does NOT overlap the integer pipes - an
// IMMA and an IMAD.WIDE per iteration take 14.9 cycles, more than the 8.4 + 4.8 they take alone. The 252 IMMAs of a
// permutation therefore cost 2.1 k issue cycles per warp to save 2.6 k cycles of integer work: a wash through
// mma.sync. Only the asynchronous path (tcgen05.mma issued by one thread, accumulators in TMEM, operands staged by
// TMA) could run beside the S-box arithmetic; that is the version worth building next.
__device__ __forceinline__ void imma_u8(int (&c)[4], const uint32_t (&a)[4], const uint2 b) {
  asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b.x), "r"(b.y));
}

// seven position sums (each < 2^23) of one output -> the Montgomery word of sum_i c_i v_i
__device__ __forceinline__ uint32_t tc_recombine(int t0, int t1, int t2, int t3, int t4, int t5, int t6) {
  const uint32_t u0 = (uint32_t)t1 * 256u + (uint32_t)t0;
  const uint32_t u1 = (uint32_t)t3 * 256u + (uint32_t)t2;
  const uint32_t u2 = (uint32_t)t5 * 256u + (uint32_t)t4;
  const uint64_t l64 = (uint64_t)u1 * 65536u + u0;
  const uint64_t h = (uint64_t)(uint32_t)t6 * 65536u + u2 + (l64 >> 32);     // < 2^40
  const uint32_t l = (uint32_t)l64;
  // L * 2^-32 mod P: m = L P^-1, (L - m P) / 2^32 = -hi(m P)
  const uint32_t m = l * MONT_PINV;
  const uint32_t nh = 0u - __umulhi(m, P);
  const uint32_t r0 = umin32(nh, nh + P);
  // H mod P with H = hh 2^32 + hl. Field words and the constants are < P < 0x78000001, so their top limbs are <= 0x78:
  // t6 <= 23 * 0x78^2 < 2^18.4, H < 2^34.5, hh <= 5 and hh * (2^32 mod P) < 5 * 2^28 < 2 P fits 32 bits.
  uint32_t r1 = (uint32_t)(h >> 32) * MONT_ONE;
  r1 = umin32(r1, r1 - P);
  uint32_t r2 = (uint32_t)h;
  r2 = umin32(r2, r2 - P);
  r2 = umin32(r2, r2 - P);
  return fp_add(fp_add(r0, r1), r2);
}

template <int MODE>
__device__ __forceinline__ void tc_matvec(uint32_t* dst /* warp's [32][28] words out */, const uint32_t* src /* same, in */,
                                         const uint2* __restrict__ bfrag, int lane) {
  const int g = lane >> 2, tig = lane & 3;
#pragma unroll 1
  for (int mt = 0; mt < 2; mt++) {
    uint32_t a[3][4];
#pragma unroll
    for (int kt = 0; kt < 3; kt++) {
      const uint32_t* r0 = src + (mt * 16 + g) * 28 + kt * 8 + tig;
      a[kt][0] = r0[0];
      a[kt][1] = r0[8 * 28];
      a[kt][2] = r0[4];
      a[kt][3] = r0[8 * 28 + 4];
    }
#pragma unroll 1
    for (int ob = 0; ob < 3; ob++) {
      int acc[7][4];
#pragma unroll
      for (int p = 0; p < 7; p++) {
#pragma unroll
        for (int q = 0; q < 4; q++) acc[p][q] = 0;
#pragma unroll
        for (int kt = 0; kt < 3; kt++) imma_u8(acc[p], a[kt], __ldg(&bfrag[((ob * 7 + p) * 3 + kt) * 32 + lane]));
      }
      uint32_t res[4];
#pragma unroll
      for (int q = 0; q < 4; q++)
        res[q] = tc_recombine(acc[0][q], acc[1][q], acc[2][q], acc[3][q], acc[4][q], acc[5][q], acc[6][q]);
      // (row g, out 2 tig), (row g, 2 tig + 1), (row g + 8, 2 tig), (row g + 8, 2 tig + 1) of this block of 8 outputs
      uint32_t* w0 = dst + (mt * 16 + g) * 28 + ob * 8 + tig * 2;
      *reinterpret_cast<uint2*>(w0) = make_uint2(res[0], res[1]);
      *reinterpret_cast<uint2*>(w0 + 8 * 28) = make_uint2(res[2], res[3]);
    }
  }
  __syncwarp();
}

__device__ __forceinline__ uint32_t shoup_mul(uint32_t x, uint32_t dn, uint32_t dq) {   // x * d mod P, d constant
  const uint32_t q = __umulhi(x, dq);
  const uint32_t r = x * dn - q * P;
  return umin32(r, r - P);
}

template <int MODE>
__device__ __forceinline__ void p2_permute_tc(uint32_t (&c)[24], uint32_t* bufv, uint32_t* bufx, int lane) {
  m_ext<MODE>(c);
#pragma unroll 1
  for (int r = 0; r < 4; r++) full_round<MODE>(c, r);
  // ---- partial rounds. bufv keeps the state at their start (elements 1..23 are only needed again at the very end, so
  // they do not occupy registers in between); bufx carries A, then s, then M2 s.
  uint4* vrow = reinterpret_cast<uint4*>(bufv + lane * 28);
  uint4* xrow = reinterpret_cast<uint4*>(bufx + lane * 28);
#pragma unroll
  for (int j = 0; j < 6; j++) vrow[j] = make_uint4(c[4 * j], c[4 * j + 1], c[4 * j + 2], c[4 * j + 3]);
  __syncwarp();
  tc_matvec<MODE>(bufx, bufv, g_p2_bfrag[0], lane);          // bufx[lane][r] = A_r, r < 21
  uint32_t s[21];
  uint32_t x0 = c[0];
#pragma unroll
  for (int r = 0; r < 21; r++) {
    const uint32_t z = sbox7_rc<MODE>(x0, c_rc_partial[r]);
    uint32_t sr = fp_add(z, bufx[lane * 28 + r]);
    if (r > 0) {
      uint32_t lo = 0, hi = 0;
#pragma unroll
      for (int j = 0; j < r; j++) {
        asm("mad.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.u32 %1, %2, %3, %1;" : "+r"(lo), "+r"(hi) : "r"(s[j]), "r"(c_Dm[r - 1 - j]));
        if ((j & 1) == 1 || j == r - 1) hi = umin32(hi, hi - P);
      }
      sr = fp_add(sr, mont_reduce(((uint64_t)hi << 32) | lo));
    }
    s[r] = sr;
    x0 = fp_add(sr, shoup_mul(z, c_diag_n[0], c_diag_q[0]));
  }
  __syncwarp();
#pragma unroll
  for (int j = 0; j < 5; j++) xrow[j] = make_uint4(s[4 * j], s[4 * j + 1], s[4 * j + 2], s[4 * j + 3]);
  xrow[5] = make_uint4(s[20], 0u, 0u, 0u);
  __syncwarp();
  tc_matvec<MODE>(bufx, bufx, g_p2_bfrag[1], lane);          // bufx[lane][i] = sum_j d_i^(20-j) s_j, 1 <= i < 24
  c[0] = x0;
#pragma unroll
  for (int j = 0; j < 6; j++) {
    const uint4 o = xrow[j], v = vrow[j];
    const uint32_t ov[4] = {o.x, o.y, o.z, o.w}, vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int q = 0; q < 4; q++) {
      const int i = 4 * j + q;
      if (i > 0) c[i] = fp_add(ov[q], shoup_mul(vv[q], c_d21_n[i], c_d21_q[i]));
    }
  }
  __syncwarp();
#pragma unroll 1
  for (int r = 4; r < 8; r++) full_round<MODE>(c, r);
}

template <int MODE>
__global__ void __launch_bounds__(128) p2_hash_rows_tc_kernel(uint32_t* __restrict__ out, const uint32_t* __restrict__ matrix,
                                                            size_t rows, uint32_t cols) {
  __shared__ __align__(16) uint32_t sbuf[4][2][32 * 28];
  const int lane = threadIdx.x & 31;
  uint32_t* bufv = sbuf[threadIdx.x >> 5][0];
  uint32_t* bufx = sbuf[threadIdx.x >> 5][1];
  const size_t row_raw = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t row = row_raw < rows ? row_raw : rows - 1;      // every lane stays in the warp-wide MMAs
  uint32_t c[24];
#pragma unroll
  for (int i = 0; i < 24; i++) c[i] = 0;
  uint32_t done = 0;
  do {
#pragma unroll
    for (int i = 0; i < 16; i++) c[i] = (done + i < cols) ? matrix[(size_t)(done + i) * rows + row] : 0u;
    p2_permute_tc<MODE>(c, bufv, bufx, lane);
    done += 16;
  } while (done < cols);
  if (row_raw < rows) {
    uint4* o = reinterpret_cast<uint4*>(out + row * 8);
    o[0] = make_uint4(c[0], c[1], c[2], c[3]);
    o[1] = make_uint4(c[4], c[5], c[6], c[7]);
  }
}

__device__ __forceinline__ void load_pair(uint32_t (&c)[24], const uint32_t* __restrict__ in) {
  const uint4* p = reinterpret_cast<const uint4*>(in);
  uint4 a = p[0], b = p[1], d = p[2], e = p[3];
  c[0] = a.x; c[1] = a.y; c[2] = a.z; c[3] = a.w;
  c[4] = b.x; c[5] = b.y; c[6] = b.z; c[7] = b.w;
  c[8] = d.x; c[9] = d.y; c[10] = d.z; c[11] = d.w;
  c[12] = e.x; c[13] = e.y; c[14] = e.z; c[15] = e.w;
#pragma unroll
  for (int i = 16; i < 24; i++) c[i] = 0;
}

// one level: io[out_size + i] = H(io[in_size + 2i] || io[in_size + 2i + 1])
template <int MODE>
__global__ void __launch_bounds__(256) p2_hash_fold_kernel(uint32_t* io, size_t in_size, size_t out_size) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= out_size) return;
  uint32_t c[24];
  load_pair(c, io + (in_size + 2 * i) * 8);
  p2_permute_m<MODE>(c);
  uint4* o = reinterpret_cast<uint4*>(io + (out_size + i) * 8);
  o[0] = make_uint4(c[0], c[1], c[2], c[3]);
  o[1] = make_uint4(c[4], c[5], c[6], c[7]);
}

// generic pair hash with independent pointers (sppark_poseidon2_fold's signature): out[i] = H(in[2i] || in[2i+1])
template <int MODE>
__global__ void __launch_bounds__(256) p2_fold_pairs_kernel(uint32_t* out, const uint32_t* in, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t c[24];
  load_pair(c, in + 16 * i);
  p2_permute_m<MODE>(c);
  uint4* o = reinterpret_cast<uint4*>(out + i * 8);
  o[0] = make_uint4(c[0], c[1], c[2], c[3]);
  o[1] = make_uint4(c[4], c[5], c[6], c[7]);
}

// Several levels per launch. A block takes 2*B consecutive nodes of the level of `in_size` nodes (B = blockDim.x),
// and produces `levels` levels (B, B/2, ... nodes), each written to its heap position nodes[size + index].
template <int MODE>
__global__ void __launch_bounds__(256) p2_fold_tree_kernel(uint32_t* nodes, size_t in_size, int levels) {
  __shared__ uint32_t sh[256 * 8];
  const int B = blockDim.x;
  const int t = threadIdx.x;
  uint32_t c[24];
  size_t out_size = in_size >> 1;
  size_t base = (size_t)blockIdx.x * B;  // first output node of this block at the current level
  int active = B;
  for (int lv = 0; lv < levels; lv++) {
    if (lv > 0) {
      __syncthreads();  // previous readers of sh are done
      if (t < active) {
#pragma unroll
        for (int i = 0; i < 8; i++) sh[t * 8 + i] = c[i];
      }
      __syncthreads();
      active >>= 1;
      out_size >>= 1;
      base >>= 1;
      if (active == 0 || out_size == 0) break;  // uniform across the block
    }
    if (t < active && base + t < out_size) {
      if (lv == 0) {
        load_pair(c, nodes + (in_size + 2 * (base + t)) * 8);
      } else {
#pragma unroll
        for (int i = 0; i < 16; i++) c[i] = sh[2 * t * 8 + i];
#pragma unroll
        for (int i = 16; i < 24; i++) c[i] = 0;
      }
      p2_permute_m<MODE>(c);
      uint4* o = reinterpret_cast<uint4*>(nodes + (out_size + base + t) * 8);
      o[0] = make_uint4(c[0], c[1], c[2], c[3]);
      o[1] = make_uint4(c[4], c[5], c[6], c[7]);
    }
  }
}

}  // namespace r0

using namespace r0;

void r0_poseidon2_init(Ctx* c) {
  R0_CUDA(cudaMemcpyToSymbolAsync(c_rc_full, R0_P2_RC_FULL_MONT, sizeof(R0_P2_RC_FULL_MONT), 0, cudaMemcpyHostToDevice,
                                  c->stream));
  R0_CUDA(cudaMemcpyToSymbolAsync(c_rc_partial, R0_P2_RC_PARTIAL_MONT, sizeof(R0_P2_RC_PARTIAL_MONT), 0,
                                  cudaMemcpyHostToDevice, c->stream));
  R0_CUDA(cudaMemcpyToSymbolAsync(c_diag, R0_P2_DIAG_MONT, sizeof(R0_P2_DIAG_MONT), 0, cudaMemcpyHostToDevice,
                                  c->stream));
  uint32_t dq[24];
  for (int i = 0; i < 24; i++) dq[i] = (uint32_t)(((uint64_t)R0_P2_DIAG[i] << 32) / P);
  R0_CUDA(cudaMemcpyToSymbolAsync(c_diag_n, R0_P2_DIAG, sizeof(R0_P2_DIAG), 0, cudaMemcpyHostToDevice, c->stream));
  R0_CUDA(cudaMemcpyToSymbolAsync(c_diag_q, dq, sizeof(dq), 0, cudaMemcpyHostToDevice, c->stream));
  const uint32_t one = 1, zero = 0;
  R0_CUDA(cudaMemcpyToSymbolAsync(c_one, &one, sizeof(one), 0, cudaMemcpyHostToDevice, c->stream));
  R0_CUDA(cudaMemcpyToSymbolAsync(c_zero, &zero, sizeof(zero), 0, cudaMemcpyHostToDevice, c->stream));
  // tables of the tensor-core partial rounds (p2_hash_rows_tc_kernel)
  {
    auto mulmod = [](uint64_t a, uint64_t b) { return (uint32_t)(a * b % P); };
    auto powmod = [&](uint32_t b, uint32_t e) {
      uint32_t r = 1;
      for (uint32_t k = 0; k < e; k++) r = mulmod(r, b);
      return r;
    };
    const uint32_t R32 = (uint32_t)((uint64_t(1) << 32) % P);
    uint32_t Dm[20], d21n[24], d21q[24];
    for (int k = 0; k < 20; k++) {
      uint64_t acc = 0;
      for (int i = 1; i < 24; i++) acc += powmod(R0_P2_DIAG[i], k);
      Dm[k] = mulmod(acc % P, R32);
    }
    for (int i = 0; i < 24; i++) {
      d21n[i] = powmod(R0_P2_DIAG[i], 21);
      d21q[i] = (uint32_t)(((uint64_t)d21n[i] << 32) / P);
    }
    // C[matrix][output n][input element k], already multiplied by 2^32 mod P
    static uint32_t C[2][24][24];
    memset(C, 0, sizeof(C));
    for (int r = 0; r < 21; r++)
      for (int i = 1; i < 24; i++) C[0][r][i] = mulmod(powmod(R0_P2_DIAG[i], r), R32);
    for (int i = 1; i < 24; i++)
      for (int j = 0; j < 21; j++) C[1][i][j] = mulmod(powmod(R0_P2_DIAG[i], 20 - j), R32);
    static uint2 frag[2][63 * 32];
    for (int mtx = 0; mtx < 2; mtx++)
      for (int ob = 0; ob < 3; ob++)
        for (int pos = 0; pos < 7; pos++)
          for (int kt = 0; kt < 3; kt++)
            for (int lane = 0; lane < 32; lane++) {
              const int g = lane >> 2, tig = lane & 3;
              uint32_t b[2] = {0, 0};
              for (int half = 0; half < 2; half++) {
                const uint32_t cst = C[mtx][ob * 8 + g][kt * 8 + half * 4 + tig];
                for (int l = 0; l < 4; l++) {
                  const int m = pos - l;     // limb of the constant that meets limb l of the value at position pos
                  if (m >= 0 && m < 4) b[half] |= ((cst >> (8 * m)) & 0xffu) << (8 * l);
                }
              }
              frag[mtx][((ob * 7 + pos) * 3 + kt) * 32 + lane] = make_uint2(b[0], b[1]);
            }
    R0_CUDA(cudaMemcpyToSymbolAsync(c_Dm, Dm, sizeof(Dm), 0, cudaMemcpyHostToDevice, c->stream));
    R0_CUDA(cudaMemcpyToSymbolAsync(c_d21_n, d21n, sizeof(d21n), 0, cudaMemcpyHostToDevice, c->stream));
    R0_CUDA(cudaMemcpyToSymbolAsync(c_d21_q, d21q, sizeof(d21q), 0, cudaMemcpyHostToDevice, c->stream));
    R0_CUDA(cudaMemcpyToSymbolAsync(g_p2_bfrag, frag, sizeof(frag), 0, cudaMemcpyHostToDevice, c->stream));
  }
  R0_CUDA(cudaStreamSynchronize(c->stream));
}

static int p2_mode() {
  static const int mode = getenv("R0B200_P2_MODE") ? atoi(getenv("R0B200_P2_MODE")) : R0_P2_DEFAULT_MODE;
  return mode;
}
// the fold kernels are compiled for the default schedule and the two all-alu reference points of the mode experiment
#define P2_FOLD_DISPATCH(CALL)                         \
  switch (p2_mode()) {                                 \
    case 32: { constexpr int M = 32; CALL; } break;    \
    case 128: { constexpr int M = 128; CALL; } break;  \
    default: { constexpr int M = R0_P2_DEFAULT_MODE; CALL; } break; \
  }

void r0_p2_hash_rows(Ctx* c, uint32_t* out, const uint32_t* matrix, size_t rows, size_t cols) {
  PhaseScope ph(c, "hash_rows", 4.0 * (double)rows * (double)cols + 32.0 * (double)rows);
  if (rows == 0) return;
  R0_CHECK(cols <= 0xffffffffull, "hash_rows: too many columns");
  const unsigned grid = (unsigned)((rows + 255) / 256);
  static const int variant = getenv("R0B200_P2_VARIANT") ? atoi(getenv("R0B200_P2_VARIANT")) : 0;
  if (variant && (rows & 1) == 0) {
    constexpr int M = R0_P2_DEFAULT_MODE;
    const unsigned g2 = (unsigned)((rows / 2 + 127) / 128);
    switch (variant) {
      case 1: p2_hash_rows_occ_kernel<M, 8><<<grid, 256, 0, c->stream>>>(out, matrix, rows, (uint32_t)cols); break;
      case 2: p2_hash_rows2_kernel<M, 1><<<g2, 128, 0, c->stream>>>(out, matrix, rows, (uint32_t)cols); break;
      case 3: p2_hash_rows2_kernel<M, 6><<<g2, 128, 0, c->stream>>>(out, matrix, rows, (uint32_t)cols); break;
      case 4: p2_hash_rows2_kernel<M, 8><<<g2, 128, 0, c->stream>>>(out, matrix, rows, (uint32_t)cols); break;
      case 5: p2_hash_rows_occ_kernel<M, 7><<<grid, 256, 0, c->stream>>>(out, matrix, rows, (uint32_t)cols); break;
      case 6: p2_hash_rows_tc_kernel<M><<<(unsigned)((rows + 127) / 128), 128, 0, c->stream>>>(out, matrix, rows, (uint32_t)cols); break;
      default: throw std::invalid_argument("R0B200_P2_VARIANT: unknown variant");
    }
    count_launch(c);
    R0_CUDA(cudaGetLastError());
    return;
  }
#define P2_ROWS(M) case M: p2_hash_rows_kernel<M><<<grid, 256, 0, c->stream>>>(out, matrix, rows, (uint32_t)cols); break;
  switch (p2_mode()) {
    P2_ROWS(0) P2_ROWS(32) P2_ROWS(63) P2_ROWS(64) P2_ROWS(127) P2_ROWS(128) P2_ROWS(192) P2_ROWS(288) P2_ROWS(304)
    P2_ROWS(372) P2_ROWS(512) P2_ROWS(544) P2_ROWS(546) P2_ROWS(563) P2_ROWS(575)
    default: throw std::invalid_argument("R0B200_P2_MODE: mask not compiled in");
  }
#undef P2_ROWS
  count_launch(c);
  R0_CUDA(cudaGetLastError());
}

void r0_p2_hash_fold(Ctx* c, uint32_t* io, size_t in_size, size_t out_size) {
  PhaseScope ph(c, "hash_fold", 96.0 * (double)out_size);
  R0_CHECK(in_size == 2 * out_size, "hash_fold: input_size must be 2 * output_size");
  if (out_size == 0) return;
  P2_FOLD_DISPATCH((p2_hash_fold_kernel<M><<<(unsigned)((out_size + 255) / 256), 256, 0, c->stream>>>(io, in_size, out_size)));
  count_launch(c);
  R0_CUDA(cudaGetLastError());
}

void r0_p2_fold_pairs(Ctx* c, uint32_t* out, const uint32_t* in, size_t n) {
  PhaseScope ph(c, "hash_fold", 96.0 * (double)n);
  if (n == 0) return;
  P2_FOLD_DISPATCH((p2_fold_pairs_kernel<M><<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(out, in, n)));
  count_launch(c);
  R0_CUDA(cudaGetLastError());
}

// All levels below `leaves` (a power of two): nodes[leaves .. 2*leaves) are the leaf digests, fills nodes[1 .. leaves).
void r0_p2_merkle_fold_all(Ctx* c, uint32_t* nodes, size_t leaves) {
  PhaseScope ph(c, "hash_fold", 96.0 * (double)(leaves - 1));
  size_t in_size = leaves;
  // wide levels: one full-width launch per level (every thread does exactly one permutation); the narrow top of the
  // tree (<= 2^13 nodes per level, latency-bound) is folded several levels per launch
  while (in_size > (size_t(1) << 14)) {
    size_t out_size = in_size / 2;
    P2_FOLD_DISPATCH((p2_hash_fold_kernel<M><<<(unsigned)((out_size + 255) / 256), 256, 0, c->stream>>>(nodes, in_size, out_size)));
    count_launch(c);
    in_size = out_size;
  }
  while (in_size > 1) {
    size_t out_size = in_size / 2;
    int B = out_size >= 256 ? 256 : (int)out_size;
    int levels = 1;
    while ((1 << levels) <= B) levels++;  // B = 2^j -> j + 1 levels (B, B/2, ..., 1)
    unsigned blocks = (unsigned)((out_size + B - 1) / B);
    P2_FOLD_DISPATCH((p2_fold_tree_kernel<M><<<blocks, B, 0, c->stream>>>(nodes, in_size, levels)));
    count_launch(c);
    in_size >>= levels;
  }
  R0_CUDA(cudaGetLastError());
}
