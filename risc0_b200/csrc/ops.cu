// Element-wise, FRI, coefficient-mixing, evaluation, division and gather kernels of the Hal trait, sm_100a.
//
// Replaces (reference CPU spec -> reference CUDA kernel):
//   eltwise_add/copy/zeroize/sum   cpu.rs:457-522        -> risc0/sys/kernels/zkp/cuda/eltwise.cu:18-86
//   fri_fold                       cpu.rs:524-553        -> kernels.cu:74-93
//   mix_poly_coeffs                cpu.rs:410-455        -> kernels.cu:116-132 (serial RMW loop per thread)
//   batch_evaluate_any             cpu.rs:362-393        -> kernels.cu:46-72   (one block per evaluation)
//   gather_sample / scatter        cpu.rs:583-615        -> kernels.cu:95-114
//   combos_prepare / combos_divide hal/mod.rs:202-257    -> combos.cu:17-51, sppark div_by_x_minus_z
//   prefix_products, eltwise_copy_elem_slice cpu.rs:617-642
// All of these are HBM-bound streaming kernels: grid-stride loops sized to a multiple of the SM count, 32-bit
// coalesced or 128-bit vector accesses, products accumulated lazily in 64 bits (fp.cuh lazy_mac) so an
// FpExt += Fp * FpExt costs four IMAD.WIDE instead of four full Montgomery products.
#include "ctx.h"
#include "launchers.h"

namespace r0 {

static inline unsigned grid_for(const Ctx* c, size_t n, int block = 256, int waves = 8) {
  size_t blocks = (n + block - 1) / block;
  size_t cap = (size_t)c->sm_count * waves;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (unsigned)blocks;
}

#define GRID_STRIDE(i, n) for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < (n); i += (size_t)gridDim.x * blockDim.x)

__global__ void k_add(uint32_t* out, const uint32_t* a, const uint32_t* b, size_t n) {
  GRID_STRIDE(i, n) out[i] = fp_add(a[i], b[i]);
}
__global__ void k_expand_zero(uint32_t* out, const uint32_t* in, size_t n_out, int bits) {
  GRID_STRIDE(w, n_out) out[w] = (w & ((size_t(1) << bits) - 1)) == 0 ? in[w >> bits] : 0u;
}
__global__ void k_mul_factor(uint32_t* io, uint32_t factor, size_t n) { GRID_STRIDE(i, n) io[i] = fp_mul(io[i], factor); }
__global__ void k_copy(uint32_t* out, const uint32_t* in, size_t n) { GRID_STRIDE(i, n) out[i] = in[i]; }
__global__ void k_copy4(uint4* out, const uint4* in, size_t n4) { GRID_STRIDE(i, n4) out[i] = in[i]; }
__global__ void k_zeroize(uint32_t* io, size_t n) {
  GRID_STRIDE(i, n) {
    uint32_t v = io[i];
    if (v == FP_INVALID) io[i] = 0;
  }
}
__global__ void k_fill(uint32_t* io, uint32_t v, size_t n) { GRID_STRIDE(i, n) io[i] = v; }

// in: to_add x count AoS FpExt ; out: 4 SoA planes of count
__global__ void k_sum_ext(uint32_t* out, const uint4* in, size_t count, size_t to_add) {
  GRID_STRIDE(i, count) {
    uint32_t s0 = 0, s1 = 0, s2 = 0, s3 = 0;
    for (size_t t = 0; t < to_add; t++) {
      uint4 v = in[t * count + i];
      s0 = fp_add(s0, v.x);
      s1 = fp_add(s1, v.y);
      s2 = fp_add(s2, v.z);
      s3 = fp_add(s3, v.w);
    }
    out[i] = s0;
    out[count + i] = s1;
    out[2 * count + i] = s2;
    out[3 * count + i] = s3;
  }
}

// out[k*count + idx] = (sum_{i<16} mix^i * in_ext[brev4(i)*count + idx])[k], in_ext component j at in[j*16*count + ..]
__global__ void k_fri_fold(uint32_t* out, const uint32_t* in, size_t count, const FpExt* mix_pows /*16*/) {
  __shared__ FpExt mp[16];
  if (threadIdx.x < 16) mp[threadIdx.x] = mix_pows[threadIdx.x];
  __syncthreads();
  GRID_STRIDE(idx, count) {
    FpExt tot = ext_zero();
#pragma unroll
    for (int i = 0; i < 16; i++) {
      const int rev = ((i & 1) << 3) | ((i & 2) << 1) | ((i & 4) >> 1) | ((i & 8) >> 3);
      const size_t p = (size_t)rev * count + idx;
      FpExt f{{in[p], in[16 * count + p], in[32 * count + p], in[48 * count + p]}};
      tot = ext_add(tot, ext_mul(mp[i], f));
    }
    out[idx] = tot.c[0];
    out[count + idx] = tot.c[1];
    out[2 * count + idx] = tot.c[2];
    out[3 * count + idx] = tot.c[3];
  }
}

// One pass over the input columns: columns are pre-sorted by combo id on the host (`order`), each thread owns one
// coefficient index and accumulates every combo in registers, then does ONE read-modify-write per touched combo.
//   out[combo*count + idx] += sum_{i in combo} mix_pows[i] * in[i*count + idx]
__global__ void k_mix_poly_coeffs(uint4* out, const uint32_t* in, size_t count, const uint32_t* order /*input_size*/,
                                  const uint32_t* seg_begin /*nseg+1*/, const uint32_t* seg_combo /*nseg*/,
                                  uint32_t nseg, const FpExt* mix_pows /*input_size, indexed by column*/) {
  GRID_STRIDE(idx, count) {
    for (uint32_t s = 0; s < nseg; s++) {
      uint64_t a0 = 0, a1 = 0, a2 = 0, a3 = 0;
      for (uint32_t q = seg_begin[s]; q < seg_begin[s + 1]; q++) {
        const uint32_t col = order[q];
        const uint32_t v = in[(size_t)col * count + idx];
        const uint4 m = *reinterpret_cast<const uint4*>(&mix_pows[col]);
        lazy_mac(a0, v, m.x);
        lazy_mac(a1, v, m.y);
        lazy_mac(a2, v, m.z);
        lazy_mac(a3, v, m.w);
      }
      uint4* o = out + (size_t)seg_combo[s] * count + idx;
      uint4 cur = *o;
      cur.x = fp_add(cur.x, lazy_finish(a0));
      cur.y = fp_add(cur.y, lazy_finish(a1));
      cur.z = fp_add(cur.z, lazy_finish(a2));
      cur.w = fp_add(cur.w, lazy_finish(a3));
      *o = cur;
    }
  }
}

// Same, four consecutive coefficient indices per thread: one 128-bit load per column keeps 16 B per thread in flight
// (the scalar kernel is latency-bound at one 4-byte load per thread: 1.7 TB/s measured) and divides the uniform
// order / mix_pows loads by four. Requires count % 4 == 0 and 16-byte aligned columns.
__global__ void __launch_bounds__(256) k_mix_poly_coeffs4(uint4* out, const uint32_t* in, size_t count,
                                                          const uint32_t* order, const uint32_t* seg_begin,
                                                          const uint32_t* seg_combo, uint32_t nseg, const FpExt* mix_pows) {
  const size_t count4 = count >> 2;
  GRID_STRIDE(i4, count4) {
    for (uint32_t s = 0; s < nseg; s++) {
      uint64_t a[4][4] = {};
      const uint32_t qe = seg_begin[s + 1];
#pragma unroll 2
      for (uint32_t q = seg_begin[s]; q < qe; q++) {
        const uint32_t col = order[q];
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(in + (size_t)col * count) + i4);
        const uint4 m = *reinterpret_cast<const uint4*>(&mix_pows[col]);
        const uint32_t vv[4] = {v.x, v.y, v.z, v.w};
        const uint32_t mm[4] = {m.x, m.y, m.z, m.w};
#pragma unroll
        for (int e = 0; e < 4; e++)
#pragma unroll
          for (int k = 0; k < 4; k++) lazy_mac(a[e][k], vv[e], mm[k]);
      }
      uint4* o = out + (size_t)seg_combo[s] * count + (i4 << 2);
#pragma unroll
      for (int e = 0; e < 4; e++) {
        uint4 cur = o[e];
        cur.x = fp_add(cur.x, lazy_finish(a[e][0]));
        cur.y = fp_add(cur.y, lazy_finish(a[e][1]));
        cur.z = fp_add(cur.z, lazy_finish(a[e][2]));
        cur.w = fp_add(cur.w, lazy_finish(a[e][3]));
        o[e] = cur;
      }
    }
  }
}

// ---- batch_evaluate_any -------------------------------------------------------------------------------------
// out[e] = sum_i coeffs[which[e]*n + i] * xs[e]^i. Block (chunk, e): 256 threads x CH coefficients each;
// thread t takes i = base + j*256 + t, so sum = x^base * sum_t x^t * sum_j c[..] * (x^256)^j. The powers x^t and
// (x^256)^j come from per-evaluation tables built by k_eval_tables; partial sums go to `partial[e][chunk]`.
constexpr int EV_T = 256;
constexpr int EV_CH = 128;  // coefficients per thread -> 32768 per block (the per-block epilogue is ~200 instructions)

__global__ void k_eval_tables(FpExt* xt /*[E][256]*/, FpExt* xq /*[E][EV_CH]*/, FpExt* xc /*[E][nchunks]*/,
                              const FpExt* xs, int nchunks) {
  const int e = blockIdx.x;
  const FpExt x = xs[e];
  const int t = threadIdx.x;
  xt[(size_t)e * EV_T + t] = ext_pow(x, t);
  const FpExt x256 = ext_pow(x, EV_T);
  if (t < EV_CH) xq[(size_t)e * EV_CH + t] = ext_pow(x256, t);
  const FpExt xblk = ext_pow(x256, EV_CH);
  for (int c = t; c < nchunks; c += blockDim.x) xc[(size_t)e * nchunks + c] = ext_pow(xblk, c);
}

// R consecutive evaluations e0 .. e0 + R - 1 of the SAME polynomial (the taps of one register at its back-points are
// adjacent in the tap set): the coefficients are read once and multiplied into R x 4 lazy 64-bit accumulators. Two
// products (< 2 P^2) on top of a folded accumulator (< P 2^32) stay under 2^64, so the high word is folded every second
// coefficient only.
// (lo, hi) += a * b as ONE IMAD.WIDE with a 64-bit accumulator operand: the carry-chained pair is the form ptxas folds
// that way. Written in C the zero-extended operand of a predicated load became a 64 x 32-bit product (IMAD.WIDE + IMAD
// + IADD per term), and a chain of mad.wide.u32 is re-associated into IMAD.WIDE + IADD3 / IADD3.X trees (5 alu-pipe
// cycles per term instead of 2).
__device__ __forceinline__ void mad_wide(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b) {
  asm("mad.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.u32 %1, %2, %3, %1;" : "+r"(lo), "+r"(hi) : "r"(a), "r"(b));
}

template <int R, bool FULL>
__global__ void __launch_bounds__(EV_T) k_eval_partial(FpExt* partial, const uint32_t* coeffs, size_t n,
                                                       const uint32_t* which, const uint32_t* run_start,
                                                       const FpExt* xt, const FpExt* xq, const FpExt* xc, int nchunks) {
  __shared__ FpExt q[R][EV_CH];
  __shared__ FpExt red[EV_T];
  const int chunk = blockIdx.x, t = threadIdx.x;
  const uint32_t e0 = run_start[blockIdx.y];
  for (int w = t; w < R * EV_CH; w += EV_T) q[w / EV_CH][w % EV_CH] = xq[(size_t)(e0 + w / EV_CH) * EV_CH + (w % EV_CH)];
  __syncthreads();
  const size_t base = (size_t)chunk * EV_T * EV_CH;
  const uint32_t* poly = coeffs + (size_t)which[e0] * n + base + t;   // FULL: every index of the chunk is < n
  const uint32_t left = FULL ? 0u : (uint32_t)(n > base + t ? (n - base - t + EV_T - 1) / EV_T : 0);   // valid j of this thread
  uint32_t lo[R][4], hi[R][4];
#pragma unroll
  for (int r = 0; r < R; r++)
#pragma unroll
    for (int k = 0; k < 4; k++) lo[r][k] = hi[r][k] = 0;
#pragma unroll 4
  for (int j = 0; j < EV_CH; j += 2) {
    uint32_t c0, c1;
    if (FULL) {
      c0 = poly[(size_t)j * EV_T];
      c1 = poly[(size_t)(j + 1) * EV_T];
    } else {
      c0 = (uint32_t)j < left ? poly[(size_t)j * EV_T] : 0u;
      c1 = (uint32_t)(j + 1) < left ? poly[(size_t)(j + 1) * EV_T] : 0u;
    }
#pragma unroll
    for (int r = 0; r < R; r++) {
#pragma unroll
      for (int k = 0; k < 4; k++) {
        mad_wide(lo[r][k], hi[r][k], c0, q[r][j].c[k]);
        mad_wide(lo[r][k], hi[r][k], c1, q[r][j + 1].c[k]);
        hi[r][k] = umin32(hi[r][k], hi[r][k] - P);
      }
    }
  }
#pragma unroll
  for (int r = 0; r < R; r++) {
    const uint32_t e = e0 + r;
    FpExt s;
#pragma unroll
    for (int k = 0; k < 4; k++) s.c[k] = lazy_finish(((uint64_t)hi[r][k] << 32) | lo[r][k]);
    if (r) __syncthreads();   // red is reused
    red[t] = ext_mul(s, xt[(size_t)e * EV_T + t]);
    __syncthreads();
    for (int w = EV_T / 2; w > 0; w >>= 1) {
      if (t < w) red[t] = ext_add(red[t], red[t + w]);
      __syncthreads();
    }
    if (t == 0) partial[(size_t)e * nchunks + chunk] = ext_mul(red[0], xc[(size_t)e * nchunks + chunk]);
  }
}

__global__ void k_eval_final(FpExt* out, const FpExt* partial, int nchunks) {
  __shared__ FpExt red[256];
  const int e = blockIdx.x, t = threadIdx.x;
  FpExt s = ext_zero();
  for (int c = t; c < nchunks; c += blockDim.x) s = ext_add(s, partial[(size_t)e * nchunks + c]);
  red[t] = s;
  __syncthreads();
  for (int w = 128; w > 0; w >>= 1) {
    if (t < w) red[t] = ext_add(red[t], red[t + w]);
    __syncthreads();
  }
  if (t == 0) out[e] = red[0];
}

// ---- gather / scatter ---------------------------------------------------------------------------------------
__global__ void k_gather(uint32_t* dst, const uint32_t* src, size_t idx, size_t size, size_t stride) {
  GRID_STRIDE(g, size) dst[g] = src[g * stride + idx];
}
// Batched form used by the product driver: every opening of every query in one launch.
// job j: dst[dst_off[j] + g] = src_j[g * stride_j + idx_j] for g < size_j
__global__ void k_gather_batched(uint32_t* dst, const GatherJob* jobs) {
  const GatherJob j = jobs[blockIdx.x];
  for (uint32_t g = threadIdx.x; g < j.size; g += blockDim.x) dst[j.dst_off + g] = j.src[(size_t)g * j.stride + j.idx];
}
// digests: dst[8*j .. 8*j+8) = nodes_j[8*idx_j ..]
__global__ void k_gather_digests(uint32_t* dst, const DigestJob* jobs, size_t njobs) {
  size_t w = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t j = w >> 3;
  if (j < njobs) dst[w] = jobs[j].nodes[jobs[j].idx * 8 + (w & 7)];
}
__global__ void k_scatter(uint32_t* into, const uint32_t* index, size_t cycles, const uint32_t* offsets,
                          const uint32_t* values) {
  GRID_STRIDE(cycle, cycles) {
    for (uint32_t k = index[cycle]; k < index[cycle + 1]; k++) into[offsets[k]] = values[k];
  }
}
__global__ void k_copy_region(uint32_t* into, const uint32_t* from, size_t rows, size_t cols, size_t from_stride,
                              size_t into_offset, size_t into_stride) {
  GRID_STRIDE(w, rows * cols) {
    size_t row = w / cols, col = w % cols;
    into[into_offset + row * into_stride + col] = from[row * from_stride + col];
  }
}

// ---- combos_prepare: tiny, done by one block so that the order of subtraction is irrelevant (exact arithmetic) ----
__global__ void k_combos_prepare(FpExt* combos, const FpExt* coeff_u, uint32_t combo_count, size_t cycles,
                                 const uint32_t* reg_sizes, const uint32_t* reg_combo_ids, const uint32_t* reg_pos,
                                 uint32_t nregs, const FpExt* mix_pows /* nregs + check_size */, uint32_t check_size,
                                 uint32_t check_pos) {
  // each (combo, i) cell receives contributions from many registers: one thread per register would race, so
  // parallelise over destination cells instead: cell = combo * max_size + i
  const uint32_t max_size = 32;
  for (uint32_t cell = threadIdx.x; cell < combo_count * max_size; cell += blockDim.x) {
    const uint32_t combo = cell / max_size, i = cell % max_size;
    FpExt acc = ext_zero();
    bool any = false;
    for (uint32_t r = 0; r < nregs; r++) {
      if (reg_combo_ids[r] == combo && i < reg_sizes[r]) {
        acc = ext_add(acc, ext_mul(mix_pows[r], coeff_u[reg_pos[r] + i]));
        any = true;
      }
    }
    if (any) combos[cycles * combo + i] = ext_sub(combos[cycles * combo + i], acc);
  }
  if (threadIdx.x == 0) {
    FpExt acc = ext_zero();
    for (uint32_t k = 0; k < check_size; k++) acc = ext_add(acc, ext_mul(mix_pows[nregs + k], coeff_u[check_pos + k]));
    combos[cycles * combo_count] = ext_sub(combos[cycles * combo_count], acc);
  }
}

// ---- combos_divide: in-place synthetic division by (x - z) as a blocked suffix scan ------------------------------
// q[i] = sum_{j > i} p[j] z^(j-i-1), remainder = sum_j p[j] z^j. A block owns DIV_SEG = 256 x 8 consecutive
// coefficients (one 128-byte line per thread). Phase A: local Horner per thread, tree-combined to the block's value
// H_b = sum_j p[b*SEG + j] z^j. Phase B (one block): carry into block b = sum_{b' > b} H_b' Z^(b'-b-1), Z = z^SEG,
// as a Kogge-Stone suffix scan. Phase C: recompute the local values, scan them inside the block, add the block carry
// and run the Horner recurrence again, this time writing the quotient. Powers of z come from a small table built on
// the host (DivPowers), so the kernels contain no exponentiation.
constexpr int DIV_E = 8;                 // coefficients per thread
constexpr int DIV_B = 256;               // threads per block
constexpr int DIV_SEG = DIV_E * DIV_B;   // coefficients per block
constexpr int DIV_MAXLG = 14;            // up to 2^14 blocks = 2^25 coefficients

struct DivPowers {
  FpExt z;                  // z
  FpExt zs[8];              // z^(8 * 2^s): in-block scan strides
  FpExt zt[DIV_B];          // z^(8 * (255 - t)): weight of the block carry for thread t
  FpExt Z[DIV_MAXLG + 1];   // (z^SEG)^(2^s): cross-block scan strides
};

__device__ __forceinline__ FpExt ld_ext(const FpExt* p) {
  const uint4 q = *reinterpret_cast<const uint4*>(p);
  return FpExt{{q.x, q.y, q.z, q.w}};
}
__device__ __forceinline__ void st_ext(FpExt* p, const FpExt& v) {
  *reinterpret_cast<uint4*>(p) = make_uint4(v.c[0], v.c[1], v.c[2], v.c[3]);
}

// loads the thread's 8 coefficients (zero beyond n) and returns their local Horner value sum_j c[j] z^j
__device__ __forceinline__ FpExt div_local(const FpExt* p, size_t n, size_t lo, const FpExt& z, FpExt (&c)[DIV_E]) {
#pragma unroll
  for (int j = 0; j < DIV_E; j++) c[j] = (lo + j < n) ? ld_ext(p + lo + j) : ext_zero();
  FpExt acc = c[DIV_E - 1];
#pragma unroll
  for (int j = DIV_E - 2; j >= 0; j--) acc = ext_add(ext_mul(acc, z), c[j]);
  return acc;
}

// Independent divisions (different combos) run side by side: blockIdx.y (blockIdx.x for the one-block carry kernel)
// selects the job; every job has the same length n.
struct DivJob {
  FpExt* p;
  FpExt* remainder;
};

__global__ void __launch_bounds__(DIV_B) k_div_block_totals(FpExt* totals, const DivJob* jobs, size_t n, const DivPowers* pw) {
  __shared__ FpExt sh[DIV_B];
  const int t = threadIdx.x;
  const FpExt* p = jobs[blockIdx.y].p;
  pw += blockIdx.y;
  totals += (size_t)blockIdx.y * gridDim.x;
  const size_t lo = ((size_t)blockIdx.x * DIV_B + t) * DIV_E;
  FpExt c[DIV_E];
  sh[t] = div_local(p, n, lo, pw->z, c);
  __syncthreads();
#pragma unroll
  for (int s = 0; s < 8; s++) {
    const int d = 1 << s;
    if ((t & (2 * d - 1)) == 0) sh[t] = ext_add(sh[t], ext_mul(pw->zs[s], sh[t + d]));
    __syncthreads();
  }
  if (t == 0) totals[blockIdx.x] = sh[0];
}

// carry[b] = sum_{b' > b} totals[b'] Z^(b'-b-1); one block, processes the blocks in windows of 1024 from the top
__global__ void __launch_bounds__(1024) k_div_block_carries(FpExt* carry, const FpExt* totals, size_t nblocks,
                                                            const DivPowers* pw) {
  __shared__ FpExt sh[2][1024];
  __shared__ FpExt above;  // carry entering the current window from the blocks above it
  const int t = threadIdx.x;
  pw += blockIdx.x;
  totals += (size_t)blockIdx.x * nblocks;
  carry += (size_t)blockIdx.x * nblocks;
  if (t == 0) above = ext_zero();
  const size_t nwin = (nblocks + 1023) / 1024;
  for (size_t wi = nwin; wi-- > 0;) {
    const size_t b = wi * 1024 + t;
    // inclusive suffix scan S_t = sum_{t' >= t} H_t' Z^(t'-t) inside the window
    int cur = 0;
    sh[0][t] = b < nblocks ? totals[b] : ext_zero();
    __syncthreads();
    for (int s = 0; s < 10; s++) {
      const int d = 1 << s;
      FpExt v = sh[cur][t];
      if (t + d < 1024) v = ext_add(v, ext_mul(pw->Z[s], sh[cur][t + d]));
      sh[cur ^ 1][t] = v;
      cur ^= 1;
      __syncthreads();
    }
    // carry[b] = S_{t+1} + above * Z^(1023 - t); Z^(1023-t) from the binary expansion of the exponent
    FpExt w = ext_one();
    const int e = 1023 - t;
    for (int s = 0; s < 10; s++)
      if ((e >> s) & 1) w = ext_mul(w, pw->Z[s]);
    FpExt cy = ext_mul(above, w);
    if (t + 1 < 1024) cy = ext_add(cy, sh[cur][t + 1]);
    if (b < nblocks) carry[b] = cy;
    __syncthreads();
    if (t == 0) above = ext_add(sh[cur][0], ext_mul(above, pw->Z[10]));
    __syncthreads();
  }
}

__global__ void __launch_bounds__(DIV_B) k_div_quotient(const DivJob* jobs, size_t n, const FpExt* carry,
                                                        const DivPowers* pw) {
  __shared__ FpExt sh[2][DIV_B];
  const int t = threadIdx.x;
  FpExt* p = jobs[blockIdx.y].p;
  FpExt* remainder = jobs[blockIdx.y].remainder;
  pw += blockIdx.y;
  carry += (size_t)blockIdx.y * gridDim.x;
  const size_t lo = ((size_t)blockIdx.x * DIV_B + t) * DIV_E;
  const FpExt z = pw->z;
  FpExt c[DIV_E];
  int cur = 0;
  sh[0][t] = div_local(p, n, lo, z, c);
  __syncthreads();
#pragma unroll
  for (int s = 0; s < 8; s++) {
    const int d = 1 << s;
    FpExt v = sh[cur][t];
    if (t + d < DIV_B) v = ext_add(v, ext_mul(pw->zs[s], sh[cur][t + d]));
    sh[cur ^ 1][t] = v;
    cur ^= 1;
    __syncthreads();
  }
  // value entering this thread's 8 coefficients from above
  FpExt v = ext_mul(carry[blockIdx.x], pw->zt[t]);
  if (t + 1 < DIV_B) v = ext_add(v, sh[cur][t + 1]);
#pragma unroll
  for (int j = DIV_E - 1; j >= 0; j--) {
    const FpExt next = ext_add(ext_mul(z, v), c[j]);
    if (lo + j < n) st_ext(p + lo + j, v);
    v = next;
  }
  if (blockIdx.x == 0 && t == 0) *remainder = v;
}

__global__ void k_prefix_products(FpExt* io, size_t n) {  // test-only op in the reference; serial, exact
  if (threadIdx.x == 0 && blockIdx.x == 0)
    for (size_t i = 1; i < n; i++) io[i] = ext_mul(io[i], io[i - 1]);
}

__global__ void k_zk_shift(uint32_t* io, size_t total, int bits, const uint32_t* p3_lo, const uint32_t* p3_hi) {
  GRID_STRIDE(i, total) {
    uint32_t pos = (uint32_t)(i & ((size_t(1) << bits) - 1));
    uint32_t e = bits ? (__brev(pos) >> (32 - bits)) : 0u;
    uint32_t s = fp_mul(p3_hi[e >> 12], p3_lo[e & 4095u]);
    io[i] = fp_mul(io[i], s);
  }
}

}  // namespace r0

using namespace r0;

#define LAUNCH_1D(kernel, n, ...)                                            \
  do {                                                                       \
    if ((n) > 0) {                                                           \
      kernel<<<grid_for(c, (n)), 256, 0, c->stream>>>(__VA_ARGS__);          \
      count_launch(c);                                                       \
      R0_CUDA(cudaGetLastError());                                           \
    }                                                                        \
  } while (0)

void r0_eltwise_add(Ctx* c, uint32_t* out, const uint32_t* a, const uint32_t* b, size_t n) { LAUNCH_1D(k_add, n, out, a, b, n); }
void r0_eltwise_copy(Ctx* c, uint32_t* out, const uint32_t* in, size_t n) {
  PhaseScope ph(c, "eltwise_copy", 8.0 * (double)n);
  if ((n % 4) == 0 && ((uintptr_t)out % 16) == 0 && ((uintptr_t)in % 16) == 0) {
    LAUNCH_1D(k_copy4, n / 4, (uint4*)out, (const uint4*)in, n / 4);
  } else {
    LAUNCH_1D(k_copy, n, out, in, n);
  }
}
void r0_eltwise_zeroize(Ctx* c, uint32_t* io, size_t n) { LAUNCH_1D(k_zeroize, n, io, n); }
void r0_fill(Ctx* c, uint32_t* io, uint32_t v, size_t n) {
  PhaseScope ph(c, "fill", 4.0 * (double)n);
  LAUNCH_1D(k_fill, n, io, v, n);
}
void r0_eltwise_sum_ext(Ctx* c, uint32_t* out, const uint32_t* in, size_t count, size_t to_add) {
  PhaseScope ph(c, "eltwise_sum_extelem", 16.0 * (double)count * (double)(to_add + 1));
  LAUNCH_1D(k_sum_ext, count, out, (const uint4*)in, count, to_add);
}
void r0_zk_shift(Ctx* c, uint32_t* io, size_t count, int bits) {
  size_t total = count << bits;
  LAUNCH_1D(k_zk_shift, total, io, total, bits, c->tab.p3_lo, c->tab.p3_hi);
}

// small host->device staging helper: stream-ordered scratch that is freed after the kernels that use it
struct Scratch {
  Ctx* c;
  void* d = nullptr;
  Scratch(Ctx* c_, const void* host, size_t bytes) : c(c_) {
    R0_CUDA(r0_malloc_async(c, &d, bytes ? bytes : 16, c->stream));
    if (bytes) {
      // The caller's buffer may be pinned (the header recommends pinned memory for witnesses), and a copy from pinned
      // memory is truly asynchronous: the caller could free or overwrite it before the GPU reads it. Stage through a
      // library-owned pageable copy instead - cudaMemcpyAsync from pageable memory returns only after the bytes
      // have been taken, so `stage` can be released right away. These arguments are small (indices, combo ids, mix).
      std::vector<unsigned char> stage((const unsigned char*)host, (const unsigned char*)host + bytes);
      R0_CUDA(cudaMemcpyAsync(d, stage.data(), bytes, cudaMemcpyHostToDevice, c->stream));
    }
  }
  Scratch(Ctx* c_, size_t bytes) : c(c_) { R0_CUDA(r0_malloc_async(c, &d, bytes ? bytes : 16, c->stream)); }
  ~Scratch() { cudaFreeAsync(d, c->stream); }
  template <typename T>
  T* as() { return (T*)d; }
};
// Host arguments passed to Scratch may be freed or modified as soon as the constructor returns (see above).

void r0_fri_fold(Ctx* c, uint32_t* out, const uint32_t* in, size_t count, const FpExt& mix) {
  PhaseScope ph(c, "fri_fold", 272.0 * (double)count);
  FpExt pows[16];
  FpExt cur = ext_one();
  for (int i = 0; i < 16; i++) {
    pows[i] = cur;
    cur = ext_mul(cur, mix);
  }
  Scratch mp(c, pows, sizeof(pows));
  LAUNCH_1D(k_fri_fold, count, out, in, count, mp.as<FpExt>());
}

#include <algorithm>
#include <vector>

void r0_mix_poly_coeffs(Ctx* c, uint32_t* out, const FpExt& mix_start, const FpExt& mix, const uint32_t* in,
                        const uint32_t* combos_host, size_t input_size, size_t count) {
  if (input_size == 0 || count == 0) return;
  PhaseScope ph(c, "mix_poly_coeffs", 4.0 * (double)input_size * (double)count);
  std::vector<FpExt> pows(input_size);
  FpExt cur = mix_start;
  for (size_t i = 0; i < input_size; i++) {
    pows[i] = cur;
    cur = ext_mul(cur, mix);
  }
  std::vector<uint32_t> order(input_size);
  for (size_t i = 0; i < input_size; i++) order[i] = (uint32_t)i;
  std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return combos_host[a] < combos_host[b]; });
  std::vector<uint32_t> seg_begin, seg_combo;
  for (size_t q = 0; q < input_size; q++) {
    if (q == 0 || combos_host[order[q]] != combos_host[order[q - 1]]) {
      seg_begin.push_back((uint32_t)q);
      seg_combo.push_back(combos_host[order[q]]);
    }
  }
  seg_begin.push_back((uint32_t)input_size);
  Scratch d_pows(c, pows.data(), pows.size() * sizeof(FpExt));
  Scratch d_order(c, order.data(), order.size() * 4);
  Scratch d_sb(c, seg_begin.data(), seg_begin.size() * 4);
  Scratch d_sc(c, seg_combo.data(), seg_combo.size() * 4);
  if (count % 4 == 0 && (reinterpret_cast<uintptr_t>(in) & 15) == 0) {
    LAUNCH_1D(k_mix_poly_coeffs4, count / 4, (uint4*)out, in, count, d_order.as<uint32_t>(), d_sb.as<uint32_t>(),
              d_sc.as<uint32_t>(), (uint32_t)seg_combo.size(), d_pows.as<FpExt>());
  } else {
    LAUNCH_1D(k_mix_poly_coeffs, count, (uint4*)out, in, count, d_order.as<uint32_t>(), d_sb.as<uint32_t>(),
              d_sc.as<uint32_t>(), (uint32_t)seg_combo.size(), d_pows.as<FpExt>());
  }
}

// distinct_polys: how many different polynomials the evaluations touch (SURVEY 8d counts 4 * P_distinct * n algorithmic
// bytes: evaluations of one register at several back-points re-read its polynomial through L2, not HBM); 0 = unknown
void r0_batch_evaluate_any(Ctx* c, const uint32_t* coeffs, size_t n, const uint32_t* which_dev, const uint32_t* xs_dev,
                           uint32_t* out_dev, size_t eval_count, size_t distinct_polys, const uint32_t* which_host_in) {
  const size_t distinct = distinct_polys && distinct_polys < eval_count ? distinct_polys : eval_count;
  PhaseScope ph(c, "batch_evaluate_any", 4.0 * (double)n * (double)distinct);
  if (eval_count == 0) return;
  const size_t per_block = (size_t)EV_T * EV_CH;
  const int nchunks = (int)((n + per_block - 1) / per_block);
  Scratch xt(c, eval_count * EV_T * sizeof(FpExt));
  Scratch xq(c, eval_count * EV_CH * sizeof(FpExt));
  Scratch xc(c, eval_count * nchunks * sizeof(FpExt));
  Scratch partial(c, eval_count * nchunks * sizeof(FpExt));
  // runs of consecutive evaluations of one polynomial, cut to at most 4, grouped by length: one launch per length
  // (callers that built `which` on the host pass it along; the C-ABI form only has the device copy)
  std::vector<uint32_t> which_copy;
  const uint32_t* which_host = which_host_in;
  if (!which_host) {
    which_copy.resize(eval_count);
    R0_CUDA(cudaMemcpyAsync(which_copy.data(), which_dev, eval_count * 4, cudaMemcpyDeviceToHost, c->stream));
    R0_CUDA(cudaStreamSynchronize(c->stream));
    which_host = which_copy.data();
  }
  const bool full = n % per_block == 0;
  std::vector<uint32_t> runs[4];
  for (size_t e = 0; e < eval_count;) {
    size_t len = 1;
    while (len < 4 && e + len < eval_count && which_host[e + len] == which_host[e]) len++;
    runs[len - 1].push_back((uint32_t)e);
    e += len;
  }
  for (size_t e0 = 0; e0 < eval_count; e0 += 65535) {
    size_t ne = eval_count - e0 < 65535 ? eval_count - e0 : 65535;
    k_eval_tables<<<(unsigned)ne, EV_T, 0, c->stream>>>(xt.as<FpExt>() + e0 * EV_T, xq.as<FpExt>() + e0 * EV_CH,
                                                        xc.as<FpExt>() + e0 * nchunks, (const FpExt*)xs_dev + e0, nchunks);
    count_launch(c);
  }
  for (int len = 1; len <= 4; len++) {
    const std::vector<uint32_t>& rs = runs[len - 1];
    for (size_t r0 = 0; r0 < rs.size(); r0 += 65535) {
      const size_t nr = rs.size() - r0 < 65535 ? rs.size() - r0 : 65535;
      Scratch d_runs(c, rs.data() + r0, nr * 4);
      const dim3 grid(nchunks, (unsigned)nr);
#define EV_LAUNCH(R)                                                                                                     \
  if (full)                                                                                                              \
    k_eval_partial<R, true><<<grid, EV_T, 0, c->stream>>>(partial.as<FpExt>(), coeffs, n, which_dev, d_runs.as<uint32_t>(), \
                                                          xt.as<FpExt>(), xq.as<FpExt>(), xc.as<FpExt>(), nchunks);     \
  else                                                                                                                   \
    k_eval_partial<R, false><<<grid, EV_T, 0, c->stream>>>(partial.as<FpExt>(), coeffs, n, which_dev, d_runs.as<uint32_t>(), \
                                                           xt.as<FpExt>(), xq.as<FpExt>(), xc.as<FpExt>(), nchunks)
      switch (len) {
        case 1: EV_LAUNCH(1); break;
        case 2: EV_LAUNCH(2); break;
        case 3: EV_LAUNCH(3); break;
        default: EV_LAUNCH(4); break;
      }
#undef EV_LAUNCH
      count_launch(c);
    }
  }
  for (size_t e0 = 0; e0 < eval_count; e0 += 65535) {
    size_t ne = eval_count - e0 < 65535 ? eval_count - e0 : 65535;
    k_eval_final<<<(unsigned)ne, 256, 0, c->stream>>>((FpExt*)out_dev + e0, partial.as<FpExt>() + e0 * nchunks, nchunks);
    count_launch(c);
  }
  R0_CUDA(cudaGetLastError());
}

void r0_gather_sample(Ctx* c, uint32_t* dst, const uint32_t* src, size_t idx, size_t size, size_t stride) {
  LAUNCH_1D(k_gather, size, dst, src, idx, size, stride);
}
void r0_scatter(Ctx* c, uint32_t* into, const uint32_t* index_host, size_t index_len, const uint32_t* offsets_host,
                const uint32_t* values_host) {
  if (index_len < 2) return;
  size_t nvals = index_host[index_len - 1];
  Scratch d_index(c, index_host, index_len * 4);
  Scratch d_off(c, offsets_host, nvals * 4);
  Scratch d_val(c, values_host, nvals * 4);
  LAUNCH_1D(k_scatter, index_len - 1, into, d_index.as<uint32_t>(), index_len - 1, d_off.as<uint32_t>(),
            d_val.as<uint32_t>());
}
// device-argument forms (the reference's FFI passes device buffers: risc0_zkp_cuda_scatter, .._eltwise_copy_fp_region)
void r0_scatter_dev(Ctx* c, uint32_t* into, const uint32_t* index_dev, size_t count, const uint32_t* offsets_dev,
                    const uint32_t* values_dev) {
  if (count == 0) return;
  LAUNCH_1D(k_scatter, count, into, index_dev, count, offsets_dev, values_dev);
}
void r0_copy_region_dev(Ctx* c, uint32_t* into, const uint32_t* from_dev, size_t from_rows, size_t from_cols,
                        size_t from_offset, size_t from_stride, size_t into_offset, size_t into_stride) {
  if (from_rows == 0 || from_cols == 0) return;
  LAUNCH_1D(k_copy_region, from_rows * from_cols, into, from_dev + from_offset, from_rows, from_cols, from_stride,
            into_offset, into_stride);
}
void r0_expand_zero_interleave(Ctx* c, uint32_t* out, const uint32_t* in, size_t n_in, int bits) {
  LAUNCH_1D(k_expand_zero, n_in << bits, out, in, n_in << bits, bits);
}
void r0_eltwise_mul_factor(Ctx* c, uint32_t* io, uint32_t factor, size_t n) { LAUNCH_1D(k_mul_factor, n, io, factor, n); }
void r0_copy_elem_slice(Ctx* c, uint32_t* into, const uint32_t* from_host, size_t from_rows, size_t from_cols,
                        size_t from_offset, size_t from_stride, size_t into_offset, size_t into_stride) {
  if (from_rows == 0 || from_cols == 0) return;
  size_t span = (from_rows - 1) * from_stride + from_cols;
  Scratch d_from(c, from_host + from_offset, span * 4);
  LAUNCH_1D(k_copy_region, from_rows * from_cols, into, d_from.as<uint32_t>(), from_rows, from_cols, from_stride,
            into_offset, into_stride);
}
void r0_prefix_products(Ctx* c, uint32_t* io, size_t n) {
  k_prefix_products<<<1, 32, 0, c->stream>>>((FpExt*)io, n);
  count_launch(c);
  R0_CUDA(cudaGetLastError());
}

void r0_combos_prepare(Ctx* c, uint32_t* combos, const FpExt* coeff_u_host, size_t coeff_u_len, uint32_t combo_count,
                       size_t cycles, const uint32_t* reg_sizes, const uint32_t* reg_combo_ids, uint32_t nregs,
                       const FpExt& mix, uint32_t check_size) {
  PhaseScope ph(c, "combos_prepare");
  std::vector<FpExt> pows(nregs + check_size);
  std::vector<uint32_t> pos(nregs);
  FpExt cur = ext_one();
  uint32_t cur_pos = 0;
  for (uint32_t r = 0; r < nregs; r++) {
    R0_CHECK(reg_sizes[r] <= 32, "combos_prepare: register with more than 32 taps");
    pows[r] = cur;
    pos[r] = cur_pos;
    cur = ext_mul(cur, mix);
    cur_pos += reg_sizes[r];
  }
  for (uint32_t k = 0; k < check_size; k++) {
    pows[nregs + k] = cur;
    cur = ext_mul(cur, mix);
  }
  R0_CHECK(cur_pos + check_size <= coeff_u_len, "combos_prepare: coeff_u too short");
  Scratch d_u(c, coeff_u_host, coeff_u_len * sizeof(FpExt));
  Scratch d_sz(c, reg_sizes, nregs * 4);
  Scratch d_id(c, reg_combo_ids, nregs * 4);
  Scratch d_pos(c, pos.data(), nregs * 4);
  Scratch d_pw(c, pows.data(), pows.size() * sizeof(FpExt));
  k_combos_prepare<<<1, 256, 0, c->stream>>>((FpExt*)combos, d_u.as<FpExt>(), combo_count, cycles, d_sz.as<uint32_t>(),
                                             d_id.as<uint32_t>(), d_pos.as<uint32_t>(), nregs, d_pw.as<FpExt>(),
                                             check_size, cur_pos);
  count_launch(c);
  R0_CUDA(cudaGetLastError());
}

// polys[j]: n FpExt coefficients, divided in place by (x - zs[j]); the remainders go to remainders_dev[j]. The jobs must
// be independent (different polynomials); successive divisions of one polynomial are successive calls.
void r0_poly_divide_batch(Ctx* c, uint32_t* const* polys, const FpExt* zs, uint32_t* const* remainders_dev, size_t njobs,
                          size_t n) {
  if (n == 0 || njobs == 0) return;
  PhaseScope ph(c, "combos_divide", 32.0 * (double)n * (double)njobs);
  const size_t nblocks = (n + DIV_SEG - 1) / DIV_SEG;
  R0_CHECK(nblocks <= (size_t(1) << DIV_MAXLG), "poly_divide: polynomial too long");
  R0_CHECK(njobs <= 65535, "poly_divide: too many jobs");
  std::vector<DivPowers> pws(njobs);
  std::vector<DivJob> jobs(njobs);
  for (size_t j = 0; j < njobs; j++) {
    DivPowers& pw = pws[j];
    pw.z = zs[j];
    FpExt z8 = ext_pow(zs[j], DIV_E);
    FpExt cur = z8;
    for (int s = 0; s < 8; s++) {
      pw.zs[s] = cur;
      cur = ext_mul(cur, cur);
    }
    // cur = z^(8 * 256) = z^SEG
    for (int s = 0; s <= DIV_MAXLG; s++) {
      pw.Z[s] = cur;
      cur = ext_mul(cur, cur);
    }
    FpExt acc = ext_one();
    for (int t = DIV_B - 1; t >= 0; t--) {
      pw.zt[t] = acc;
      acc = ext_mul(acc, z8);
    }
    jobs[j] = DivJob{(FpExt*)polys[j], (FpExt*)remainders_dev[j]};
  }
  Scratch d_pw(c, pws.data(), njobs * sizeof(DivPowers));
  Scratch d_jobs(c, jobs.data(), njobs * sizeof(DivJob));
  Scratch totals(c, njobs * nblocks * sizeof(FpExt));
  Scratch carry(c, njobs * nblocks * sizeof(FpExt));
  const dim3 grid((unsigned)nblocks, (unsigned)njobs);
  k_div_block_totals<<<grid, DIV_B, 0, c->stream>>>(totals.as<FpExt>(), d_jobs.as<DivJob>(), n, d_pw.as<DivPowers>());
  k_div_block_carries<<<(unsigned)njobs, 1024, 0, c->stream>>>(carry.as<FpExt>(), totals.as<FpExt>(), nblocks,
                                                                d_pw.as<DivPowers>());
  k_div_quotient<<<grid, DIV_B, 0, c->stream>>>(d_jobs.as<DivJob>(), n, carry.as<FpExt>(), d_pw.as<DivPowers>());
  count_launch(c, 3);
  R0_CUDA(cudaGetLastError());
}

void r0_poly_divide(Ctx* c, uint32_t* poly, size_t n, const FpExt& z, uint32_t* remainder_dev) {
  r0_poly_divide_batch(c, &poly, &z, &remainder_dev, 1, n);
}

// ---- batched query openings (product driver) ------------------------------------------------------------------
void r0_gather_batched(Ctx* c, uint32_t* dst, const GatherJob* jobs_host, size_t njobs) {
  PhaseScope ph(c, "gather");
  if (njobs == 0) return;
  Scratch d_jobs(c, jobs_host, njobs * sizeof(GatherJob));
  k_gather_batched<<<(unsigned)njobs, 128, 0, c->stream>>>(dst, d_jobs.as<GatherJob>());
  count_launch(c);
  R0_CUDA(cudaGetLastError());
}
void r0_gather_digests(Ctx* c, uint32_t* dst, const DigestJob* jobs_host, size_t njobs) {
  PhaseScope ph(c, "gather");
  if (njobs == 0) return;
  Scratch d_jobs(c, jobs_host, njobs * sizeof(DigestJob));
  k_gather_digests<<<(unsigned)((njobs * 8 + 255) / 256), 256, 0, c->stream>>>(dst, d_jobs.as<DigestJob>(), njobs);
  count_launch(c);
  R0_CUDA(cudaGetLastError());
}
