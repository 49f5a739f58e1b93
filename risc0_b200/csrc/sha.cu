// SHA-256 row hashing and Merkle folding for the "sha-256" hash suite, sm_100a.
//
// Replaces: Hal::hash_rows / hash_fold with Sha256HashSuite (risc0/zkp/src/core/hash/sha/mod.rs:312-326 ->
// sha/cpu.rs:56-98): the raw little-endian words of the row are compressed block by block WITHOUT padding or length,
// a short last block is zero-filled, and the digest words are the byte-swapped state. Reference GPU kernels:
// risc0/sys/kernels/zkp/cuda/sha.cu:17-29 + sha256.h:156-237. One thread per row / node, state and message schedule
// in registers (rolling 16-word window).
#include "ctx.h"

namespace r0 {

__constant__ uint32_t c_sha_k[64] = {
    0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5, 0xd807aa98,
    0x12835b01, 0x243185be, 0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174, 0xe49b69c1, 0xefbe4786,
    0x0fc19dc6, 0x240ca1cc, 0x2de92c6f, 0x4a7484aa, 0x5cb0a9dc, 0x76f988da, 0x983e5152, 0xa831c66d, 0xb00327c8,
    0xbf597fc7, 0xc6e00bf3, 0xd5a79147, 0x06ca6351, 0x14292967, 0x27b70a85, 0x2e1b2138, 0x4d2c6dfc, 0x53380d13,
    0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85, 0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3, 0xd192e819,
    0xd6990624, 0xf40e3585, 0x106aa070, 0x19a4c116, 0x1e376c08, 0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a,
    0x5b9cca4f, 0x682e6ff3, 0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208, 0x90befffa, 0xa4506ceb, 0xbef9a3f7,
    0xc67178f2};

__device__ __forceinline__ uint32_t rotr(uint32_t x, int n) { return __funnelshift_r(x, x, n); }
__device__ __forceinline__ uint32_t bswap(uint32_t x) { return __byte_perm(x, 0, 0x0123); }

__device__ __forceinline__ void sha_init(uint32_t (&st)[8]) {
  st[0] = 0x6a09e667; st[1] = 0xbb67ae85; st[2] = 0x3c6ef372; st[3] = 0xa54ff53a;
  st[4] = 0x510e527f; st[5] = 0x9b05688c; st[6] = 0x1f83d9ab; st[7] = 0x5be0cd19;
}

__device__ __forceinline__ void sha_compress(uint32_t (&st)[8], uint32_t (&w)[16]) {
  uint32_t a = st[0], b = st[1], c = st[2], d = st[3], e = st[4], f = st[5], g = st[6], h = st[7];
#pragma unroll
  for (int i = 0; i < 64; i++) {
    uint32_t wi;
    if (i < 16) {
      wi = w[i];
    } else {
      uint32_t w15 = w[(i + 1) & 15], w2 = w[(i + 14) & 15];
      uint32_t s0 = rotr(w15, 7) ^ rotr(w15, 18) ^ (w15 >> 3);
      uint32_t s1 = rotr(w2, 17) ^ rotr(w2, 19) ^ (w2 >> 10);
      wi = w[i & 15] + s0 + w[(i + 9) & 15] + s1;
      w[i & 15] = wi;
    }
    uint32_t S1 = rotr(e, 6) ^ rotr(e, 11) ^ rotr(e, 25);
    uint32_t ch = (e & f) ^ (~e & g);
    uint32_t t1 = h + S1 + ch + c_sha_k[i] + wi;
    uint32_t S0 = rotr(a, 2) ^ rotr(a, 13) ^ rotr(a, 22);
    uint32_t mj = (a & b) ^ (a & c) ^ (b & c);
    uint32_t t2 = S0 + mj;
    h = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
  }
  st[0] += a; st[1] += b; st[2] += c; st[3] += d; st[4] += e; st[5] += f; st[6] += g; st[7] += h;
}

__global__ void __launch_bounds__(256) sha_rows_kernel(uint32_t* __restrict__ out, const uint32_t* __restrict__ matrix,
                                                     size_t rows, uint32_t cols) {
  size_t row = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= rows) return;
  uint32_t st[8], w[16];
  sha_init(st);
  for (uint32_t done = 0; done < cols; done += 16) {
#pragma unroll
    for (int i = 0; i < 16; i++) w[i] = (done + i < cols) ? bswap(matrix[(size_t)(done + i) * rows + row]) : 0u;
    sha_compress(st, w);
  }
#pragma unroll
  for (int i = 0; i < 8; i++) out[row * 8 + i] = bswap(st[i]);
}

__global__ void __launch_bounds__(256) sha_fold_kernel(uint32_t* io, size_t in_size, size_t out_size) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= out_size) return;
  uint32_t st[8], w[16];
  sha_init(st);
  const uint32_t* in = io + (in_size + 2 * i) * 8;
#pragma unroll
  for (int k = 0; k < 16; k++) w[k] = bswap(in[k]);
  sha_compress(st, w);
#pragma unroll
  for (int k = 0; k < 8; k++) io[(out_size + i) * 8 + k] = bswap(st[k]);
}

// generic pair hash with independent pointers (risc0_zkp_cuda_sha_fold's signature): out[i] = H(in[2i] || in[2i+1])
__global__ void __launch_bounds__(256) sha_fold_pairs_kernel(uint32_t* out, const uint32_t* in, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t st[8], w[16];
  sha_init(st);
#pragma unroll
  for (int k = 0; k < 16; k++) w[k] = bswap(in[16 * i + k]);
  sha_compress(st, w);
#pragma unroll
  for (int k = 0; k < 8; k++) out[i * 8 + k] = bswap(st[k]);
}

}  // namespace r0

using namespace r0;

void r0_sha_fold_pairs(Ctx* c, uint32_t* out, const uint32_t* in, size_t n) {
  PhaseScope ph(c, "hash_fold", 96.0 * (double)n);
  if (n == 0) return;
  sha_fold_pairs_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(out, in, n);
  count_launch(c);
  R0_CUDA(cudaGetLastError());
}

void r0_sha_hash_rows(Ctx* c, uint32_t* out, const uint32_t* matrix, size_t rows, size_t cols) {
  PhaseScope ph(c, "hash_rows", 4.0 * (double)rows * (double)cols + 32.0 * (double)rows);
  if (rows == 0) return;
  sha_rows_kernel<<<(unsigned)((rows + 255) / 256), 256, 0, c->stream>>>(out, matrix, rows, (uint32_t)cols);
  count_launch(c);
  R0_CUDA(cudaGetLastError());
}
void r0_sha_hash_fold(Ctx* c, uint32_t* io, size_t in_size, size_t out_size) {
  PhaseScope ph(c, "hash_fold", 96.0 * (double)out_size);
  R0_CHECK(in_size == 2 * out_size, "hash_fold: input_size must be 2 * output_size");
  if (out_size == 0) return;
  sha_fold_kernel<<<(unsigned)((out_size + 255) / 256), 256, 0, c->stream>>>(io, in_size, out_size);
  count_launch(c);
  R0_CUDA(cudaGetLastError());
}
