/* Plain-C user of the r0b200 C ABI (no Python, no C++): what a cgo / Rust-FFI caller does. Interpolates a batch of
 * columns, evaluates them back without expansion and checks the round trip; then builds a Poseidon2 Merkle tree.
 *
 *   gcc -std=c99 -Iinclude examples/hal_roundtrip.c -Lrisc0_b200/lib -lr0b200 -Wl,-rpath,$PWD/risc0_b200/lib -o /tmp/hal_roundtrip
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "r0b200.h"

#define CHECK(call)                                  \
  do {                                               \
    r0b200_err e_ = (call);                          \
    if (e_) {                                        \
      fprintf(stderr, "%s failed: %s\n", #call, e_); \
      r0b200_free_error(e_);                         \
      return 1;                                      \
    }                                                \
  } while (0)

int main(void) {
  const uint32_t P = 15u * (1u << 27) + 1u;
  const uint32_t lg_n = 12;
  const size_t n = (size_t)1 << lg_n, cols = 7, words = n * cols;
  r0b200_ctx* ctx = NULL;
  CHECK(r0b200_create(0, &ctx));
  uint32_t* host = (uint32_t*)malloc(words * 4);
  uint32_t* back = (uint32_t*)malloc(words * 4);
  uint64_t s = 88172645463325252ull;
  for (size_t i = 0; i < words; i++) { /* xorshift: any canonical word is a valid Montgomery element */
    s ^= s << 13;
    s ^= s >> 7;
    s ^= s << 17;
    host[i] = (uint32_t)(s % P);
  }
  void *d_io = NULL, *d_out = NULL, *d_nodes = NULL;
  CHECK(r0b200_alloc(ctx, words * 4, &d_io));
  CHECK(r0b200_alloc(ctx, words * 4, &d_out));
  CHECK(r0b200_alloc(ctx, 2 * n * 32, &d_nodes));
  CHECK(r0b200_copy_h2d(ctx, d_io, host, words * 4));
  CHECK(r0b200_batch_interpolate_ntt(ctx, (uint32_t*)d_io, cols, lg_n));
  CHECK(r0b200_batch_expand_into_evaluate_ntt(ctx, (uint32_t*)d_out, (const uint32_t*)d_io, cols, lg_n, 0));
  CHECK(r0b200_copy_d2h(ctx, back, d_out, words * 4));
  if (memcmp(host, back, words * 4) != 0) {
    fprintf(stderr, "NTT round trip mismatch\n");
    return 1;
  }
  CHECK(r0b200_merkle_build(ctx, R0B200_HASH_POSEIDON2, (uint32_t*)d_nodes, (const uint32_t*)d_out, n, cols));
  uint32_t root[8];
  CHECK(r0b200_copy_d2h(ctx, root, (const char*)d_nodes + 32, 32));
  printf("ok: %zu columns of 2^%u round-tripped; Merkle root %08x %08x ...; %llu kernel launches\n", cols, lg_n, root[0],
         root[1], (unsigned long long)r0b200_launch_count(ctx));
  CHECK(r0b200_free(ctx, d_io));
  CHECK(r0b200_free(ctx, d_out));
  CHECK(r0b200_free(ctx, d_nodes));
  r0b200_destroy(ctx);
  free(host);
  free(back);
  return 0;
}
