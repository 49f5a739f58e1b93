/* r0b200 — C ABI of the B200-native (sm_100a) backend for the RISC Zero STARK prover hot path.
 *
 * This is the drop-in boundary: plain pointers and sizes, no C++ or torch types. Each entry point replaces one
 * method of the reference `risc0_zkp::hal::Hal` trait (risc0/zkp/src/hal/mod.rs:55-258), of `CircuitHal`
 * (hal/mod.rs:265-290), or one symbol of the reference's existing FFI tables
 * (`risc0_zkp_cuda_*`: risc0/sys/kernels/zkp/cuda/ffi.cu:25-145, `sppark_*`: risc0/sys/src/cuda.rs:19-80,
 * `risc0_circuit_rv32im_cuda_eval_check`: risc0/circuit/rv32im-sys/src/lib.rs:87-133). INTEGRATION.md shows the Rust
 * `impl Hal for B200Hal` that binds them.
 *
 * Conventions
 *  - Error convention is the reference's (risc0/sys/src/lib.rs:53-75): every call returns `const char*`; NULL means
 *    success, otherwise a heap string the caller releases with r0b200_free_error() (libc free works too).
 *  - All `uint32_t*` buffers are DEVICE pointers unless the parameter name ends in `_host`. Field elements are
 *    BabyBear Montgomery words exactly as in the reference's Buffer<Elem> (canonical, < P, or 0xffffffff = INVALID);
 *    extension elements are 4 consecutive words; digests are 8 words. Matrices are column-major: buf[col*rows + row].
 *  - Work is stream-ordered on the context's stream and returns without waiting for the GPU (the reference's CudaHal
 *    blocks on every call, hal/cuda.rs + cuda.h:77-100). Results are visible to the host after r0b200_copy_d2h()
 *    (which synchronises, like Buffer::view / get_at) or r0b200_sync().
 *  - One context = one device ordinal + one stream + one prover at a time (the reference hard-codes device 0,
 *    hal/cuda.rs:406); contexts on different devices are independent, which is how segments shard across GPUs.
 */
#ifndef R0B200_H
#define R0B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct r0b200_ctx r0b200_ctx;
typedef const char* r0b200_err;

enum { R0B200_HASH_POSEIDON2 = 0, R0B200_HASH_SHA256 = 1 };
enum { R0B200_CIRCUIT_RV32IM = 0, R0B200_CIRCUIT_RECURSION = 1 };

/* ---- context (CudaHal::new_from_hash, hal/cuda.rs:397-421; sppark_init, supra/ntt.cu:4-32) ---- */
r0b200_err r0b200_create(int device, r0b200_ctx** out);
void r0b200_destroy(r0b200_ctx* ctx);
void r0b200_free_error(const char* err);
r0b200_err r0b200_sync(r0b200_ctx* ctx);
void* r0b200_stream(r0b200_ctx* ctx);              /* the cudaStream_t, for callers that enqueue their own work */
uint64_t r0b200_launch_count(r0b200_ctx* ctx);     /* kernels launched so far through this context */
uint64_t r0b200_bytes_peak(r0b200_ctx* ctx);       /* MemoryTracker.peak analogue, hal/mod.rs:297-317 */
/* CUDA-event timer on the context's stream (the stream the kernels are launched on): start records an event, stop
 * records a second one, waits for it and returns the elapsed device time in milliseconds. */
r0b200_err r0b200_timer_start(r0b200_ctx* ctx);
r0b200_err r0b200_timer_stop(r0b200_ctx* ctx, float* ms);

/* Per-op device timing (the reference's NVTX ranges / scope! timers, hal/cuda.rs + risc0_core::scope): begin clears
 * and enables; end synchronises and writes JSON {"op": {"ms": device ms, "n": count, "bytes": algorithmic bytes}}. */
r0b200_err r0b200_profile_begin(r0b200_ctx* ctx);
r0b200_err r0b200_profile_end(r0b200_ctx* ctx, char* json_out, size_t cap);

/* ---- buffers (Hal::alloc_* / copy_from_* hal/mod.rs:67-100; Buffer::view/get_at/view_mut hal/mod.rs:39-53) ---- */
r0b200_err r0b200_alloc(r0b200_ctx* ctx, size_t bytes, void** dptr);
r0b200_err r0b200_free(r0b200_ctx* ctx, void* dptr);
r0b200_err r0b200_copy_h2d(r0b200_ctx* ctx, void* dst, const void* src_host, size_t bytes);
r0b200_err r0b200_copy_d2h(r0b200_ctx* ctx, void* dst_host, const void* src, size_t bytes); /* synchronises */
r0b200_err r0b200_fill_u32(r0b200_ctx* ctx, uint32_t* dst, uint32_t word, size_t count);   /* alloc_elem_init */

/* ---- NTT family ---- */
/* Hal::batch_interpolate_ntt (cpu.rs:342-350; sppark_batch_iNTT): `count` rows of 2^lg_n, in place;
 * natural-order values in, bit-reversed coefficients out, scaled by 2^-lg_n. */
r0b200_err r0b200_batch_interpolate_ntt(r0b200_ctx* ctx, uint32_t* io, size_t count, uint32_t lg_n);
/* Fused make_coeffs tail (prove/prover.rs:38-48): batch_interpolate_ntt followed by zk_shift in one pass. */
r0b200_err r0b200_batch_interpolate_ntt_zk(r0b200_ctx* ctx, uint32_t* io, size_t count, uint32_t lg_n);
/* Hal::zk_shift (cpu.rs:395-408; sppark_batch_zk_shift): io[p*n + i] *= 3^brev(i). */
r0b200_err r0b200_zk_shift(r0b200_ctx* ctx, uint32_t* io, size_t count, uint32_t lg_n);
/* Hal::batch_expand_into_evaluate_ntt (cpu.rs:305-340; sppark_batch_expand + sppark_batch_NTT): rows of
 * 2^lg_in bit-reversed coefficients -> rows of 2^(lg_in+expand_bits) natural-order evaluations. expand_bits in {0, 2}. */
r0b200_err r0b200_batch_expand_into_evaluate_ntt(r0b200_ctx* ctx, uint32_t* out, const uint32_t* in, size_t count,
                                                 uint32_t lg_in, uint32_t expand_bits);
/* Hal::batch_bit_reverse (cpu.rs:352-360; risc0_zkp_cuda_batch_bit_reverse). */
r0b200_err r0b200_batch_bit_reverse(r0b200_ctx* ctx, uint32_t* io, size_t count, uint32_t lg_n);

/* ---- hashing / Merkle ---- */
/* Hal::hash_rows (cpu.rs:555-567; sppark_poseidon2_rows / risc0_zkp_cuda_sha_rows): out[r] = H(matrix[:, r]). */
r0b200_err r0b200_hash_rows(r0b200_ctx* ctx, int hash, uint32_t* out_digests, const uint32_t* matrix, size_t rows,
                            size_t cols);
/* Hal::hash_fold (cpu.rs:569-581; sppark_poseidon2_fold / risc0_zkp_cuda_sha_fold): one heap level. */
r0b200_err r0b200_hash_fold(r0b200_ctx* ctx, int hash, uint32_t* io_digests, size_t input_size, size_t output_size);
/* MerkleTreeProver::new (prove/merkle.rs:54-81) in one call: nodes[rows..2*rows) = hash_rows, then every fold level;
 * nodes has 2*rows digests. */
r0b200_err r0b200_merkle_build(r0b200_ctx* ctx, int hash, uint32_t* nodes, const uint32_t* matrix, size_t rows,
                               size_t cols);

/* ---- element-wise / FRI / mixing ---- */
r0b200_err r0b200_eltwise_add_elem(r0b200_ctx* ctx, uint32_t* out, const uint32_t* a, const uint32_t* b, size_t n);
r0b200_err r0b200_eltwise_copy_elem(r0b200_ctx* ctx, uint32_t* out, const uint32_t* in, size_t n);
r0b200_err r0b200_eltwise_zeroize_elem(r0b200_ctx* ctx, uint32_t* io, size_t n);
/* Hal::eltwise_sum_extelem (cpu.rs:475-500): in = to_add x count AoS ext elems -> out = 4 SoA planes of count. */
r0b200_err r0b200_eltwise_sum_extelem(r0b200_ctx* ctx, uint32_t* out, const uint32_t* in, size_t count, size_t to_add);
/* Hal::eltwise_copy_elem_slice (cpu.rs:617-635): strided copy from a HOST slice. */
r0b200_err r0b200_eltwise_copy_elem_slice(r0b200_ctx* ctx, uint32_t* into, const uint32_t* from_host, size_t from_rows,
                                          size_t from_cols, size_t from_offset, size_t from_stride, size_t into_offset,
                                          size_t into_stride);
/* Hal::fri_fold (cpu.rs:524-553): out = 4 planes of count; in = 4 planes of 16*count; mix = 4 host words. */
r0b200_err r0b200_fri_fold(r0b200_ctx* ctx, uint32_t* out, const uint32_t* in, size_t count, const uint32_t* mix_host);
/* Hal::mix_poly_coeffs (cpu.rs:410-455): out (ext, combos x count) += mix_start*mix^i * in[i] for column i.
 * combos_host: combo id per input column (the reference passes it as a device u32 buffer built from host data). */
r0b200_err r0b200_mix_poly_coeffs(r0b200_ctx* ctx, uint32_t* out, const uint32_t* mix_start_host,
                                  const uint32_t* mix_host, const uint32_t* in, const uint32_t* combos_host,
                                  size_t input_size, size_t count);
/* Hal::batch_evaluate_any (cpu.rs:362-393): out[e] = sum_i coeffs[which[e]*n + i] * xs[e]^i ; which/xs/out on device. */
r0b200_err r0b200_batch_evaluate_any(r0b200_ctx* ctx, const uint32_t* coeffs, size_t poly_count, uint32_t lg_n,
                                     const uint32_t* which, const uint32_t* xs, uint32_t* out, size_t eval_count);
/* Hal::gather_sample (cpu.rs:583-596). */
r0b200_err r0b200_gather_sample(r0b200_ctx* ctx, uint32_t* dst, const uint32_t* src, size_t idx, size_t size,
                                size_t stride);
/* Hal::scatter (cpu.rs:598-615): CSR scatter from HOST index/offsets/values. */
r0b200_err r0b200_scatter(r0b200_ctx* ctx, uint32_t* into, const uint32_t* index_host, size_t index_len,
                          const uint32_t* offsets_host, const uint32_t* values_host);
/* Hal::prefix_products (cpu.rs:637-642). */
r0b200_err r0b200_prefix_products(r0b200_ctx* ctx, uint32_t* io_ext, size_t n);
/* Hal::combos_prepare (hal/mod.rs:202-234). */
r0b200_err r0b200_combos_prepare(r0b200_ctx* ctx, uint32_t* combos, const uint32_t* coeff_u_host, size_t coeff_u_len,
                                 uint32_t combo_count, size_t cycles, const uint32_t* reg_sizes_host,
                                 const uint32_t* reg_combo_ids_host, uint32_t nregs, const uint32_t* mix_host);
/* Hal::combos_divide (hal/mod.rs:236-257): chunk i (cycles ext elems) is divided in place by (x - pows[k]) for
 * k in [pow_begin[i], pow_begin[i+1]). Fails if any remainder is non-zero (the reference asserts). Synchronises. */
r0b200_err r0b200_combos_divide(r0b200_ctx* ctx, uint32_t* combos, size_t nchunks, const uint32_t* pow_begin_host,
                                const uint32_t* pows_host, size_t cycles);

/* ---- circuit (CircuitHal) ---- */
/* CircuitHal::eval_check for the rv32im circuit (hal/mod.rs:279-289; risc0_circuit_rv32im_cuda_eval_check,
 * rv32im-sys/src/lib.rs:87-133; CPU spec rv32im/src/prove/hal/cpu.rs:145-208). check: 4 x (4 << po2) words out;
 * groups are the evaluated matrices (cols x (4 << po2)): ctrl = code group (1 col, unused by the constraints, kept for
 * signature parity), data (211), accum (103); mix (36 words) and out = globals (90 words) are device buffers;
 * poly_mix_host = 4 words. The poly-mix power table and the root of unity are derived inside. */
r0b200_err r0b200_eval_check_rv32im(r0b200_ctx* ctx, uint32_t* check, const uint32_t* ctrl, const uint32_t* data,
                                    const uint32_t* accum, const uint32_t* mix, const uint32_t* out,
                                    const uint32_t* poly_mix_host, uint32_t po2);

/* Same for the recursion circuit (risc0_circuit_recursion_cpu_eval_check, recursion-sys/kernels/cxx/ffi.cpp:219-246;
 * recursion/src/prove/hal/cpu.rs:106-150): ctrl 23, data 128, accum 12 columns; mix 20 words, out 32 words. */
r0b200_err r0b200_eval_check_recursion(r0b200_ctx* ctx, uint32_t* check, const uint32_t* ctrl, const uint32_t* data,
                                       const uint32_t* accum, const uint32_t* mix, const uint32_t* out,
                                       const uint32_t* poly_mix_host, uint32_t po2);

/* ---- whole segment (the Hal's caller on the hot path) ---- */
/* prove_core's prove_inner block (rv32im/src/prove/hal/mod.rs:171-222 -> zkp/src/prove/prover.rs:81-393 ->
 * prove/fri.rs:77-126), split where the protocol splits it: the accum matrix is computed FROM the mix values the
 * transcript yields after code and data have been committed (hal/mod.rs:209-217: commit_group(CODE), commit_group(DATA),
 * mix = MIX_SIZE x iop.random_elem(), witgen.accum(.., &mix), commit_group(ACCUM), finalize), so a caller cannot hand
 * all three matrices over at once.
 *
 *   r0b200_prove_begin : version word, info + header commits, code (1 x N) and data (211 x N) groups, N = 2^po2; draws
 *                        the mix and returns it (36 words rv32im, 20 recursion) together with a proof handle.
 *   r0b200_prove_finish: commits accum (103 x N), runs eval_check, DEEP and FRI, writes the seal (u32 words, the
 *                        reference's seal format) to host memory and releases the handle (also on failure).
 *   r0b200_prove_abort : releases a handle without finishing.
 *
 * code/data/accum are column-major witness matrices: device pointers, or host pointers when the *_on_host flag is set
 * (copied with stream-ordered H2D copies on the context's copy stream, chunk by chunk, overlapped with compute; pinned
 * memory makes them asynchronous; they must stay valid until the call that consumes them returns). A device-resident
 * witness is left untouched (the first NTT pass reads it out of place), so the data matrix can still be read by the
 * accum step between begin and finish. `uploaded` (may be NULL) is a witness from r0b200_witness_upload; when given,
 * code/data/po2 are taken from it. global_host: 90 words (rv32im) / 32 (recursion).
 * Optional outputs of finish (may be NULL): every committed Merkle root in commit order, and the 50 query positions. */
typedef struct r0b200_proof r0b200_proof;
typedef struct r0b200_witness r0b200_witness;
r0b200_err r0b200_prove_begin(r0b200_ctx* ctx, int circuit, int hash, uint32_t po2, const uint32_t* code,
                              const uint32_t* data, int witness_on_host, r0b200_witness* uploaded,
                              const uint32_t* global_host, uint32_t* mix_out_host, size_t mix_cap, r0b200_proof** out);
r0b200_err r0b200_prove_finish(r0b200_proof* proof, const uint32_t* accum, int accum_on_host, uint32_t* seal_out_host,
                               size_t seal_cap, size_t* seal_len, uint32_t* roots_out_host, size_t roots_cap,
                               size_t* nroots, uint32_t* query_pos_out_host);
void r0b200_prove_abort(r0b200_proof* proof);

/* One-call form = prove_begin + prove_finish with an accum matrix that was fixed BEFORE the mix was drawn. Only
 * meaningful for witnesses whose accum does not depend on the mix (the synthetic witnesses of the parity tests and
 * micro-benchmarks); a real prove_core must use the two-phase form above. */
r0b200_err r0b200_prove_rv32im(r0b200_ctx* ctx, int hash, uint32_t po2, const uint32_t* code, const uint32_t* data,
                               const uint32_t* accum, int witness_on_host, const uint32_t* global_host,
                               uint32_t* seal_out_host, size_t seal_cap, size_t* seal_len, uint32_t* roots_out_host,
                               size_t roots_cap, size_t* nroots, uint32_t* query_pos_out_host);

/* RecursionProverImpl::prove's prove block (recursion/src/prove/mod.rs:179-224), one-call form (same caveat): ctrl
 * (23 x N), data (128 x N), accum (12 x N), global_host 32 words; lift / join / resolve programs all have this shape
 * (po2 = 18 by default). No seal version word. */
r0b200_err r0b200_prove_recursion(r0b200_ctx* ctx, int hash, uint32_t po2, const uint32_t* ctrl, const uint32_t* data,
                                  const uint32_t* accum, int witness_on_host, const uint32_t* global_host,
                                  uint32_t* seal_out_host, size_t seal_cap, size_t* seal_len, uint32_t* roots_out_host,
                                  size_t roots_cap, size_t* nroots, uint32_t* query_pos_out_host);

/* ---- witness generation on the device (CircuitWitnessGenerator / CircuitAccumulator) ----
 * rv32im/src/prove/hal/mod.rs:82-102; replaces risc0_circuit_rv32im_cuda_witgen / _cuda_accum
 * (rv32im-sys/src/lib.rs:105-119, kernels/cuda/ffi.cu:431-512; CPU spec kernels/cxx/ffi.cpp:265-365).
 * r0b200_preflight_trace is RawPreflightTrace (rv32im-sys/src/lib.rs:63-72) with HOST pointers: `cycles` = array of
 * RawPreflightCycle (36 bytes each), `txns` = array of RawMemoryTransaction (20 bytes each). */
typedef struct {
  const void* cycles;
  const void* txns;
  const uint8_t* bigint_bytes;
  uint32_t txns_len;
  uint32_t bigint_bytes_len;
  uint32_t table_split_cycle;
} r0b200_preflight_trace;
typedef struct r0b200_trace r0b200_trace;
/* Copies a trace of `cycles` (= 2^po2) cycles to the device (stream-ordered; pinned memory makes it asynchronous) and
 * bucket-sorts the cycles of each phase by (major, minor) there, which is the order the step kernels walk. */
r0b200_err r0b200_trace_upload(r0b200_ctx* ctx, const r0b200_preflight_trace* trace_host, uint32_t cycles,
                               r0b200_trace** out);
void r0b200_trace_free(r0b200_trace* trace);
/* generate_witness: fills `data` (211 x cycles, device) and the output cells of `global` (90 words, device) for every
 * cycle. As in WitnessGenerator::new (prove/witgen/mod.rs:144-165) the caller has filled data with INVALID and applied
 * the injector scatter, and uploaded the global vector. mode: StepMode (0 parallel, 1 forward, 2 reverse) - the step
 * functions are order-independent inside a phase, so all three run the same sorted parallel schedule. Synchronises and
 * reports the first failed check ("Inconsistent set", "Read of unset value", eqz, trace mismatch) as the error. */
r0b200_err r0b200_witgen_rv32im(r0b200_ctx* ctx, uint32_t mode, r0b200_trace* trace, uint32_t* global, uint32_t* data);
/* step_accum: fills `accum` (103 x cycles, device; INVALID-filled, bigint-accum cells scattered by the caller) from
 * data, global and mix (36 words, device), then the 4-column prefix sum and the back-propagation of the totals. */
r0b200_err r0b200_accum_rv32im(r0b200_ctx* ctx, r0b200_trace* trace, uint32_t* data, uint32_t* accum, uint32_t* global,
                               uint32_t* mix);

/* SegmentProverImpl::prove_core (rv32im/src/prove/hal/mod.rs:143-224) from a PreflightResults
 * (prove/witgen/mod.rs:55-88), everything on the device: WitnessGenerator::new (INVALID fill, injector scatter,
 * generate_witness, zeroize), commit code + data, mix draw, WitnessGenerator::accum (step_accum from that mix), commit
 * accum, finalize -> seal.
 *   r0b200_segment_upload : copies a segment's prover inputs to the device on the context's COPY stream and returns at
 *                           once: the trace, the global vector (90 words, INVALID where the witness generator fills
 *                           the value) and the injector as the reference builds it (CSR: index[cycles + 1], word
 *                           offsets col * cycles + row, Montgomery values; witgen/mod.rs:226-270,320-370). Host
 *                           buffers must stay valid until the segment has been proved (pinned memory makes the copies
 *                           asynchronous). Uploading segment s + 1 before proving segment s overlaps the transfer with
 *                           compute - the reference's CPU -> GPU queue (r0vm/src/actors/worker.rs:70-76,585).
 *   r0b200_prove_segment  : proves an uploaded segment (it can be proved again; free it with r0b200_segment_free).
 *                           global_out_host (90 words, may be NULL) receives the globals after witness generation.
 *   r0b200_prove_segment_rv32im : upload + prove + free in one call.
 * Segments with bigint cycles: the BigIntAccumState cells of those rows depend on the mix (witgen/mod.rs:186-207,
 * byte_poly.rs:403-475). r0b200_segment_upload reads what they need out of the trace while it is still on the host
 * (the 16 witness bytes and the verify-program word of each bigint cycle), and r0b200_prove_segment evaluates them at
 * the drawn mix and scatters them into accum before step_accum - nothing extra to pass. A caller that owns witness
 * generation itself (r0b200_prove_begin / r0b200_accum_rv32im / r0b200_prove_finish) scatters them as the reference does. */
typedef struct r0b200_segment r0b200_segment;
r0b200_err r0b200_segment_upload(r0b200_ctx* ctx, uint32_t po2, const r0b200_preflight_trace* trace_host,
                                 const uint32_t* global_host, const uint32_t* inj_index_host, size_t inj_index_len,
                                 const uint32_t* inj_offsets_host, const uint32_t* inj_values_host, r0b200_segment** out);
void r0b200_segment_free(r0b200_segment* segment);
r0b200_err r0b200_prove_segment(r0b200_ctx* ctx, int hash, r0b200_segment* segment, uint32_t* seal_out_host,
                                size_t seal_cap, size_t* seal_len, uint32_t* roots_out_host, size_t roots_cap,
                                size_t* nroots, uint32_t* query_pos_out_host, uint32_t* global_out_host);
r0b200_err r0b200_prove_segment_rv32im(r0b200_ctx* ctx, int hash, uint32_t po2, const r0b200_preflight_trace* trace_host,
                                       const uint32_t* global_host, const uint32_t* inj_index_host, size_t inj_index_len,
                                       const uint32_t* inj_offsets_host, const uint32_t* inj_values_host,
                                       uint32_t* seal_out_host, size_t seal_cap, size_t* seal_len, uint32_t* roots_out_host,
                                       size_t roots_cap, size_t* nroots, uint32_t* query_pos_out_host,
                                       uint32_t* global_out_host);

/* ---- pipelined host witnesses ----
 * The reference keeps the GPU busy with depth-2 work queues (r0vm/src/actors/worker.rs:70-76,185-204): while segment s is
 * proved, segment s+1 is prepared. r0b200_witness_upload() enqueues the host->device copy of a witness (circuit 0 =
 * rv32im, 1 = recursion; column-major host matrices, pinned memory makes the copies asynchronous) on the context's
 * copy stream and returns at once; r0b200_prove_uploaded() consumes it (the handle's buffers pass to the proof; free
 * the handle afterwards). Uploading s+1 before proving s hides the PCIe transfer behind compute. accum_host may be
 * NULL (the normal case: accum does not exist before prove_begin has drawn the mix); the handle is then consumed by
 * r0b200_prove_begin(.., uploaded, ..) and accum goes to r0b200_prove_finish. */
r0b200_err r0b200_witness_upload(r0b200_ctx* ctx, int circuit, uint32_t po2, const uint32_t* code_host,
                                 const uint32_t* data_host, const uint32_t* accum_host, r0b200_witness** out);
void r0b200_witness_free(r0b200_witness* witness);
r0b200_err r0b200_prove_uploaded(r0b200_ctx* ctx, int hash, r0b200_witness* witness, const uint32_t* global_host,
                                 uint32_t* seal_out_host, size_t seal_cap, size_t* seal_len, uint32_t* roots_out_host,
                                 size_t roots_cap, size_t* nroots, uint32_t* query_pos_out_host);

#ifdef __cplusplus
}
#endif
#endif /* R0B200_H */
