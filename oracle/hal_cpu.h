// ORACLE — TEST INFRASTRUCTURE ONLY. CPU restatement of the reference `CpuHal` ops; never linked into the product.
//
// Each function follows the `impl Hal for CpuHal` method of the same name in
// /root/reference/risc0/zkp/src/hal/cpu.rs:305-651 and the defaulted trait methods `combos_prepare` /
// `combos_divide` in hal/mod.rs:202-257. Buffers are plain arrays with the reference's layouts
// (column-major matrices `buf[col*rows + row]`, Montgomery words). OpenMP replaces rayon at the same loops.
#pragma once
#include <stdexcept>
#include <vector>

#include "field.h"
#include "hashes.h"
#include "ntt.h"

namespace oracle {

constexpr size_t INV_RATE = 4, FRI_FOLD = 16, FRI_MIN_DEGREE = 256, QUERIES = 50, EXT_SIZE = 4;
constexpr size_t CHECK_SIZE = INV_RATE * EXT_SIZE;

// cpu.rs:305-340
inline void batch_expand_into_evaluate_ntt(Fp* out, size_t out_size, const Fp* in, size_t in_size, size_t count,
                                           unsigned /*expand_bits_arg*/) {
  size_t out_row = out_size / count, in_row = in_size / count;
  unsigned expand_bits = log2_ceil(out_row / in_row);
  if (out_row != (in_row << expand_bits)) throw std::runtime_error("expand size mismatch");
#pragma omp parallel for schedule(dynamic)
  for (size_t c = 0; c < count; c++) {
    expand(out + c * out_row, in + c * in_row, in_row, expand_bits);
    evaluate_ntt(out + c * out_row, out_row, expand_bits);
  }
}

// cpu.rs:342-350
inline void batch_interpolate_ntt(Fp* io, size_t size, size_t count) {
  size_t row = size / count;
#pragma omp parallel for schedule(dynamic)
  for (size_t c = 0; c < count; c++) interpolate_ntt(io + c * row, row);
}

// cpu.rs:352-360
inline void batch_bit_reverse(Fp* io, size_t size, size_t count) {
  size_t row = size / count;
#pragma omp parallel for schedule(dynamic)
  for (size_t c = 0; c < count; c++) bit_reverse(io + c * row, row);
}

// cpu.rs:362-393
inline void batch_evaluate_any(const Fp* coeffs, size_t coeffs_size, size_t poly_count, const uint32_t* which,
                               const FpExt* xs, FpExt* out, size_t eval_count) {
  size_t n = coeffs_size / poly_count;
#pragma omp parallel for schedule(dynamic)
  for (size_t e = 0; e < eval_count; e++) {
    FpExt tot, cur = FpExt::one();
    const Fp* local = coeffs + n * which[e];
    for (size_t i = 0; i < n; i++) {
      tot += cur * local[i];
      cur *= xs[e];
    }
    out[e] = tot;
  }
}

// cpu.rs:395-408
inline void zk_shift(Fp* io, size_t size, size_t poly_count) {
  unsigned bits = log2_ceil(size / poly_count);
  const Fp three(3);
#pragma omp parallel for
  for (size_t idx = 0; idx < size; idx++) {
    size_t pos = idx & ((size_t(1) << bits) - 1);
    uint32_t rev = brev(uint32_t(pos), bits);
    io[idx] *= three.pow(rev);
  }
}

// cpu.rs:410-455
inline void mix_poly_coeffs(FpExt* out, size_t out_size, FpExt mix_start, FpExt mix, const Fp* in,
                            const uint32_t* combos, size_t input_size, size_t count) {
  std::vector<FpExt> mix_pows(input_size);
  FpExt cur = mix_start;
  for (size_t i = 0; i < input_size; i++) {
    mix_pows[i] = cur;
    cur *= mix;
  }
  size_t chunks = out_size / count;
#pragma omp parallel for schedule(dynamic)
  for (size_t id = 0; id < chunks; id++) {
    FpExt* oc = out + id * count;
    for (size_t i = 0; i < input_size; i++) {
      if (combos[i] != id) continue;
      for (size_t idx = 0; idx < count; idx++) oc[idx] += mix_pows[i] * in[count * i + idx];
    }
  }
}

// cpu.rs:457-473
inline void eltwise_add_elem(Fp* out, const Fp* a, const Fp* b, size_t n) {
  for (size_t i = 0; i < n; i++) out[i] = a[i] + b[i];
}

// cpu.rs:475-500 : input is (to_add x count) AoS FpExt; output is 4 SoA planes of `count`
inline void eltwise_sum_extelem(Fp* out, size_t out_size, const FpExt* in, size_t in_size) {
  size_t count = out_size / EXT_SIZE;
  size_t to_add = in_size / count;
#pragma omp parallel for
  for (size_t idx = 0; idx < count; idx++) {
    FpExt sum;
    for (size_t t = 0; t < to_add; t++) sum += in[t * count + idx];
    for (size_t k = 0; k < EXT_SIZE; k++) out[k * count + idx] = sum.e[k];
  }
}

// cpu.rs:502-516
inline void eltwise_copy_elem(Fp* out, const Fp* in, size_t n) {
  for (size_t i = 0; i < n; i++) out[i] = in[i];
}

// cpu.rs:518-522
inline void eltwise_zeroize_elem(Fp* io, size_t n) {
  for (size_t i = 0; i < n; i++) io[i] = io[i].valid_or_zero();
}

// cpu.rs:524-553
inline void fri_fold(Fp* out, size_t out_size, const Fp* in, FpExt mix) {
  size_t count = out_size / EXT_SIZE;
#pragma omp parallel for
  for (size_t idx = 0; idx < count; idx++) {
    FpExt tot, cur_mix = FpExt::one();
    for (size_t i = 0; i < FRI_FOLD; i++) {
      size_t rev_i = brev(uint32_t(i), 4);
      size_t rev_idx = rev_i * count + idx;
      FpExt factor(in[0 * count * FRI_FOLD + rev_idx], in[1 * count * FRI_FOLD + rev_idx],
                   in[2 * count * FRI_FOLD + rev_idx], in[3 * count * FRI_FOLD + rev_idx]);
      tot += cur_mix * factor;
      cur_mix *= mix;
    }
    for (size_t k = 0; k < EXT_SIZE; k++) out[count * k + idx] = tot.e[k];
  }
}

// cpu.rs:555-567
inline void hash_rows(const HashSuite& suite, Digest* out, size_t row_size, const Fp* matrix, size_t matrix_size) {
  size_t col_size = matrix_size / row_size;
#pragma omp parallel for
  for (size_t idx = 0; idx < row_size; idx++) out[idx] = suite.hash_elem_slice(matrix + idx, col_size, row_size);
}

// cpu.rs:569-581
inline void hash_fold(const HashSuite& suite, Digest* io, size_t input_size, size_t output_size) {
  if (input_size != 2 * output_size) throw std::runtime_error("hash_fold size mismatch");
  Digest* output = io + output_size;
  const Digest* input = io + input_size;
#pragma omp parallel for
  for (size_t idx = 0; idx < output_size; idx++) output[idx] = suite.hash_pair(input[2 * idx], input[2 * idx + 1]);
}

// cpu.rs:583-596
inline void gather_sample(Fp* dst, const Fp* src, size_t idx, size_t size, size_t stride) {
  for (size_t g = 0; g < size; g++) dst[g] = src[g * stride + idx];
}

// cpu.rs:598-615
inline void scatter(Fp* into, const uint32_t* index, size_t index_len, const uint32_t* offsets, const Fp* values) {
  if (index_len == 0) return;
  for (size_t cycle = 0; cycle + 1 < index_len; cycle++)
    for (uint32_t idx = index[cycle]; idx < index[cycle + 1]; idx++) into[offsets[idx]] = values[idx];
}

// cpu.rs:617-635
inline void eltwise_copy_elem_slice(Fp* into, const Fp* from, size_t from_rows, size_t from_cols, size_t from_offset,
                                    size_t from_stride, size_t into_offset, size_t into_stride) {
  for (size_t row = 0; row < from_rows; row++)
    for (size_t col = 0; col < from_cols; col++)
      into[into_offset + row * into_stride + col] = from[from_offset + row * from_stride + col];
}

// cpu.rs:637-642
inline void prefix_products(FpExt* io, size_t n) {
  for (size_t i = 1; i < n; i++) io[i] = io[i] * io[i - 1];
}

// hal/mod.rs:202-234
inline void combos_prepare(FpExt* combos, const FpExt* coeff_u, size_t combo_count, size_t cycles,
                           const uint32_t* reg_sizes, const uint32_t* reg_combo_ids, size_t nregs, FpExt mix) {
  size_t cur_pos = 0;
  FpExt cur = FpExt::one();
  for (size_t r = 0; r < nregs; r++) {
    for (size_t i = 0; i < reg_sizes[r]; i++) combos[cycles * reg_combo_ids[r] + i] -= cur * coeff_u[cur_pos + i];
    cur *= mix;
    cur_pos += reg_sizes[r];
  }
  for (size_t i = 0; i < CHECK_SIZE; i++) {
    combos[cycles * combo_count] -= cur * coeff_u[cur_pos];
    cur_pos++;
    cur *= mix;
  }
}

// hal/mod.rs:236-257 ; chunk i is divided by every pow in pows[i]; returns false if any remainder != 0
inline bool combos_divide(FpExt* combos, const std::vector<std::vector<FpExt>>& pows, size_t cycles) {
  bool ok = true;
#pragma omp parallel for schedule(dynamic)
  for (size_t i = 0; i < pows.size(); i++) {
    for (const FpExt& z : pows[i]) {
      FpExt rem = poly_divide(combos + i * cycles, cycles, z);
      if (rem != FpExt()) ok = false;
    }
  }
  return ok;
}

}  // namespace oracle
