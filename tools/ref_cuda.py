"""ctypes binding of oracle/_ref/libref_zkp_cuda.so: the REFERENCE'S OWN CUDA kernels (risc0/sys/kernels/zkp/cuda/
{ffi,kernels,eltwise,combos,sha}.cu, the subset that does not need sppark) compiled for sm_100a by `make -C oracle
refcuda`. Test / benchmark infrastructure only: a reported baseline and a second parity oracle, never part of the product.
Entry points and argument order: risc0/sys/kernels/zkp/cuda/ffi.cu:25-145; launch geometry as hal/cuda.rs passes it."""
import ctypes as C
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "oracle", "_ref", "libref_zkp_cuda.so")
_lib = None


def available():
    return os.path.exists(LIB)


def use(path):
    """bind the harness to another library exporting the same symbol table (libr0b200.so's compat layer, so that
    the same calls can be run against the reference's kernels and against ours)"""
    global _lib
    _lib = None
    lib(path)


def lib(path=None):
    global _lib
    if _lib is None:
        _lib = C.CDLL(path or LIB)
        for n in ("risc0_zkp_cuda_batch_bit_reverse", "risc0_zkp_cuda_fri_fold", "risc0_zkp_cuda_mix_poly_coeffs",
                  "risc0_zkp_cuda_batch_evaluate_any", "risc0_zkp_cuda_eltwise_sum_fpext", "risc0_zkp_cuda_eltwise_copy_fp",
                  "risc0_zkp_cuda_sha_rows", "risc0_zkp_cuda_sha_fold", "risc0_zkp_cuda_gather_sample",
                  "risc0_zkp_cuda_eltwise_zeroize_fp", "risc0_zkp_cuda_eltwise_add_fp"):
            getattr(_lib, n).restype = C.c_char_p
    return _lib


def _ok(err):
    if err:
        raise RuntimeError(err.decode())


u32 = C.c_uint32


def batch_bit_reverse(io, n_bits, total):            # hal/cuda.rs:594-613 passes io.size() as count
    _ok(lib().risc0_zkp_cuda_batch_bit_reverse(io.ptr, u32(n_bits), u32(total)))


def fri_fold(out, inp, mix_dev, count):
    _ok(lib().risc0_zkp_cuda_fri_fold(out.ptr, inp.ptr, mix_dev.ptr, u32(count)))


def mix_poly_coeffs(out, inp, combos_dev, mix_start_dev, mix_dev, input_size, count):
    _ok(lib().risc0_zkp_cuda_mix_poly_coeffs(out.ptr, inp.ptr, combos_dev.ptr, mix_start_dev.ptr, mix_dev.ptr,
                                             u32(input_size), u32(count)))


def batch_evaluate_any(out, coeffs, which_dev, xs_dev, evals, deg):   # hal/cuda.rs:615-660
    _ok(lib().risc0_zkp_cuda_batch_evaluate_any(out.ptr, coeffs.ptr, which_dev.ptr, xs_dev.ptr, u32(256 * 16),
                                                u32(evals * 256), u32(deg)))


def eltwise_sum_fpext(out, inp, to_add, count):
    _ok(lib().risc0_zkp_cuda_eltwise_sum_fpext(out.ptr, inp.ptr, u32(to_add), u32(count)))


def eltwise_copy_fp(out, inp, count):
    _ok(lib().risc0_zkp_cuda_eltwise_copy_fp(out.ptr, inp.ptr, u32(count)))


def sha_rows(out, matrix, rows, cols):
    _ok(lib().risc0_zkp_cuda_sha_rows(out.ptr, matrix.ptr, u32(rows), u32(cols)))


def sha_fold(out_ptr, in_ptr, count):
    _ok(lib().risc0_zkp_cuda_sha_fold(out_ptr, in_ptr, u32(count)))


def gather_sample(dst, src, idx, size, stride):
    _ok(lib().risc0_zkp_cuda_gather_sample(dst.ptr, src.ptr, u32(idx), u32(size), u32(stride)))
