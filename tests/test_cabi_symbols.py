"""CPU-only: the C-ABI library is built, loads, and exports every symbol include/r0b200.h declares; the product
refuses to run without a GPU instead of falling back."""
import pytest

import risc0_b200
from risc0_b200 import _lib


def test_library_exports_every_declared_symbol():
    names = _lib.declared_symbols()
    assert len(names) >= 30
    lib = _lib.load_library()
    for n in names:
        assert hasattr(lib, n), n


def test_library_exports_the_reference_symbol_table():
    """include/r0b200_compat.h: the names risc0/sys/src/cuda.rs:19-80, risc0/sys/kernels/zkp/cuda/ffi.cu:25-145 and
    the circuit -sys crates declare, so the unmodified hal/cuda.rs can link against libr0b200.so"""
    names = _lib.compat_symbols()
    want = {"risc0_zkp_cuda_eltwise_add_fp", "risc0_zkp_cuda_eltwise_mul_factor_fp", "risc0_zkp_cuda_eltwise_copy_fp",
            "risc0_zkp_cuda_eltwise_copy_fp_region", "risc0_zkp_cuda_eltwise_sum_fpext", "risc0_zkp_cuda_eltwise_zeroize_fp",
            "risc0_zkp_cuda_eltwise_zeroize_fpext", "risc0_zkp_cuda_fri_fold", "risc0_zkp_cuda_mix_poly_coeffs",
            "risc0_zkp_cuda_batch_bit_reverse", "risc0_zkp_cuda_batch_evaluate_any", "risc0_zkp_cuda_gather_sample",
            "risc0_zkp_cuda_scatter", "risc0_zkp_cuda_sha_rows", "risc0_zkp_cuda_sha_fold", "risc0_zkp_cuda_combos_prepare",
            "sppark_init", "sppark_batch_expand", "sppark_batch_NTT", "sppark_batch_iNTT", "sppark_batch_zk_shift",
            "sppark_poseidon2_fold", "sppark_poseidon2_rows", "sppark_poseidon254_fold", "sppark_poseidon254_rows",
            "supra_poly_divide", "risc0_circuit_rv32im_cuda_eval_check", "risc0_circuit_recursion_cuda_eval_check",
            "risc0_circuit_rv32im_cuda_witgen", "risc0_circuit_rv32im_cuda_accum"}
    assert want <= set(names), want - set(names)
    lib = _lib.load_library()
    for n in names:
        assert hasattr(lib, n), n


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(risc0_b200.R0B200Error) as ei:
        risc0_b200.B200Hal()
    assert "no CPU fallback" in str(ei.value)


def test_product_does_not_reference_oracle():
    """the shipped package must not import, link or call anything under oracle/"""
    import os
    root = os.path.dirname(os.path.abspath(risc0_b200.__file__))
    for dirpath, _, files in os.walk(root):
        if os.path.basename(dirpath) in ("build", "lib", "__pycache__"):
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert "oracle_lib" not in text and "liboracle" not in text and "oracle/" not in text, f
