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
