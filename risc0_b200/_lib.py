"""ctypes loader for libr0b200.so. Fails loudly when the CUDA extension is missing (no fallback path exists)."""
import ctypes as C
import os
import re

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("R0B200_LIB") or os.path.join(HERE, "lib", "libr0b200.so")
HEADER = os.path.join(HERE, "..", "include", "r0b200.h")


class R0B200Error(RuntimeError):
    pass


_lib = None


def lib_available():
    return os.path.exists(LIB_PATH)


def declared_symbols():
    """every function name declared in include/r0b200.h"""
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(r0b200_[a-z0-9_]+)\s*\(", text)))


COMPAT_HEADER = os.path.join(HERE, "..", "include", "r0b200_compat.h")


def compat_symbols():
    """every function name declared in include/r0b200_compat.h (the reference's own FFI symbol table)"""
    text = open(COMPAT_HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b((?:risc0_zkp_cuda|risc0_circuit|sppark|supra)_[A-Za-z0-9_]+)\s*\(", text)))


def load_library():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise R0B200Error(
                "libr0b200.so is not built (%s). Run `python -m risc0_b200.build`; this backend has no CPU fallback."
                % LIB_PATH)
        _lib = C.CDLL(LIB_PATH)
        for name in declared_symbols():
            fn = getattr(_lib, name)  # AttributeError if the library does not export a declared symbol
            fn.restype = C.c_void_p
        _lib.r0b200_destroy.restype = None
        _lib.r0b200_free_error.restype = None
        _lib.r0b200_witness_free.restype = None
        _lib.r0b200_prove_abort.restype = None
        _lib.r0b200_trace_free.restype = None
        _lib.r0b200_segment_free.restype = None
        _lib.r0b200_launch_count.restype = C.c_uint64
        _lib.r0b200_bytes_peak.restype = C.c_uint64
        _lib.r0b200_stream.restype = C.c_void_p
    return _lib


def check(err):
    if err:
        msg = C.cast(err, C.c_char_p).value.decode(errors="replace")
        load_library().r0b200_free_error(C.c_void_p(err))
        raise R0B200Error(msg)
