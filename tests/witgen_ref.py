"""TEST INFRASTRUCTURE: ctypes harness for
  * the REFERENCE's own CPU witness generator, compiled where it lies by `make -C oracle ref`
    (oracle/_ref/librv32im_witgen_ref.so: risc0_circuit_rv32im_cpu_witgen / _cpu_accum,
    rv32im-sys/kernels/cxx/ffi.cpp:275-365, called the way rv32im/src/prove/hal/cpu.rs:49-142 and
    prove/witgen/mod.rs:130-224 call them), and
  * the host build of this repo's GENERATED witgen code (tests/witgen_host_check.cpp), used to validate the generator
    and the circuit IR on the CPU.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_LIB = os.path.join(ROOT, "oracle", "_ref", "librv32im_witgen_ref.so")
HOST_LIB = os.path.join(ROOT, "tests", "_build", "libwitgen_hostcheck.so")
INVALID = 0xFFFFFFFF
N_DATA, N_ACCUM, N_GLOBAL, N_MIX = 211, 103, 90, 36


class RawBuffer(C.Structure):
    _fields_ = [("buf", C.c_void_p), ("rows", C.c_size_t), ("cols", C.c_size_t), ("checked", C.c_bool)]


class RawExecBuffers(C.Structure):
    _fields_ = [("global_", RawBuffer), ("data", RawBuffer)]


class RawAccumBuffers(C.Structure):
    _fields_ = [("data", RawBuffer), ("accum", RawBuffer), ("global_", RawBuffer), ("mix", RawBuffer)]


class RawPreflightTrace(C.Structure):
    _fields_ = [("cycles", C.c_void_p), ("txns", C.c_void_p), ("bigint_bytes", C.c_void_p), ("txns_len", C.c_uint32),
                ("bigint_bytes_len", C.c_uint32), ("table_split_cycle", C.c_uint32)]


def have_ref():
    return os.path.exists(REF_LIB)


_ref = None


def ref_lib():
    global _ref
    if _ref is None:
        _ref = C.CDLL(REF_LIB)
        _ref.risc0_circuit_rv32im_cpu_witgen.restype = C.c_char_p
        _ref.risc0_circuit_rv32im_cpu_accum.restype = C.c_char_p
    return _ref


def _buf(a, rows, cols, checked=True):
    return RawBuffer(a.ctypes.data, rows, cols, checked)


def _trace(pf, cycles, txns, bigint):
    return RawPreflightTrace(cycles.ctypes.data, txns.ctypes.data, bigint.ctypes.data, len(txns), len(bigint),
                             pf.table_split_cycle)


def _bigint_bytes(pf):
    b = np.ascontiguousarray(pf.bigint_bytes, dtype=np.uint8)
    return b.copy() if len(b) else np.zeros(1, dtype=np.uint8)


def _inject_bigint_accum(pf, accum, mix):
    """witgen/mod.rs:186-207: segments with bigint cycles get their BigIntAccumState cells (a function of the mix)
    scattered into accum before step_accum"""
    inj = pf.bigint_accum_injector(mix) if getattr(pf, "has_bigint", False) else None
    if inj is not None:
        scatter(accum, *inj)


def scatter(into, index, offsets, values):
    into[offsets] = values    # Hal::scatter (cpu.rs:598-615); offsets are unique per (row, col)
    return into


def zeroize(a):
    a[a == INVALID] = 0
    return a


def ref_generate_witness(pf):
    """WitnessGenerator::new (witgen/mod.rs:130-176) with the reference C++ step_Top. Returns (global, data), zeroized."""
    rows = pf.rows
    data = np.full(N_DATA * rows, INVALID, dtype=np.uint32)
    scatter(data, *pf.injector)
    glob = pf.global_.copy()
    cycles, txns, bigint = pf.cycles.copy(), pf.txns.copy(), _bigint_bytes(pf)
    bufs = RawExecBuffers(_buf(glob, 1, N_GLOBAL), _buf(data, rows, N_DATA))
    tr = _trace(pf, cycles, txns, bigint)
    err = ref_lib().risc0_circuit_rv32im_cpu_witgen(C.c_uint32(0), C.byref(bufs), C.byref(tr), C.c_uint32(rows))
    if err:
        raise RuntimeError("reference witgen: " + err.decode())
    return zeroize(glob), zeroize(data)


def ref_accum(pf, glob, data, mix):
    """WitnessGenerator::accum (witgen/mod.rs:178-224) with the reference C++ step_TopAccum + prefix sums"""
    accum = np.full(N_ACCUM * pf.rows, INVALID, dtype=np.uint32)
    _inject_bigint_accum(pf, accum, np.ascontiguousarray(mix, dtype=np.uint32))
    return ref_accum_prefilled(pf, glob, data, mix, accum)


def ref_accum_prefilled(pf, glob, data, mix, accum):
    """the reference step_accum on an accum matrix the caller has INVALID-filled and injected"""
    rows = pf.rows
    data, glob, mix = data.copy(), glob.copy(), np.ascontiguousarray(mix, dtype=np.uint32).copy()
    cycles, txns, bigint = pf.cycles.copy(), pf.txns.copy(), _bigint_bytes(pf)
    bufs = RawAccumBuffers(_buf(data, rows, N_DATA), _buf(accum, rows, N_ACCUM), _buf(glob, 1, N_GLOBAL), _buf(mix, 1, N_MIX))
    tr = _trace(pf, cycles, txns, bigint)
    err = ref_lib().risc0_circuit_rv32im_cpu_accum(C.byref(bufs), C.byref(tr), C.c_uint32(rows))
    if err:
        raise RuntimeError("reference accum: " + err.decode())
    return zeroize(accum)


# ---- host build of the generated code
def build_host_check(force=False):
    src = os.path.join(ROOT, "tests", "witgen_host_check.cpp")
    deps = [src, os.path.join(ROOT, "risc0_b200", "csrc", "witgen_rt.cuh"), os.path.join(ROOT, "risc0_b200", "csrc", "fp.cuh"),
            os.path.join(ROOT, "risc0_b200", "csrc", "gen", "witgen_rv32im.inc")]
    if not os.path.exists(deps[-1]):
        subprocess.check_call(["python", os.path.join(ROOT, "tools", "gen_witgen.py")])
    if force or not os.path.exists(HOST_LIB) or os.path.getmtime(HOST_LIB) < max(os.path.getmtime(d) for d in deps):
        os.makedirs(os.path.dirname(HOST_LIB), exist_ok=True)
        subprocess.check_call(["g++", "-std=c++17", "-O1", "-fPIC", "-fopenmp", "-shared", "-o", HOST_LIB, src])
    return HOST_LIB


_host = None


def host_lib():
    global _host
    if _host is None:
        _host = C.CDLL(build_host_check())
    return _host


ERR_NAMES = {1: "Inconsistent set", 2: "Read of unset value", 3: "eqz failure", 4: "txn cycle mismatch",
             5: "memory peek not in preflight", 6: "Invalid lookup table", 7: "u8/16 table error", 8: "unreachable mux arm"}


def _check_err(err, what):
    if err[0]:
        raise RuntimeError("%s: %d failures, first: %s at cycle %d (detail %d)" % (what, err[0], ERR_NAMES.get(int(err[1]), "?"),
                                                                                 err[2], err[3]))


def host_generate_witness(pf):
    rows = pf.rows
    data = np.full(N_DATA * rows, INVALID, dtype=np.uint32)
    scatter(data, *pf.injector)
    glob = pf.global_.copy()
    bigint = _bigint_bytes(pf)
    err = np.zeros(4, dtype=np.uint32)
    host_lib().wg_host_witgen(C.c_void_p(pf.cycles.ctypes.data), C.c_uint32(rows), C.c_void_p(pf.txns.ctypes.data),
                              C.c_uint32(len(pf.txns)), C.c_void_p(bigint.ctypes.data), C.c_uint32(len(pf.bigint_bytes)),
                              C.c_uint32(pf.table_split_cycle), C.c_void_p(glob.ctypes.data), C.c_void_p(data.ctypes.data),
                              C.c_uint32(rows), C.c_void_p(err.ctypes.data))
    _check_err(err, "generated witgen (host build)")
    return zeroize(glob), zeroize(data)


def host_accum(pf, glob, data, mix, accum=None):
    """generated step_TopAccum (host build); `accum`: a matrix the caller has INVALID-filled and injected itself"""
    rows = pf.rows
    data, glob, mix = data.copy(), glob.copy(), np.ascontiguousarray(mix, dtype=np.uint32).copy()
    if accum is None:
        accum = np.full(N_ACCUM * rows, INVALID, dtype=np.uint32)
        _inject_bigint_accum(pf, accum, mix)
    bigint = _bigint_bytes(pf)
    err = np.zeros(4, dtype=np.uint32)
    host_lib().wg_host_accum(C.c_void_p(pf.cycles.ctypes.data), C.c_uint32(rows), C.c_void_p(pf.txns.ctypes.data),
                             C.c_uint32(len(pf.txns)), C.c_void_p(bigint.ctypes.data), C.c_uint32(len(pf.bigint_bytes)),
                             C.c_void_p(data.ctypes.data), C.c_void_p(accum.ctypes.data), C.c_void_p(glob.ctypes.data),
                             C.c_void_p(mix.ctypes.data), C.c_uint32(rows), C.c_void_p(err.ctypes.data))
    _check_err(err, "generated accum (host build)")
    return zeroize(accum)
