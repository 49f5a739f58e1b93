"""ctypes binding of oracle/liboracle.so — TEST INFRASTRUCTURE (checker only).

The oracle is the CPU restatement of the reference (`oracle/*.h`, each citing the reference file:line it follows) plus
`oracle/_ref/librv32im_poly_fp_ref.so`, the reference's own generated C++ compiled as-is. Only tests, smoke() and
bench.py's cpu_baseline / --impl reference legs import this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
LIB = os.path.join(ORACLE_DIR, "liboracle.so")
REF_LIB = os.path.join(ORACLE_DIR, "_ref", "librv32im_poly_fp_ref.so")
REF_LIB_RECURSION = os.path.join(ORACLE_DIR, "_ref", "librecursion_poly_fp_ref.so")

P = 15 * 2**27 + 1
POSEIDON2, SHA256 = 0, 1

_u32p = C.POINTER(C.c_uint32)
_u64 = C.c_uint64


def build(force=False):
    srcs = [os.path.join(ORACLE_DIR, f) for f in os.listdir(ORACLE_DIR) if f.endswith((".cpp", ".h"))]
    srcs += [os.path.join(ORACLE_DIR, "tables", f) for f in os.listdir(os.path.join(ORACLE_DIR, "tables"))]
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < max(os.path.getmtime(s) for s in srcs):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "oracle"], stdout=subprocess.DEVNULL)
    ref_witgen = os.path.join(ORACLE_DIR, "_ref", "librv32im_witgen_ref.so")
    if not (os.path.exists(REF_LIB) and os.path.exists(REF_LIB_RECURSION) and os.path.exists(ref_witgen)) and \
            os.path.isdir("/root/reference"):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "-j5", "ref"], stdout=subprocess.DEVNULL)


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(LIB)
        for name in ("orc_load_ref", "orc_batch_expand_into_evaluate_ntt", "orc_hash_fold", "orc_combos_divide",
                     "orc_prove_rv32im", "orc_prove_hello", "orc_verify_hello", "orc_verify_rv32im",
                     "orc_rv32im_eval_check", "orc_load_ref_recursion", "orc_recursion_eval_check",
                     "orc_prove_recursion", "orc_verify_recursion", "orc_prove_rv32im_mix", "orc_prove_rv32im_cb",
                     "orc_verify_rv32im_ext", "orc_verify_recursion_ext", "orc_rv32im_check_constraints"):
            getattr(_lib, name).restype = C.c_void_p
        for name in ("orc_fp_encode", "orc_fp_decode", "orc_fp_add", "orc_fp_sub", "orc_fp_mul", "orc_fp_pow",
                     "orc_fp_inv", "orc_rou_fwd", "orc_rou_rev", "orc_rng_elem", "orc_rng_bits"):
            getattr(_lib, name).restype = C.c_uint32
        _lib.orc_rng_new.restype = C.c_void_p
        _lib.orc_fp_pow.argtypes = [C.c_uint32, _u64]
    return _lib


def _check(err):
    if err:
        msg = C.cast(err, C.c_char_p).value.decode()
        lib().orc_free_str(C.c_void_p(err))
        raise RuntimeError(msg)


def have_ref():
    return os.path.exists(REF_LIB)


def load_ref():
    if not lib().orc_ref_loaded():
        _check(lib().orc_load_ref(REF_LIB.encode()))


def have_ref_recursion():
    return os.path.exists(REF_LIB_RECURSION)


def load_ref_recursion():
    if not lib().orc_ref_recursion_loaded():
        _check(lib().orc_load_ref_recursion(REF_LIB_RECURSION.encode()))


def ptr(a):
    assert a.dtype == np.uint32 and a.flags["C_CONTIGUOUS"], (a.dtype, a.flags)
    return a.ctypes.data_as(_u32p)


def u32(a):
    return np.ascontiguousarray(a, dtype=np.uint32)


# ---------------------------------------------------------------- field helpers (numpy, vectorised)
R_MOD_P = 2**32 % P


def encode(x):
    """normal form -> Montgomery form (numpy uint32 array or int)"""
    return (np.asarray(x, dtype=np.uint64) % P * R_MOD_P % P).astype(np.uint32)


RINV = pow(R_MOD_P, -1, P)


def decode(x):
    return (np.asarray(x, dtype=np.uint64) * RINV % P).astype(np.uint32)


def rand_elems(rng, n):
    """n uniform field elements in Montgomery form"""
    return encode(rng.integers(0, P, size=n, dtype=np.uint64))


def rand_ext(rng):
    return rand_elems(rng, 4)


# ---------------------------------------------------------------- thin wrappers (names = Hal trait methods)
def poseidon2_mix(cells_mont):
    c = u32(cells_mont).copy()
    lib().orc_poseidon2_mix(ptr(c))
    return c


def hash_elems(kind, data, stride=1, n=None):
    data = u32(data)
    if n is None:
        n = len(data) // stride if stride > 1 else len(data)
    out = np.zeros(8, dtype=np.uint32)
    lib().orc_hash_elems(kind, ptr(out), ptr(data) if len(data) else None, _u64(n), _u64(stride))
    return out


def hash_pair(kind, a, b):
    out = np.zeros(8, dtype=np.uint32)
    lib().orc_hash_pair(kind, ptr(out), ptr(u32(a)), ptr(u32(b)))
    return out


class Rng:
    def __init__(self, kind=POSEIDON2):
        self.h = C.c_void_p(lib().orc_rng_new(kind))

    def mix(self, d):
        lib().orc_rng_mix(self.h, ptr(u32(d)))

    def elem(self):
        return lib().orc_rng_elem(self.h)

    def bits(self, b):
        return lib().orc_rng_bits(self.h, b)

    def ext(self):
        return np.array([self.elem() for _ in range(4)], dtype=np.uint32)

    def __del__(self):
        try:
            lib().orc_rng_free(self.h)
        except Exception:
            pass


def batch_expand_into_evaluate_ntt(inp, count, expand_bits=2):
    inp = u32(inp)
    out = np.zeros(len(inp) << expand_bits, dtype=np.uint32)
    _check(lib().orc_batch_expand_into_evaluate_ntt(ptr(out), _u64(len(out)), ptr(inp), _u64(len(inp)), _u64(count),
                                                    C.c_uint32(expand_bits)))
    return out


def batch_interpolate_ntt(io, count):
    io = u32(io).copy()
    lib().orc_batch_interpolate_ntt(ptr(io), _u64(len(io)), _u64(count))
    return io


def batch_bit_reverse(io, count):
    io = u32(io).copy()
    lib().orc_batch_bit_reverse(ptr(io), _u64(len(io)), _u64(count))
    return io


def zk_shift(io, count):
    io = u32(io).copy()
    lib().orc_zk_shift(ptr(io), _u64(len(io)), _u64(count))
    return io


def batch_evaluate_any(coeffs, poly_count, which, xs):
    coeffs, which, xs = u32(coeffs), u32(which), u32(xs)
    out = np.zeros(4 * len(which), dtype=np.uint32)
    lib().orc_batch_evaluate_any(ptr(coeffs), _u64(len(coeffs)), _u64(poly_count), ptr(which), ptr(xs), ptr(out),
                                 _u64(len(which)))
    return out


def mix_poly_coeffs(out, mix_start, mix, inp, combos, input_size, count):
    out = u32(out).copy()
    lib().orc_mix_poly_coeffs(ptr(out), _u64(len(out) // 4), ptr(u32(mix_start)), ptr(u32(mix)), ptr(u32(inp)),
                              ptr(u32(combos)), _u64(input_size), _u64(count))
    return out


def eltwise_add_elem(a, b):
    a, b = u32(a), u32(b)
    out = np.zeros_like(a)
    lib().orc_eltwise_add_elem(ptr(out), ptr(a), ptr(b), _u64(len(a)))
    return out


def eltwise_sum_extelem(inp, count):
    inp = u32(inp)
    out = np.zeros(4 * count, dtype=np.uint32)
    lib().orc_eltwise_sum_extelem(ptr(out), _u64(len(out)), ptr(inp), _u64(len(inp) // 4))
    return out


def eltwise_zeroize_elem(io):
    io = u32(io).copy()
    lib().orc_eltwise_zeroize_elem(ptr(io), _u64(len(io)))
    return io


def fri_fold(inp, mix):
    inp = u32(inp)
    out = np.zeros(len(inp) // 16, dtype=np.uint32)
    lib().orc_fri_fold(ptr(out), _u64(len(out)), ptr(inp), ptr(u32(mix)))
    return out


def hash_rows(kind, matrix, rows):
    matrix = u32(matrix)
    out = np.zeros(8 * rows, dtype=np.uint32)
    lib().orc_hash_rows(kind, ptr(out), _u64(rows), ptr(matrix), _u64(len(matrix)))
    return out


def hash_fold(kind, io, input_size, output_size):
    io = u32(io).copy()
    _check(lib().orc_hash_fold(kind, ptr(io), _u64(input_size), _u64(output_size)))
    return io


def merkle_tree(kind, matrix, rows):
    """nodes heap (2*rows digests) as MerkleTreeProver::new builds it"""
    nodes = np.zeros(16 * rows, dtype=np.uint32)
    nodes[8 * rows:] = hash_rows(kind, matrix, rows)
    size = rows
    while size > 1:
        nodes = hash_fold(kind, nodes, size, size // 2)
        size //= 2
    return nodes


def gather_sample(src, idx, size, stride):
    src = u32(src)
    out = np.zeros(size, dtype=np.uint32)
    lib().orc_gather_sample(ptr(out), ptr(src), _u64(idx), _u64(size), _u64(stride))
    return out


def scatter(into, index, offsets, values):
    into = u32(into).copy()
    lib().orc_scatter(ptr(into), ptr(u32(index)), _u64(len(index)), ptr(u32(offsets)), ptr(u32(values)))
    return into


def eltwise_copy_elem_slice(into, frm, from_rows, from_cols, from_offset, from_stride, into_offset, into_stride):
    into = u32(into).copy()
    lib().orc_eltwise_copy_elem_slice(ptr(into), ptr(u32(frm)), _u64(from_rows), _u64(from_cols), _u64(from_offset),
                                      _u64(from_stride), _u64(into_offset), _u64(into_stride))
    return into


def prefix_products(io):
    io = u32(io).copy()
    lib().orc_prefix_products(ptr(io), _u64(len(io) // 4))
    return io


def combos_prepare(combos, coeff_u, combo_count, cycles, reg_sizes, reg_combo_ids, mix):
    combos = u32(combos).copy()
    lib().orc_combos_prepare(ptr(combos), ptr(u32(coeff_u)), _u64(combo_count), _u64(cycles), ptr(u32(reg_sizes)),
                             ptr(u32(reg_combo_ids)), _u64(len(reg_sizes)), ptr(u32(mix)))
    return combos


def combos_divide(combos, pow_begin, pows, cycles):
    combos = u32(combos).copy()
    _check(lib().orc_combos_divide(ptr(combos), _u64(len(pow_begin) - 1), ptr(u32(pow_begin)), ptr(u32(pows)),
                                   _u64(cycles)))
    return combos


def poly_divide_unchecked(poly, z):
    """single synthetic division; returns (quotient array, ok flag)"""
    poly = u32(poly).copy()
    err = lib().orc_combos_divide(ptr(poly), _u64(1), ptr(u32([0, 1])), ptr(u32(z)), _u64(len(poly) // 4))
    ok = not err
    if err:
        lib().orc_free_str(C.c_void_p(err))
    return poly, ok


def poly_divide(poly, z):
    """core/poly.rs:81-89: returns (quotient in place of the input, remainder)"""
    poly = u32(poly).copy()
    rem = np.zeros(4, dtype=np.uint32)
    lib().orc_poly_divide(ptr(poly), _u64(len(poly) // 4), ptr(u32(z)), ptr(rem))
    return poly, rem


def rv32im_poly_mix_pows(poly_mix):
    out = np.zeros(4 * 458, dtype=np.uint32)
    lib().orc_rv32im_poly_mix_pows.restype = C.c_uint32
    n = lib().orc_rv32im_poly_mix_pows(ptr(u32(poly_mix)), ptr(out))
    return out[:4 * n].copy()


def rou_fwd(k):
    lib().orc_rou_fwd.restype = C.c_uint32
    return int(lib().orc_rou_fwd(C.c_uint32(k)))


def rv32im_eval_check(accum, data, mix, out, poly_mix, po2, begin=0, end=None):
    load_ref()
    domain = 4 << po2
    if end is None:
        end = domain
    check = np.zeros(4 * domain, dtype=np.uint32)
    _check(lib().orc_rv32im_eval_check(ptr(check), ptr(u32(accum)), ptr(u32(data)), ptr(u32(mix)), ptr(u32(out)),
                                       ptr(u32(poly_mix)), C.c_uint32(po2), _u64(begin), _u64(end)))
    return check


def prove_rv32im(po2, code, data, accum, glob, kind=POSEIDON2):
    load_ref()
    cap = 1 << 20
    seal = np.zeros(cap, dtype=np.uint32)
    roots = np.zeros(8 * 16, dtype=np.uint32)
    qpos = np.zeros(50, dtype=np.uint32)
    n, nr = _u64(0), _u64(0)
    _check(lib().orc_prove_rv32im(kind, C.c_uint32(po2), ptr(u32(code)), ptr(u32(data)), ptr(u32(accum)),
                                  ptr(u32(glob)), ptr(seal), _u64(cap), C.byref(n), ptr(roots), _u64(16), C.byref(nr),
                                  ptr(qpos)))
    return seal[:n.value].copy(), roots[:8 * nr.value].reshape(-1, 8).copy(), qpos


def rv32im_check_constraints(accum, data, mix, out, poly_mix, po2):
    """(number of trace rows on which the rv32im constraint polynomial is non-zero, first such row or None)"""
    load_ref()
    bad, first = C.c_uint64(0), C.c_uint64(0)
    _check(lib().orc_rv32im_check_constraints(ptr(u32(accum)), ptr(u32(data)), ptr(u32(mix)), ptr(u32(out)),
                                              ptr(u32(poly_mix)), C.c_uint32(po2), C.byref(bad), C.byref(first)))
    return bad.value, (first.value if bad.value else None)


def prove_rv32im_mix(po2, code, data, glob, kind=POSEIDON2):
    """the accum mix (36 words) the transcript yields after the code and data commits
    (rv32im/src/prove/hal/mod.rs:209-213)"""
    mix = np.zeros(36, dtype=np.uint32)
    _check(lib().orc_prove_rv32im_mix(C.c_int(kind), C.c_uint32(po2), ptr(u32(code)), ptr(u32(data)), ptr(u32(glob)),
                                      ptr(mix)))
    return mix


def prove_rv32im_two_phase(po2, code, data, glob, accum_fn, kind=POSEIDON2):
    """CPU prove_core in the protocol order: accum_fn(mix) -> accum matrix is called after code and data have been
    committed and the mix drawn (rv32im/src/prove/hal/mod.rs:209-217). Returns (seal, roots, query positions)."""
    load_ref()
    n = 1 << po2
    u32p = C.POINTER(C.c_uint32)
    proto = C.CFUNCTYPE(None, u32p, u32p)

    def cb(mix_ptr, accum_ptr):
        mix = np.ctypeslib.as_array(mix_ptr, shape=(36,)).copy()
        out = np.ctypeslib.as_array(accum_ptr, shape=(103 * n,))
        out[:] = u32(accum_fn(mix))

    seal = np.zeros(1 << 20, dtype=np.uint32)
    roots = np.zeros(8 * 16, dtype=np.uint32)
    qpos = np.zeros(50, dtype=np.uint32)
    n_, nr = _u64(0), _u64(0)
    keep = proto(cb)
    _check(lib().orc_prove_rv32im_cb(C.c_int(kind), C.c_uint32(po2), ptr(u32(code)), ptr(u32(data)), ptr(u32(glob)), keep,
                                     ptr(seal), _u64(seal.size), C.byref(n_), ptr(roots), _u64(16), C.byref(nr), ptr(qpos)))
    return seal[:n_.value].copy(), roots[:8 * nr.value].reshape(-1, 8).copy(), qpos


def prove_hello(po2, accum, code, data, kind=POSEIDON2):
    cap = 1 << 20
    seal = np.zeros(cap, dtype=np.uint32)
    n = _u64(0)
    _check(lib().orc_prove_hello(kind, C.c_uint32(po2), ptr(u32(accum)), ptr(u32(code)), ptr(u32(data)), ptr(seal),
                                 _u64(cap), C.byref(n)))
    return seal[:n.value].copy()


def verify_hello(seal, po2, kind=POSEIDON2):
    seal = u32(seal)
    roots = np.zeros(8 * 16, dtype=np.uint32)
    nr = _u64(0)
    _check(lib().orc_verify_hello(kind, ptr(seal), _u64(len(seal)), C.c_uint32(po2), ptr(roots), C.byref(nr)))
    return roots[:8 * nr.value].reshape(-1, 8).copy()


def verify_rv32im(seal, kind=POSEIDON2):
    seal = u32(seal)
    roots = np.zeros(8 * 16, dtype=np.uint32)
    nr = _u64(0)
    _check(lib().orc_verify_rv32im(kind, ptr(seal), _u64(len(seal)), ptr(roots), C.byref(nr)))
    return roots[:8 * nr.value].reshape(-1, 8).copy()


def synthetic_witness(po2, seed=None):
    """SURVEY §8(d) synthetic segment: every cell uniform in [0,P) (Montgomery), code column all zero."""
    seed = 0x5EED0000 + po2 if seed is None else seed
    rng = np.random.Generator(np.random.PCG64(seed))
    n = 1 << po2
    code = np.zeros(n, dtype=np.uint32)
    data = rand_elems(rng, 211 * n)
    accum = rand_elems(rng, 103 * n)
    glob = rand_elems(rng, 90)
    return code, data, accum, glob


# ---------------------------------------------------------------- recursion circuit (ctrl 23, data 128, accum 12 columns)
REC_COLS = dict(accum=12, ctrl=23, data=128)


def recursion_eval_check(ctrl, data, accum, mix, glob, poly_mix, po2, begin=0, end=None):
    load_ref_recursion()
    domain = 4 << po2
    if end is None:
        end = domain
    check = np.zeros(4 * domain, dtype=np.uint32)
    _check(lib().orc_recursion_eval_check(ptr(check), ptr(u32(ctrl)), ptr(u32(data)), ptr(u32(accum)), ptr(u32(mix)),
                                          ptr(u32(glob)), ptr(u32(poly_mix)), C.c_uint32(po2), _u64(begin), _u64(end)))
    return check


def prove_recursion(po2, ctrl, data, accum, glob, kind=POSEIDON2):
    load_ref_recursion()
    cap = 1 << 20
    seal = np.zeros(cap, dtype=np.uint32)
    roots = np.zeros(8 * 16, dtype=np.uint32)
    qpos = np.zeros(50, dtype=np.uint32)
    n, nr = _u64(0), _u64(0)
    _check(lib().orc_prove_recursion(kind, C.c_uint32(po2), ptr(u32(ctrl)), ptr(u32(data)), ptr(u32(accum)),
                                     ptr(u32(glob)), ptr(seal), _u64(cap), C.byref(n), ptr(roots), _u64(16), C.byref(nr),
                                     ptr(qpos)))
    return seal[:n.value].copy(), roots[:8 * nr.value].reshape(-1, 8).copy(), qpos


def verify_recursion(seal, kind=POSEIDON2):
    seal = u32(seal)
    roots = np.zeros(8 * 16, dtype=np.uint32)
    nr = _u64(0)
    _check(lib().orc_verify_recursion(kind, ptr(seal), _u64(len(seal)), ptr(roots), C.byref(nr)))
    return roots[:8 * nr.value].reshape(-1, 8).copy()


def verify_with_validity(seal, circuit="rv32im", kind=POSEIDON2):
    """the restated verifier INCLUDING the constraint check check(z) * ((3z)^N - 1) == poly_ext(eval_u)
    (verify/mod.rs:370-390), poly_ext evaluated from the committed circuit IR (tests/poly_ext_ir.py).
    Returns (roots, validity_checked); raises on any failed check."""
    import poly_ext_ir
    seal = u32(seal)
    roots = np.zeros(8 * 16, dtype=np.uint32)
    nr = _u64(0)
    checked = C.c_int(0)
    fn = lib().orc_verify_rv32im_ext if circuit == "rv32im" else lib().orc_verify_recursion_ext
    _check(fn(kind, ptr(seal), _u64(len(seal)), ptr(roots), C.byref(nr), poly_ext_ir.poly_ext(circuit).callback(),
              C.byref(checked)))
    return roots[:8 * nr.value].reshape(-1, 8).copy(), bool(checked.value)


def synthetic_witness_recursion(po2, seed=None):
    """same generator as synthetic_witness, recursion shapes; the control columns are random too (the prover never
    checks constraint satisfaction)"""
    seed = 0x5EED1000 + po2 if seed is None else seed
    rng = np.random.Generator(np.random.PCG64(seed))
    n = 1 << po2
    return rand_elems(rng, 23 * n), rand_elems(rng, 128 * n), rand_elems(rng, 12 * n), rand_elems(rng, 32)
