"""The reference's own FFI symbol table (include/r0b200_compat.h) exported by libr0b200.so: every risc0_zkp_cuda_* /
sppark_* / supra_poly_divide / risc0_circuit_*_cuda_eval_check entry point called with the reference's argument
conventions (risc0/sys/kernels/zkp/cuda/ffi.cu:25-145, risc0/sys/src/cuda.rs:19-80, launch geometry as
zkp/src/hal/cuda.rs passes it) and compared bit for bit with the oracle. The same harness (tools/ref_cuda.py) drives
the reference's own compiled kernels in test_gpu_vs_ref_cuda.py."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

import oracle_lib as O
from risc0_b200 import B200Hal, _lib

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))

pytestmark = pytest.mark.gpu
u32 = C.c_uint32


class SpparkError(C.Structure):   # sppark::Error, returned by value
    _fields_ = [("code", C.c_int32), ("message", C.c_char_p)]


@pytest.fixture(scope="module")
def hal():
    h = B200Hal(0, "poseidon2")
    yield h
    h.close()


@pytest.fixture(scope="module")
def L():
    lib = C.CDLL(_lib.LIB_PATH)
    for n in _lib.compat_symbols():
        getattr(lib, n).restype = SpparkError if n.startswith(("sppark_", "supra_")) else C.c_char_p
    return lib


def dev(hal, kind, name, host):
    """upload through `hal` and wait: the compat entry points run on the library's default context (another stream)"""
    buf = getattr(hal, "copy_from_" + kind)(name, host)
    hal.sync()
    return buf


def ok(e):
    if isinstance(e, SpparkError):
        assert e.code == 0, e.message
    else:
        assert e is None, e


def test_ref_harness_against_compat_layer(hal):
    """tools/ref_cuda.py (written for the reference's compiled kernels) bound to libr0b200.so instead"""
    import ref_cuda as R
    R.use(_lib.LIB_PATH)
    try:
        rng = np.random.default_rng(7)
        lg, cols = 13, 5
        vals = O.rand_elems(rng, cols << lg)
        b = dev(hal, "elem", "b", vals)
        R.batch_bit_reverse(b, lg, cols << lg)
        assert np.array_equal(b.view(), O.batch_bit_reverse(vals, cols))
        count = 1 << 9
        inp_h = O.rand_elems(rng, 64 * count)
        mix = O.rand_ext(rng)
        o = hal.alloc_elem("o", 4 * count)
        R.fri_fold(o, dev(hal, "elem", "in", inp_h), dev(hal, "extelem", "mix", mix), count)
        assert np.array_equal(o.view(), O.fri_fold(inp_h, mix))
        to_add = 5
        src_h = O.rand_elems(rng, 4 * to_add * count)
        s = hal.alloc_elem("s", 4 * count)
        R.eltwise_sum_fpext(s, dev(hal, "extelem", "src", src_h), to_add, count)
        assert np.array_equal(s.view(), O.eltwise_sum_extelem(src_h, count))
        # mix_poly_coeffs + batch_evaluate_any
        n, S = 1 << 11, 23
        in_h = O.rand_elems(rng, S * n)
        combos = rng.integers(0, 4, size=S).astype(np.uint32)
        out0 = O.rand_elems(rng, 4 * 5 * n)
        ms, m = O.rand_ext(rng), O.rand_ext(rng)
        d_in = dev(hal, "elem", "in", in_h)
        d_out = dev(hal, "extelem", "out", out0)
        R.mix_poly_coeffs(d_out, d_in, dev(hal, "u32", "c", combos), dev(hal, "extelem", "ms", ms),
                          dev(hal, "extelem", "m", m), S, n)
        assert np.array_equal(d_out.view(), O.mix_poly_coeffs(out0, ms, m, in_h, combos, S, n))
        E = 17
        which = rng.integers(0, S, size=E).astype(np.uint32)
        xs = O.rand_elems(rng, 4 * E)
        e = hal.alloc_extelem("e", E)
        R.batch_evaluate_any(e, d_in, dev(hal, "u32", "w", which), dev(hal, "extelem", "xs", xs), E, n)
        assert np.array_equal(e.view(), O.batch_evaluate_any(in_h, S, which, xs))
        # sha rows / fold / gather
        rows, c = 1 << 9, 19
        m_h = O.rand_elems(rng, rows * c)
        d_m = dev(hal, "elem", "m", m_h)
        nodes = hal.alloc_digest("n", 2 * rows)
        R.sha_rows(nodes.slice(rows, rows), d_m, rows, c)
        size = rows
        while size > 1:
            R.sha_fold(C.c_void_p(nodes.alloc.ptr + (size // 2) * 32), C.c_void_p(nodes.alloc.ptr + size * 32), size // 2)
            size //= 2
        assert np.array_equal(nodes.view()[8:], O.merkle_tree(O.SHA256, m_h, rows)[8:])
        g = hal.alloc_elem("g", c)
        R.gather_sample(g, d_m, 77, c, rows)
        assert np.array_equal(g.view(), O.gather_sample(m_h, 77, c, rows))
    finally:
        if R.available():
            R.use(R.LIB)


@pytest.mark.parametrize("lg,count", [(0, 2), (3, 5), (10, 7), (13, 3), (16, 8)])
def test_sppark_ntt_family(hal, L, lg, count):
    rng = np.random.default_rng(100 + lg)
    vals = O.rand_elems(rng, count << lg)
    io = dev(hal, "elem", "io", vals)
    ok(L.sppark_init())
    ok(L.sppark_batch_iNTT(io.ptr, u32(lg), u32(count)))
    coeffs = O.batch_interpolate_ntt(vals, count)
    assert np.array_equal(io.view(), coeffs)
    ok(L.sppark_batch_zk_shift(io.ptr, u32(lg), u32(count)))
    shifted = O.zk_shift(coeffs, count)
    assert np.array_equal(io.view(), shifted)
    # hal/cuda.rs:523-573: expand then NTT
    out = hal.alloc_elem("out", count << (lg + 2))
    ok(L.sppark_batch_expand(out.ptr, io.ptr, u32(lg), u32(2), u32(count)))
    ok(L.sppark_batch_NTT(out.ptr, u32(lg + 2), u32(count)))
    if lg == 0:   # supra/ntt.cu:35-36 returns before touching anything when lg_domain_size == 0
        return
    assert np.array_equal(out.view(), O.batch_expand_into_evaluate_ntt(shifted, count, 2))


@pytest.mark.parametrize("rows,cols", [(1, 16), (3, 17), (10, 128), (1 << 12, 211), (1 << 10, 1)])
def test_sppark_poseidon2(hal, L, rows, cols):
    rng = np.random.default_rng(200 + cols)
    m_h = O.rand_elems(rng, rows * cols)
    out = hal.alloc_digest("d", rows)
    d_m = dev(hal, "elem", "m", m_h)
    ok(L.sppark_poseidon2_rows(out.ptr, d_m.ptr, u32(rows), u32(cols)))
    want = O.hash_rows(O.POSEIDON2, m_h, rows)
    assert np.array_equal(out.view(), want)
    if rows >= 2 and rows & (rows - 1) == 0:
        nodes_h = np.zeros(16 * rows, dtype=np.uint32)
        nodes_h[8 * rows:] = want
        nodes = dev(hal, "digest", "n", nodes_h)
        size = rows
        while size > 1:   # hal/cuda.rs:145-155: output = io + output_size, input = io + 2 * output_size
            ok(L.sppark_poseidon2_fold(C.c_void_p(nodes.alloc.ptr + (size // 2) * 32), C.c_void_p(nodes.alloc.ptr + size * 32),
                                       C.c_size_t(size // 2)))
            size //= 2
        assert np.array_equal(nodes.view()[8:], O.merkle_tree(O.POSEIDON2, m_h, rows)[8:])
    e = L.sppark_poseidon254_rows(None, None, C.c_size_t(0), u32(0))
    assert e.code != 0 and b"scope" in e.message


def test_supra_poly_divide_and_eltwise(hal, L):
    rng = np.random.default_rng(300)
    n = 1 << 12
    poly = O.rand_elems(rng, 4 * n)
    z = O.rand_ext(rng)
    d = dev(hal, "extelem", "p", poly)
    rem = np.zeros(4, dtype=np.uint32)
    ok(L.supra_poly_divide(d.ptr, C.c_size_t(n), rem.ctypes.data_as(C.POINTER(u32)), z.ctypes.data_as(C.POINTER(u32))))
    want_q, want_r = O.poly_divide(poly, z)
    assert np.array_equal(rem, want_r)
    assert np.array_equal(d.view()[:4 * (n - 1)], want_q[:4 * (n - 1)])
    # eltwise family with u32 counts (ragged sizes of hal::testutil)
    for count in (1, 9, 1001, 1025):
        a_h, b_h = O.rand_elems(rng, count), O.rand_elems(rng, count)
        a, b, o = hal.copy_from_elem("a", a_h), hal.copy_from_elem("b", b_h), hal.alloc_elem("o", count)
        hal.sync()
        ok(L.risc0_zkp_cuda_eltwise_add_fp(o.ptr, a.ptr, b.ptr, u32(count)))
        assert np.array_equal(o.view(), O.eltwise_add_elem(a_h, b_h))
        ok(L.risc0_zkp_cuda_eltwise_copy_fp(o.ptr, a.ptr, u32(count)))
        assert np.array_equal(o.view(), a_h)
        factor = int(O.encode(np.array([7], dtype=np.uint32))[0])
        ok(L.risc0_zkp_cuda_eltwise_mul_factor_fp(o.ptr, u32(factor), u32(count)))
        assert np.array_equal(O.decode(o.view()), (O.decode(a_h).astype(np.uint64) * 7 % O.P).astype(np.uint32))
        inv = a_h.copy()
        inv[::3] = 0xFFFFFFFF
        z_ = dev(hal, "elem", "z", inv)
        ok(L.risc0_zkp_cuda_eltwise_zeroize_fp(z_.ptr, u32(count)))
        assert np.array_equal(z_.view(), O.eltwise_zeroize_elem(inv))
    # scatter / copy_region with device-side arguments (hal/cuda.rs:850-935)
    rows, cols = 64, 5
    into_h = np.full(rows * cols, 0xFFFFFFFF, dtype=np.uint32)
    index = np.arange(0, 2 * rows + 1, 2, dtype=np.uint32)
    offsets = rng.permutation(rows * cols)[:2 * rows].astype(np.uint32)
    values = O.rand_elems(rng, 2 * rows)
    into = hal.copy_from_elem("into", into_h)
    d_i, d_o, d_v = dev(hal, "u32", "i", index), dev(hal, "u32", "o", offsets), dev(hal, "elem", "v", values)
    ok(L.risc0_zkp_cuda_scatter(into.ptr, d_i.ptr, d_o.ptr, d_v.ptr, u32(rows)))
    assert np.array_equal(into.view(), O.scatter(into_h, index, offsets, values))
    frm = O.rand_elems(rng, 4 * 10)
    into2_h = O.rand_elems(rng, 200)
    into2, d_f = dev(hal, "elem", "into2", into2_h), dev(hal, "elem", "f", frm)
    ok(L.risc0_zkp_cuda_eltwise_copy_fp_region(into2.ptr, d_f.ptr, u32(3), u32(7), u32(2), u32(10), u32(11), u32(50)))
    assert np.array_equal(into2.view(), O.eltwise_copy_elem_slice(into2_h, frm, 3, 7, 2, 10, 11, 50))


@pytest.mark.skipif(not O.have_ref(), reason="reference poly_fp not built")
def test_rv32im_eval_check_compat(hal, L):
    po2 = 9
    n, dom = 1 << po2, 4 << po2
    rng = np.random.default_rng(400)
    accum, code, data = O.rand_elems(rng, 103 * dom), np.zeros(dom, dtype=np.uint32), O.rand_elems(rng, 211 * dom)
    mix, out, poly_mix = O.rand_elems(rng, 36), O.rand_elems(rng, 90), O.rand_ext(rng)
    pows = O.rv32im_poly_mix_pows(poly_mix)
    rou = np.array([O.rou_fwd(po2 + 2)], dtype=np.uint32)
    check = hal.alloc_elem("check", 4 * dom)
    p = lambda a: a.ctypes.data_as(C.POINTER(u32))
    bufs = [dev(hal, "elem", n_, a_) for n_, a_ in (("c", code), ("d", data), ("a", accum), ("m", mix), ("o", out))]
    ok(L.risc0_circuit_rv32im_cuda_eval_check(check.ptr, bufs[0].ptr, bufs[1].ptr, bufs[2].ptr, bufs[3].ptr, bufs[4].ptr,
                                              p(rou), u32(po2), u32(dom), p(pows)))
    want = O.rv32im_eval_check(accum, data, mix, out, poly_mix, po2)
    assert np.array_equal(check.view(), want)
    bad = rou.copy()
    bad[0] ^= 1
    e = L.risc0_circuit_rv32im_cuda_eval_check(check.ptr, None, None, None, None, None, p(bad), u32(po2), u32(dom), p(pows))
    assert e is not None and b"rou" in e
