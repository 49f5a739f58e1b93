"""Error behaviour of the C ABI (the reference convention, risc0/sys/src/lib.rs:53-75: NULL = ok, otherwise a message
the caller frees): bad arguments are reported as errors, never crash, and leave the context usable."""
import ctypes as C

import numpy as np
import pytest

import oracle_lib as O
from risc0_b200 import B200Hal, R0B200Error, SegmentProver
from risc0_b200._lib import load_library

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def hal():
    h = B200Hal(0, "poseidon2")
    yield h
    h.close()


def test_bad_arguments_are_errors_not_crashes(hal):
    lib = load_library()
    buf = hal.alloc_elem("x", 1 << 10)
    # NTT size beyond MAX_CYCLES_PO2 + 2
    with pytest.raises(R0B200Error):
        hal._l  # noqa: B018
        from risc0_b200._lib import check
        check(lib.r0b200_batch_interpolate_ntt(hal._ctx, buf.ptr, C.c_size_t(1), C.c_uint32(25)))
    # expand_bits other than 0 / 2
    out = hal.alloc_elem("y", 1 << 11)
    with pytest.raises(R0B200Error):
        hal.batch_expand_into_evaluate_ntt(out, buf, 1, 1)
    # unknown hash suite
    from risc0_b200._lib import check
    with pytest.raises(R0B200Error):
        check(lib.r0b200_hash_rows(hal._ctx, 7, out.ptr, buf.ptr, C.c_size_t(4), C.c_size_t(4)))
    # hash_fold with input_size != 2 * output_size
    nodes = hal.alloc_digest("n", 64)
    with pytest.raises(R0B200Error):
        check(lib.r0b200_hash_fold(hal._ctx, 0, nodes.ptr, C.c_size_t(16), C.c_size_t(4)))
    # merkle_build on a non power of two
    with pytest.raises(R0B200Error):
        check(lib.r0b200_merkle_build(hal._ctx, 0, nodes.ptr, buf.ptr, C.c_size_t(24), C.c_size_t(2)))
    # null context
    with pytest.raises(R0B200Error):
        check(lib.r0b200_sync(None))
    # device ordinal out of range
    ctx = C.c_void_p()
    with pytest.raises(R0B200Error):
        check(lib.r0b200_create(99, C.byref(ctx)))
    # the context still works
    vals = O.rand_elems(np.random.default_rng(0), 1 << 10)
    io = hal.copy_from_elem("io", vals)
    hal.batch_interpolate_ntt(io, 1)
    assert np.array_equal(io.view(), O.batch_interpolate_ntt(vals, 1))


def test_prove_rejects_out_of_range_po2(hal):
    code, data, accum, glob = O.synthetic_witness(9)
    with pytest.raises(R0B200Error):
        SegmentProver(hal).prove(8, code[:256], data[:211 * 256], accum[:103 * 256], glob)


def test_two_contexts_on_one_device_are_independent():
    # one prover per context; two contexts (e.g. two host threads) may share a GPU
    a, b = B200Hal(0), B200Hal(0)
    try:
        rng = np.random.default_rng(5)
        va, vb = O.rand_elems(rng, 4 << 12), O.rand_elems(rng, 4 << 12)
        ba, bb = a.copy_from_elem("a", va), b.copy_from_elem("b", vb)
        a.batch_interpolate_ntt(ba, 4)
        b.batch_interpolate_ntt_zk(bb, 4)
        assert np.array_equal(ba.view(), O.batch_interpolate_ntt(va, 4))
        assert np.array_equal(bb.view(), O.zk_shift(O.batch_interpolate_ntt(vb, 4), 4))
    finally:
        a.close()
        b.close()
