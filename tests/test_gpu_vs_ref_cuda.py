"""Second parity oracle on the GPU: our kernels against the REFERENCE'S OWN CUDA kernels (the non-sppark subset,
compiled for sm_100a by `make -C oracle refcuda`), bit for bit on the same device buffers."""
import os
import sys

import numpy as np
import pytest

import oracle_lib as O
from risc0_b200 import B200Hal

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
import ref_cuda as R  # noqa: E402

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not R.available(), reason="oracle/_ref/libref_zkp_cuda.so not built")]


@pytest.fixture(scope="module")
def hal():
    h = B200Hal(0, "sha-256")
    yield h
    h.close()


def test_bit_reverse_fri_fold_sum(hal):
    rng = np.random.default_rng(41)
    lg, cols = 14, 6
    vals = O.rand_elems(rng, cols << lg)
    a, b = hal.copy_from_elem("a", vals), hal.copy_from_elem("b", vals)
    hal.batch_bit_reverse(a, cols)
    hal.sync()
    R.batch_bit_reverse(b, lg, cols << lg)
    assert np.array_equal(a.view(), b.view())
    # fri_fold
    count = 1 << 10
    inp = hal.copy_from_elem("in", O.rand_elems(rng, 64 * count))
    mix = O.rand_ext(rng)
    o1, o2 = hal.alloc_elem("o1", 4 * count), hal.alloc_elem("o2", 4 * count)
    hal.fri_fold(o1, inp, mix)
    hal.sync()
    R.fri_fold(o2, inp, hal.copy_from_extelem("mix", mix), count)
    assert np.array_equal(o1.view(), o2.view())
    # eltwise_sum_extelem
    to_add = 5
    src = hal.copy_from_extelem("src", O.rand_elems(rng, 4 * to_add * count))
    s1, s2 = hal.alloc_elem("s1", 4 * count), hal.alloc_elem("s2", 4 * count)
    hal.eltwise_sum_extelem(s1, src)
    hal.sync()
    R.eltwise_sum_fpext(s2, src, to_add, count)
    assert np.array_equal(s1.view(), s2.view())


def test_mix_poly_coeffs_and_evaluate_any(hal):
    rng = np.random.default_rng(42)
    count, S = 1 << 12, 37
    inp = hal.copy_from_elem("in", O.rand_elems(rng, S * count))
    combos = rng.integers(0, 4, size=S).astype(np.uint32)
    out0 = O.rand_elems(rng, 4 * 5 * count)
    o1, o2 = hal.copy_from_extelem("o1", out0), hal.copy_from_extelem("o2", out0)
    mix_start, mix = O.rand_ext(rng), O.rand_ext(rng)
    hal.mix_poly_coeffs(o1, mix_start, mix, inp, combos, S, count)
    hal.sync()
    R.mix_poly_coeffs(o2, inp, hal.copy_from_u32("c", combos), hal.copy_from_extelem("ms", mix_start),
                      hal.copy_from_extelem("m", mix), S, count)
    assert np.array_equal(o1.view(), o2.view())
    # batch_evaluate_any
    E = 29
    which = hal.copy_from_u32("w", rng.integers(0, S, size=E).astype(np.uint32))
    xs = hal.copy_from_extelem("xs", O.rand_elems(rng, 4 * E))
    e1, e2 = hal.alloc_extelem("e1", E), hal.alloc_extelem("e2", E)
    hal.batch_evaluate_any(inp, S, which, xs, e1)
    hal.sync()
    R.batch_evaluate_any(e2, inp, which, xs, E, count)
    assert np.array_equal(e1.view(), e2.view())


def test_sha_rows_and_fold(hal):
    rng = np.random.default_rng(43)
    rows, cols = 1 << 10, 19
    m = hal.copy_from_elem("m", O.rand_elems(rng, rows * cols))
    n1, n2 = hal.alloc_digest("n1", 2 * rows), hal.alloc_digest("n2", 2 * rows)
    hal.hash_rows(n1.slice(rows, rows), m)
    hal.sync()
    R.sha_rows(n2.slice(rows, rows), m, rows, cols)
    assert np.array_equal(n1.view()[8 * rows:], n2.view()[8 * rows:])
    import ctypes as C
    size = rows
    while size > 1:
        hal.hash_fold(n1, size, size // 2)
        hal.sync()
        R.sha_fold(C.c_void_p(n2.alloc.ptr + (size // 2) * 32), C.c_void_p(n2.alloc.ptr + size * 32), size // 2)
        size //= 2
    assert np.array_equal(n1.view()[8:], n2.view()[8:])
