"""GPU parity tests: every Hal op through the C ABI vs the oracle, bit-exact. Shapes follow the reference's own
A/B harness `hal::testutil` (risc0/zkp/src/hal/mod.rs:319-616, driven for CUDA from hal/cuda.rs:1051-1138), plus the
edge cases it covers (tiny / ragged counts, unaligned sponge tails) and the large two-pass NTT sizes."""
import numpy as np
import pytest

import oracle_lib as O
from risc0_b200 import B200Hal

pytestmark = pytest.mark.gpu
P = O.P


@pytest.fixture(scope="module")
def hal():
    h = B200Hal(0, "poseidon2")
    yield h
    h.close()


@pytest.fixture(scope="module")
def hal_sha():
    h = B200Hal(0, "sha-256")
    yield h
    h.close()


def rng_for(*key):
    return np.random.default_rng([ord(c) for c in "".join(str(k) for k in key)])


# ------------------------------------------------------------------ NTT family
@pytest.mark.parametrize("lg,count", [(0, 3), (1, 5), (2, 7), (3, 2), (4, 3), (5, 9), (7, 4), (8, 1), (10, 6), (11, 3),
                                      (12, 2), (13, 5), (14, 3), (16, 224), (17, 2), (20, 3), (22, 1), (24, 1)])
def test_batch_interpolate_ntt(hal, lg, count):
    rng = rng_for("intt", lg, count)
    vals = O.rand_elems(rng, count << lg)
    io = hal.copy_from_elem("io", vals)
    hal.batch_interpolate_ntt(io, count)
    want = O.batch_interpolate_ntt(vals, count)
    assert np.array_equal(io.view(), want)


@pytest.mark.parametrize("lg,count", [(1, 2), (8, 1000), (12, 900), (13, 3), (20, 2)])
def test_zk_shift_and_fused(hal, lg, count):
    # hal testutil zk_shift shapes: 1000 x 2^8, 900 x 2^12
    rng = rng_for("zk", lg, count)
    vals = O.rand_elems(rng, count << lg)
    io = hal.copy_from_elem("io", vals)
    hal.zk_shift(io, count)
    assert np.array_equal(io.view(), O.zk_shift(vals, count))
    io2 = hal.copy_from_elem("io2", vals)
    hal.batch_interpolate_ntt_zk(io2, count)
    assert np.array_equal(io2.view(), O.zk_shift(O.batch_interpolate_ntt(vals, count), count))


@pytest.mark.parametrize("lg_in,count,eb", [(0, 3, 2), (1, 5, 2), (2, 4, 2), (3, 3, 2), (6, 2, 2), (8, 7, 2), (10, 5, 2),
                                            (11, 2, 2), (12, 3, 2), (13, 2, 2), (16, 224, 2), (18, 4, 2), (20, 4, 2),
                                            (22, 1, 2), (5, 3, 0), (13, 2, 0), (16, 3, 0)])
def test_batch_expand_into_evaluate_ntt(hal, lg_in, count, eb):
    rng = rng_for("lde", lg_in, count, eb)
    coeffs = O.rand_elems(rng, count << lg_in)
    inp = hal.copy_from_elem("in", coeffs)
    out = hal.alloc_elem("out", count << (lg_in + eb))
    hal.batch_expand_into_evaluate_ntt(out, inp, count, eb)
    want = O.batch_expand_into_evaluate_ntt(coeffs, count, eb)
    assert np.array_equal(out.view(), want)
    assert np.array_equal(inp.view(), coeffs)  # input untouched


def test_ntt_roundtrip_large(hal):
    # size-independent property at full size: interpolate then evaluate (no expansion) returns the input; 16 x 2^20
    rng = rng_for("roundtrip")
    vals = O.rand_elems(rng, 16 << 20)
    io = hal.copy_from_elem("io", vals)
    hal.batch_interpolate_ntt(io, 16)
    out = hal.alloc_elem("out", 16 << 20)
    hal.batch_expand_into_evaluate_ntt(out, io, 16, 0)
    assert np.array_equal(out.view(), vals)


@pytest.mark.parametrize("lg,count", [(0, 2), (1, 3), (2, 3), (3, 5), (4, 2), (7, 3), (9, 2), (10, 4), (11, 3), (14, 224),
                                      (20, 2), (22, 1)])
def test_batch_bit_reverse(hal, lg, count):
    rng = rng_for("brev", lg, count)
    vals = O.rand_elems(rng, count << lg)
    io = hal.copy_from_elem("io", vals)
    hal.batch_bit_reverse(io, count)
    assert np.array_equal(io.view(), O.batch_bit_reverse(vals, count))
    hal.batch_bit_reverse(io, count)  # involution
    assert np.array_equal(io.view(), vals)


@pytest.mark.parametrize("lg,count,first", [(10, 3, 1), (12, 5, 2), (16, 7, 3)])
def test_batch_bit_reverse_on_a_slice(hal, lg, count, first):
    # the TMA kernel (csrc/bitrev_tma.cu) describes the columns it is given as one tensor: a slice that starts at a
    # later column must leave the columns before and after it untouched
    rng = rng_for("brev_slice", lg, count)
    total = count + first + 1
    vals = O.rand_elems(rng, total << lg)
    io = hal.copy_from_elem("io", vals)
    hal.batch_bit_reverse(io.slice(first << lg, count << lg), count)
    want = vals.copy()
    want[first << lg:(first + count) << lg] = O.batch_bit_reverse(vals[first << lg:(first + count) << lg], count)
    assert np.array_equal(io.view(), want)


def test_batch_evaluate_any(hal):
    # hal testutil: 223 polys x 2^16, 865 evaluation points
    rng = rng_for("evalany")
    polys, lg, evals = 223, 16, 865
    coeffs = O.rand_elems(rng, polys << lg)
    which = rng.integers(0, polys, size=evals).astype(np.uint32)
    xs = O.rand_elems(rng, 4 * evals)
    out = hal.alloc_extelem("out", evals)
    hal.batch_evaluate_any(hal.copy_from_elem("coeffs", coeffs), polys, hal.copy_from_u32("which", which),
                           hal.copy_from_extelem("xs", xs), out)
    assert np.array_equal(out.view(), O.batch_evaluate_any(coeffs, polys, which, xs))


@pytest.mark.parametrize("lg,polys,evals", [(3, 2, 3), (8, 3, 4), (13, 2, 5), (14, 1, 1)])
def test_batch_evaluate_any_small(hal, lg, polys, evals):
    rng = rng_for("evalany", lg)
    coeffs = O.rand_elems(rng, polys << lg)
    which = rng.integers(0, polys, size=evals).astype(np.uint32)
    xs = O.rand_elems(rng, 4 * evals)
    out = hal.alloc_extelem("out", evals)
    hal.batch_evaluate_any(hal.copy_from_elem("coeffs", coeffs), polys, hal.copy_from_u32("which", which),
                           hal.copy_from_extelem("xs", xs), out)
    assert np.array_equal(out.view(), O.batch_evaluate_any(coeffs, polys, which, xs))


@pytest.mark.parametrize("count", [1 << 12, 1002, 7, 4])   # multiples of 4 take the 128-bit kernel, the others the scalar one
def test_mix_poly_coeffs(hal, count):
    # hal testutil: 16 x 2^14 inputs -> (100 + 1) combos x 2^12 ... we keep its structure: several inputs per combo,
    # non-zero initial output, unsorted combo ids
    rng = rng_for("mix", count)
    input_size, combo_count = 37, 6
    inp = O.rand_elems(rng, input_size * count)
    combos = rng.integers(0, combo_count, size=input_size).astype(np.uint32)
    out0 = O.rand_elems(rng, 4 * combo_count * count)
    mix_start, mix = O.rand_ext(rng), O.rand_ext(rng)
    out = hal.copy_from_extelem("out", out0)
    hal.mix_poly_coeffs(out, mix_start, mix, hal.copy_from_elem("in", inp), combos, input_size, count)
    want = O.mix_poly_coeffs(out0, mix_start, mix, inp, combos, input_size, count)
    assert np.array_equal(out.view(), want)


# ------------------------------------------------------------------ element-wise ops (counts from hal testutil)
COUNTS = [1, 9, 12, 1001, 1024, 1025, 1 << 20]


@pytest.mark.parametrize("n", COUNTS)
def test_eltwise_add_copy(hal, n):
    rng = rng_for("elt", n)
    a, b = O.rand_elems(rng, n), O.rand_elems(rng, n)
    out = hal.alloc_elem("out", n)
    hal.eltwise_add_elem(out, hal.copy_from_elem("a", a), hal.copy_from_elem("b", b))
    assert np.array_equal(out.view(), O.eltwise_add_elem(a, b))
    hal.eltwise_copy_elem(out, hal.copy_from_elem("a", a))
    assert np.array_equal(out.view(), a)


@pytest.mark.parametrize("n", COUNTS)
def test_eltwise_zeroize(hal, n):
    rng = rng_for("zero", n)
    a = O.rand_elems(rng, n)
    a[rng.integers(0, n, size=max(1, n // 3))] = 0xFFFFFFFF
    io = hal.copy_from_elem("io", a)
    hal.eltwise_zeroize_elem(io)
    assert np.array_equal(io.view(), O.eltwise_zeroize_elem(a))


def test_alloc_init_and_slices(hal):
    b = hal.alloc_elem_init("init", 1000, 0xFFFFFFFF)
    assert np.all(b.view() == 0xFFFFFFFF)
    z = hal.alloc_extelem_zeroed("z", 77)
    assert np.all(z.view() == 0)
    # hal testutil `slice`: 4096 x 256 buffer, operate on a slice only
    rng = rng_for("slice")
    data = O.rand_elems(rng, 4096 * 4)
    buf = hal.copy_from_elem("buf", data)
    s = buf.slice(4096, 2 * 4096)
    hal.batch_interpolate_ntt(s, 2)
    want = data.copy()
    want[4096:3 * 4096] = O.batch_interpolate_ntt(data[4096:3 * 4096], 2)
    assert np.array_equal(buf.view(), want)
    assert np.array_equal(s.get_at(5), want[4096 + 5:4096 + 6])


@pytest.mark.parametrize("count,to_add", [(1, 1), (9, 5), (1025, 5), (1 << 16, 5), (1000, 1)])
def test_eltwise_sum_extelem(hal, count, to_add):
    rng = rng_for("sum", count, to_add)
    inp = O.rand_elems(rng, 4 * count * to_add)
    out = hal.alloc_elem("out", 4 * count)
    hal.eltwise_sum_extelem(out, hal.copy_from_extelem("in", inp))
    assert np.array_equal(out.view(), O.eltwise_sum_extelem(inp, count))


@pytest.mark.parametrize("count", [1, 9, 12, 256, 1001, 1 << 16])
def test_fri_fold(hal, count):
    rng = rng_for("fri", count)
    inp = O.rand_elems(rng, 4 * 16 * count)
    mix = O.rand_ext(rng)
    out = hal.alloc_elem("out", 4 * count)
    hal.fri_fold(out, hal.copy_from_elem("in", inp), mix)
    assert np.array_equal(out.view(), O.fri_fold(inp, mix))


# ------------------------------------------------------------------ hashing
@pytest.mark.parametrize("rows", [1, 2, 3, 4, 10, 1000])
@pytest.mark.parametrize("cols", [0, 1, 15, 16, 17, 32, 64, 103, 128, 211])
def test_hash_rows_poseidon2(hal, rows, cols):
    rng = rng_for("rowsp2", rows, cols)
    m = O.rand_elems(rng, rows * cols)
    out = hal.alloc_digest("out", rows)
    hal.hash_rows(out, hal.copy_from_elem("m", m) if cols else hal.alloc_elem("m", 0))
    assert np.array_equal(out.view(), O.hash_rows(O.POSEIDON2, m, rows))


def test_hash_rows_tensor_core_variant():
    # poseidon2.cu's opt-in kernel that runs the 21 partial rounds' linear algebra as u8 IMMA products
    # (R0B200_P2_VARIANT=6; the switch is read once per process, hence the subprocess): same digests as the oracle,
    # incl. ragged last sponge blocks and a row count that is not a multiple of the warp size
    import os
    import subprocess
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    code = (
        "import sys, numpy as np; sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
        "import oracle_lib as O\n"
        "from risc0_b200 import B200Hal\n"
        "hal = B200Hal(0, 'poseidon2')\n"
        "for rows, cols in [(64, 16), (256, 1), (1024, 37), (4096, 211), (96, 103), (10, 64)]:\n"
        "    rng = np.random.default_rng(rows * 1000 + cols)\n"
        "    m = O.rand_elems(rng, rows * cols)\n"
        "    out = hal.alloc_digest('out', rows)\n"
        "    hal.hash_rows(out, hal.copy_from_elem('m', m))\n"
        "    assert np.array_equal(out.view(), O.hash_rows(O.POSEIDON2, m, rows)), (rows, cols)\n"
        "hal.close(); print('tc variant ok')\n" % (os.path.dirname(here), here))
    r = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, R0B200_P2_VARIANT="6"), capture_output=True,
                       text=True, timeout=600)
    assert r.returncode == 0 and "tc variant ok" in r.stdout, r.stdout + r.stderr


@pytest.mark.parametrize("rows,cols", [(1, 16), (3, 32), (10, 64), (1000, 17), (257, 211), (4, 0)])
def test_hash_rows_sha(hal_sha, rows, cols):
    rng = rng_for("rowssha", rows, cols)
    m = O.rand_elems(rng, rows * cols)
    out = hal_sha.alloc_digest("out", rows)
    hal_sha.hash_rows(out, hal_sha.copy_from_elem("m", m) if cols else hal_sha.alloc_elem("m", 0))
    assert np.array_equal(out.view(), O.hash_rows(O.SHA256, m, rows))


def test_hash_rows_sha_kat(hal_sha):
    # hal/cpu.rs:726-733
    out = hal_sha.alloc_digest("out", 1)
    hal_sha.hash_rows(out, hal_sha.alloc_elem_init("m", 16, 0))
    assert out.view().tobytes().hex() == "da5698be17b9b46962335799779fbeca8ce5d491c0d26243bafef9ea1837a9d8"


@pytest.mark.parametrize("kind", ["poseidon2", "sha-256"])
def test_hash_fold(hal, hal_sha, kind):
    # hal testutil: 1024 digests whose words are valid field elements
    h = hal if kind == "poseidon2" else hal_sha
    k = O.POSEIDON2 if kind == "poseidon2" else O.SHA256
    rng = rng_for("fold", kind)
    n = 1024
    nodes = np.zeros(16 * n, dtype=np.uint32)
    nodes[8 * n:] = O.rand_elems(rng, 8 * n)
    io = h.copy_from_digest("io", nodes)
    want = nodes
    size = n
    while size > 1:
        h.hash_fold(io, size, size // 2)
        want = O.hash_fold(k, want, size, size // 2)
        size //= 2
    assert np.array_equal(io.view()[8:], want[8:])


@pytest.mark.parametrize("kind", ["poseidon2", "sha-256"])
@pytest.mark.parametrize("rows,cols", [(2, 5), (8, 16), (512, 33), (1 << 14, 16), (1 << 16, 3)])
def test_merkle_build(hal, hal_sha, kind, rows, cols):
    h = hal if kind == "poseidon2" else hal_sha
    k = O.POSEIDON2 if kind == "poseidon2" else O.SHA256
    rng = rng_for("merkle", kind, rows, cols)
    m = O.rand_elems(rng, rows * cols)
    nodes = h.alloc_digest("nodes", 2 * rows)
    h.merkle_build(nodes, h.copy_from_elem("m", m), rows, cols)
    assert np.array_equal(nodes.view()[8:], O.merkle_tree(k, m, rows)[8:])


# ------------------------------------------------------------------ gather / scatter / misc
def test_gather_sample(hal):
    # hal testutil: 1000 x 900
    rng = rng_for("gather")
    rows, cols = 1000, 900
    src = O.rand_elems(rng, rows * cols)
    d_src = hal.copy_from_elem("src", src)
    for idx in (0, 1, 499, 999):
        dst = hal.alloc_elem("dst", cols)
        hal.gather_sample(dst, d_src, idx, cols, rows)
        assert np.array_equal(dst.view(), O.gather_sample(src, idx, cols, rows))


def test_scatter(hal):
    rng = rng_for("scatter")
    cycles, n = 300, 5000
    per = rng.integers(0, 6, size=cycles)
    index = np.concatenate([[0], np.cumsum(per)]).astype(np.uint32)
    total = int(index[-1])
    offsets = rng.permutation(n)[:total].astype(np.uint32)
    values = O.rand_elems(rng, total)
    into0 = np.full(n, 0xFFFFFFFF, dtype=np.uint32)
    into = hal.copy_from_elem("into", into0)
    hal.scatter(into, index, offsets, values)
    assert np.array_equal(into.view(), O.scatter(into0, index, offsets, values))
    hal.scatter(into, np.zeros(0, dtype=np.uint32), offsets, values)  # empty index is a no-op


def test_eltwise_copy_elem_slice(hal):
    rng = rng_for("copyslice")
    frm = O.rand_elems(rng, 50 * 40)
    into0 = O.rand_elems(rng, 100 * 64)
    into = hal.copy_from_elem("into", into0)
    args = dict(from_rows=20, from_cols=33, from_offset=7, from_stride=40, into_offset=11, into_stride=64)
    hal.eltwise_copy_elem_slice(into, frm, **args)
    want = O.eltwise_copy_elem_slice(into0, frm, 20, 33, 7, 40, 11, 64)
    assert np.array_equal(into.view(), want)


def test_prefix_products(hal):
    rng = rng_for("prefix")
    io0 = O.rand_elems(rng, 4 * 100)
    io = hal.copy_from_extelem("io", io0)
    hal.prefix_products(io)
    assert np.array_equal(io.view(), O.prefix_products(io0))


def test_combos_prepare(hal):
    rng = rng_for("prep")
    cycles, combo_count = 256, 4
    reg_sizes = np.array([1, 2, 6, 5, 1, 2, 6], dtype=np.uint32)
    reg_combo = np.array([0, 1, 2, 3, 0, 1, 2], dtype=np.uint32)
    coeff_u = O.rand_elems(rng, 4 * (int(reg_sizes.sum()) + 16))
    combos0 = O.rand_elems(rng, 4 * cycles * (combo_count + 1))
    mix = O.rand_ext(rng)
    combos = hal.copy_from_extelem("combos", combos0)
    hal.combos_prepare(combos, coeff_u, combo_count, cycles, reg_sizes, reg_combo, mix)
    want = O.combos_prepare(combos0, coeff_u, combo_count, cycles, reg_sizes, reg_combo, mix)
    assert np.array_equal(combos.view(), want)


def _poly_with_roots(rng, n, roots):
    """coefficients of q(x) * prod (x - r): guaranteed zero remainders"""
    L = O.lib()
    q = O.rand_elems(rng, 4 * (n - len(roots)))
    poly = np.concatenate([q, np.zeros(4 * len(roots), dtype=np.uint32)])
    tmp = np.zeros(4, dtype=np.uint32)
    for r in roots:  # multiply by (x - r): new[i] = old[i-1] - r*old[i]
        new = np.zeros_like(poly)
        for i in range(n):
            acc = np.zeros(4, dtype=np.uint32)
            if i > 0:
                acc = poly[4 * (i - 1):4 * i].copy()
            L.orc_fpext_mul(O.ptr(tmp), O.ptr(O.u32(r)), O.ptr(O.u32(poly[4 * i:4 * i + 4])))
            new[4 * i:4 * i + 4] = [L.orc_fp_sub(int(a), int(b)) for a, b in zip(acc, tmp)]
        poly = new
    return poly


@pytest.mark.parametrize("cycles", [64, 100, 256, 1 << 12, (1 << 14) + 64])
def test_combos_divide(hal, cycles):
    rng = rng_for("div", cycles)
    roots_a = [O.rand_ext(rng) for _ in range(3)]
    roots_b = [O.rand_ext(rng)]
    small = cycles <= 256
    if small:
        pa, pb = _poly_with_roots(rng, cycles, roots_a), _poly_with_roots(rng, cycles, roots_b)
    else:
        # build divisible polynomials with the oracle itself: take random p, subtract the remainder polynomial is
        # expensive; instead divide-then-multiply is equivalent to checking quotient parity on arbitrary input, so use
        # the unchecked single division for parity and the checked API only on the small sizes
        pa = pb = None
    if small:
        combos0 = np.concatenate([pa, pb])
        combos = hal.copy_from_extelem("combos", combos0)
        hal.combos_divide(combos, [(0, roots_a), (1, roots_b)], cycles)
        pow_begin = np.array([0, 3, 4], dtype=np.uint32)
        want = O.combos_divide(combos0, pow_begin, np.concatenate(roots_a + roots_b), cycles)
        assert np.array_equal(combos.view(), want)
        # a non-divisible polynomial must be reported, as the reference asserts remainder == 0
        bad = combos0.copy()
        bad[0] ^= 1
        with pytest.raises(RuntimeError):
            hal.combos_divide(hal.copy_from_extelem("bad", bad), [(0, roots_a), (1, roots_b)], cycles)
    else:
        p = O.rand_elems(rng, 4 * cycles)
        z = O.rand_ext(rng)
        want, _ = O.poly_divide_unchecked(p, z)
        buf = hal.copy_from_extelem("p", p)
        try:
            hal.combos_divide(buf, [(0, [z])], cycles)
        except RuntimeError:
            pass  # remainder is non-zero for a random polynomial; the quotient must still match
        assert np.array_equal(buf.view(), want)
