"""Out-of-bounds write detection without a sanitizer (compute-sanitizer is closed on this pool): every output buffer is a
slice of a larger allocation whose margins are filled with a pattern; after each op the margins must be untouched and the
payload must equal the oracle. Sizes are deliberately ragged."""
import numpy as np
import pytest

import oracle_lib as O
from risc0_b200 import B200Hal

pytestmark = pytest.mark.gpu
GUARD = 4096  # words on each side
PATTERN = 0xDEADBEEF


@pytest.fixture(scope="module")
def hal():
    h = B200Hal(0, "poseidon2")
    yield h
    h.close()


class Guarded:
    def __init__(self, hal, words_per_elem, size, data=None):
        self.hal, self.w, self.size = hal, words_per_elem, size
        total = size * words_per_elem + 2 * GUARD
        host = np.full(total, PATTERN, dtype=np.uint32)
        if data is not None:
            host[GUARD:GUARD + size * words_per_elem] = data
        mk = {1: hal.copy_from_elem, 4: hal.copy_from_extelem, 8: hal.copy_from_digest}[words_per_elem]
        pad = (-total) % words_per_elem
        self.whole = hal.copy_from_elem("guarded", np.concatenate([host, np.full(pad, PATTERN, dtype=np.uint32)]))
        assert GUARD % words_per_elem == 0
        # a typed view of the payload
        from risc0_b200.hal import Buffer
        self.buf = Buffer(hal, "payload", size, words_per_elem, self.whole.alloc, GUARD // words_per_elem)
        del mk

    def check(self):
        all_ = self.whole.view()
        n = self.size * self.w
        assert np.all(all_[:GUARD] == PATTERN), "write before the buffer"
        assert np.all(all_[GUARD + n:GUARD + n + GUARD] == PATTERN), "write past the buffer"
        return all_[GUARD:GUARD + n]


def test_ntt_family_stays_in_bounds(hal):
    rng = np.random.default_rng(77)
    for lg, cols in [(5, 3), (11, 5), (13, 3), (17, 2)]:
        vals = O.rand_elems(rng, cols << lg)
        io = Guarded(hal, 1, cols << lg, vals)
        hal.batch_interpolate_ntt_zk(io.buf, cols)
        co = io.check()
        assert np.array_equal(co, O.zk_shift(O.batch_interpolate_ntt(vals, cols), cols))
        out = Guarded(hal, 1, cols << (lg + 2))
        hal.batch_expand_into_evaluate_ntt(out.buf, io.buf, cols, 2)
        assert np.array_equal(out.check(), O.batch_expand_into_evaluate_ntt(co, cols, 2))
        io.check()
        hal.batch_bit_reverse(io.buf, cols)
        assert np.array_equal(io.check(), O.batch_bit_reverse(co, cols))


def test_merkle_and_ops_stay_in_bounds(hal):
    rng = np.random.default_rng(78)
    rows, cols = 1 << 9, 19
    m = O.rand_elems(rng, rows * cols)
    d_m = hal.copy_from_elem("m", m)
    nodes = Guarded(hal, 8, 2 * rows)
    hal.merkle_build(nodes.buf, d_m, rows, cols)
    got = nodes.check()
    assert np.array_equal(got[8:], O.merkle_tree(O.POSEIDON2, m, rows)[8:])
    # mix_poly_coeffs into a guarded accumulator, then the 5-way sum
    count, S = 1000, 11
    inp = hal.copy_from_elem("in", O.rand_elems(rng, S * count))
    combos = rng.integers(0, 4, size=S).astype(np.uint32)
    acc0 = O.rand_elems(rng, 4 * 5 * count)
    acc = Guarded(hal, 4, 5 * count, acc0)
    ms, mx = O.rand_ext(rng), O.rand_ext(rng)
    hal.mix_poly_coeffs(acc.buf, ms, mx, inp, combos, S, count)
    want = O.mix_poly_coeffs(acc0, ms, mx, O.u32(inp.view()), combos, S, count)
    assert np.array_equal(acc.check(), want)
    out = Guarded(hal, 1, 4 * count)
    hal.eltwise_sum_extelem(out.buf, acc.buf)
    assert np.array_equal(out.check(), O.eltwise_sum_extelem(want, count))
    # division of a ragged-length polynomial
    n = 3001
    p = O.rand_elems(rng, 4 * n)
    z = O.rand_ext(rng)
    g = Guarded(hal, 4, n, p)
    try:
        hal.combos_divide(g.buf, [(0, [z])], n)
    except RuntimeError:
        pass
    want, _ = O.poly_divide_unchecked(p, z)
    assert np.array_equal(g.check(), want)


@pytest.mark.skipif(not O.have_ref(), reason="oracle/_ref not built")
def test_eval_check_stays_in_bounds(hal):
    rng = np.random.default_rng(79)
    po2 = 9
    n, domain = 1 << po2, 4 << po2
    accum, data = O.rand_elems(rng, 103 * domain), O.rand_elems(rng, 211 * domain)
    mix, out, pm = O.rand_elems(rng, 36), O.rand_elems(rng, 90), O.rand_ext(rng)
    check = Guarded(hal, 1, 4 * domain)
    hal.eval_check_rv32im(check.buf, [hal.copy_from_elem("a", accum), hal.alloc_elem_init("c", domain, 0),
                                      hal.copy_from_elem("d", data)],
                          [hal.copy_from_elem("mix", mix), hal.copy_from_elem("out", out)], pm, po2, n)
    assert np.array_equal(check.check(), O.rv32im_eval_check(accum, data, mix, out, pm, po2))
