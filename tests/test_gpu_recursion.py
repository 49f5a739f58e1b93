"""GPU parity for the recursion circuit (BASELINE config 4: every lift / join is one fixed-size
recursion proof): eval_check vs the reference's own compiled recursion poly_fp, and whole proofs vs the oracle's
restated prover, bit-exact, through the C ABI."""
import numpy as np
import pytest

import oracle_lib as O
from risc0_b200 import B200Hal, SegmentProver

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not O.have_ref_recursion(), reason="oracle/_ref recursion not built")]


@pytest.fixture(scope="module")
def hal():
    h = B200Hal(0, "poseidon2")
    yield h
    h.close()


@pytest.mark.parametrize("po2", [9, 12])
def test_eval_check_recursion_matches_reference(hal, po2):
    rng = np.random.default_rng(300 + po2)
    n, domain = 1 << po2, 4 << po2
    ctrl, data, accum = O.rand_elems(rng, 23 * domain), O.rand_elems(rng, 128 * domain), O.rand_elems(rng, 12 * domain)
    mix, out, poly_mix = O.rand_elems(rng, 20), O.rand_elems(rng, 32), O.rand_ext(rng)
    d = [hal.copy_from_elem("m", x) for x in (accum, ctrl, data)]
    check = hal.alloc_elem("check", 4 * domain)
    hal.eval_check_recursion(check, d, [hal.copy_from_elem("mix", mix), hal.copy_from_elem("out", out)], poly_mix, po2, n)
    want = O.recursion_eval_check(ctrl, data, accum, mix, out, poly_mix, po2)
    assert np.array_equal(check.view(), want)


@pytest.mark.parametrize("po2", [9, 13])
def test_prove_recursion_bit_exact(hal, po2):
    ctrl, data, accum, glob = O.synthetic_witness_recursion(po2)
    want_seal, want_roots, want_qpos = O.prove_recursion(po2, ctrl, data, accum, glob)
    seal, roots, qpos = SegmentProver(hal).prove(po2, ctrl, data, accum, glob, circuit="recursion")
    assert np.array_equal(roots, want_roots)
    assert np.array_equal(qpos, want_qpos)
    assert np.array_equal(seal, want_seal)


def test_prove_recursion_po2_18_properties(hal):
    # the size every lift / join / resolve proof has (recursion programs run at po2 = 18)
    po2 = 18
    ctrl, data, accum, glob = O.synthetic_witness_recursion(po2)
    seal, roots, qpos = SegmentProver(hal).prove(po2, ctrl, data, accum, glob, circuit="recursion")
    vroots = O.verify_recursion(seal)
    assert np.array_equal(vroots, roots)
    # seal-size formula: hdr 33 (32 globals + po2, no version word), G = 3, 643 taps, 163 columns
    rows, size = [], 1 << po2
    while size > 256:
        rows.append(size * 4 // 16)
        size //= 16
    R = len(rows)
    expect = 33 + 3 * 256 + 256 + 4 * (643 + 16) + R * 256 + 4 * size + 50 * (
        163 + 16 + 4 * 8 * (po2 + 2 - 5) + R * 64 + 8 * sum(int(np.log2(r)) - 5 for r in rows))
    assert len(seal) == expect
