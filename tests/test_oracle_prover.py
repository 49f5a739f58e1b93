"""Self-consistency of the oracle: NTT contracts, prover <-> verifier round trips (the verifier is pinned by the
reference's golden seal in test_oracle_kats.py). CPU only."""
import numpy as np
import pytest

import oracle_lib as O

P = O.P


def naive_eval(coeffs_normal, x):
    acc = 0
    for c in reversed(coeffs_normal):
        acc = (acc * x + int(c)) % P
    return acc


def test_ntt_contracts_small():
    # SURVEY Appendix A closed-form buffer contracts, n = 64 (ntt.rs:345-434 uses the same naive check)
    L = O.lib()
    rng = np.random.default_rng(1)
    n, bits = 64, 6
    vals = rng.integers(0, P, size=n, dtype=np.uint64)
    w = int(O.decode(L.orc_rou_fwd(bits)))
    co = O.batch_interpolate_ntt(O.encode(vals), 1)
    co_nat = O.decode(O.batch_bit_reverse(co, 1))
    for k in (0, 1, 5, 63):
        assert naive_eval(co_nat, pow(w, k, P)) == int(vals[k])
    # expand x4 + evaluate over 4n (no zk shift): evaluated[i] = f(w4n^i)
    ev = O.decode(O.batch_expand_into_evaluate_ntt(co, 1, 2))
    w4 = int(O.decode(L.orc_rou_fwd(bits + 2)))
    for i in (0, 1, 2, 3, 77, 255):
        assert int(ev[i]) == naive_eval(co_nat, pow(w4, i, P))
    # with zk shift: f(3 * w4n^i)
    ev3 = O.decode(O.batch_expand_into_evaluate_ntt(O.zk_shift(co, 1), 1, 2))
    for i in (0, 9, 200):
        assert int(ev3[i]) == naive_eval(co_nat, 3 * pow(w4, i, P) % P)


def test_ntt_roundtrip_batched():
    rng = np.random.default_rng(3)
    n, c = 1024, 5
    vals = O.rand_elems(rng, n * c)
    co = O.batch_interpolate_ntt(vals, c)
    # forward NTT without expansion = expand_bits 0 on bit-reversed coefficients gives the values back
    back = O.batch_expand_into_evaluate_ntt(co, c, 0)
    assert np.array_equal(back, vals)


def hello_witness(po2, seed=7):
    rng = np.random.default_rng(seed)
    n = 1 << po2
    accum = np.zeros(n, dtype=np.uint32)
    code = np.zeros(n, dtype=np.uint32)
    data = O.encode(rng.integers(0, 2, size=n, dtype=np.uint64))
    return accum, code, data


@pytest.mark.parametrize("po2", [9, 12])
def test_hello_prove_verify(po2):
    accum, code, data = hello_witness(po2)
    seal = O.prove_hello(po2, accum, code, data)
    if po2 == 12:
        assert len(seal) == 22930  # same size as the reference's golden seal for this circuit/po2
    O.verify_hello(seal, po2)
    bad = seal.copy()
    bad[len(bad) // 2] ^= 4
    with pytest.raises(RuntimeError):
        O.verify_hello(bad, po2)


def test_hello_bad_witness_rejected():
    # a witness violating u2*(u2-1)=0 gives a check polynomial that is not low degree -> verifier must reject
    accum, code, data = hello_witness(9)
    data[5] = O.encode([2])[0]
    seal = O.prove_hello(9, accum, code, data)
    with pytest.raises(RuntimeError):
        O.verify_hello(seal, 9)


@pytest.mark.skipif(not O.have_ref(), reason="oracle/_ref not built")
def test_rv32im_prove_verify_small():
    po2 = 9
    code, data, accum, glob = O.synthetic_witness(po2)
    seal, roots, qpos = O.prove_rv32im(po2, code, data, accum, glob)
    # seal-size formula, SURVEY Appendix A
    G, taps, R = 3, 790, 1
    rows_r = [(4 << po2) // 16]
    final_words = 4 * ((1 << po2) // 16)  # one fold: 512 -> 32 coefficients (1024 words once the final degree is 256)
    expect = 92 + G * 256 + 256 + 4 * (taps + 16) + R * 256 + final_words + 50 * (
        315 + 16 + (G + 1) * 8 * (po2 + 2 - 5) + R * 64 + 8 * sum(int(np.log2(r)) - 5 for r in rows_r))
    assert len(seal) == expect
    assert seal[0] == 2
    vroots = O.verify_rv32im(seal)
    assert np.array_equal(vroots, roots)
    bad = seal.copy()
    bad[2000] ^= 1
    with pytest.raises(RuntimeError):
        O.verify_rv32im(bad)
