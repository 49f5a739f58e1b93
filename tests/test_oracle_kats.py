"""Pins the oracle (CPU restatement) against every golden vector the reference's own tests hold for the hot path
(SURVEY §8c). CPU only."""
import json
import os

import numpy as np
import pytest

import oracle_lib as O

HERE = os.path.dirname(os.path.abspath(__file__))
KATS = json.load(open(os.path.join(HERE, "golden", "kats.json")))
P = O.P


def test_fp_pow_kat():
    # baby_bear.rs:893-894
    L = O.lib()
    assert L.orc_fp_decode(L.orc_fp_pow(L.orc_fp_encode(5), 1000)) == KATS["fp_5_pow_1000"]


def test_fp_vs_u64_arithmetic():
    # baby_bear.rs:903-915 (compare against plain modular arithmetic)
    L = O.lib()
    rng = np.random.default_rng(2)
    for _ in range(2000):
        a, b = (int(x) for x in rng.integers(0, P, size=2))
        ea, eb = L.orc_fp_encode(a), L.orc_fp_encode(b)
        assert L.orc_fp_decode(L.orc_fp_add(ea, eb)) == (a + b) % P
        assert L.orc_fp_decode(L.orc_fp_sub(ea, eb)) == (a - b) % P
        assert L.orc_fp_decode(L.orc_fp_mul(ea, eb)) == a * b % P
    assert L.orc_fp_inv(0) == 0
    assert L.orc_fp_decode(L.orc_fp_mul(L.orc_fp_inv(L.orc_fp_encode(12345)), L.orc_fp_encode(12345))) == 1
    # numpy encode/decode helpers agree with the oracle
    xs = rng.integers(0, P, size=100, dtype=np.uint64)
    assert [int(v) for v in O.encode(xs)] == [L.orc_fp_encode(int(x)) for x in xs]
    assert np.array_equal(O.decode(O.encode(xs)), xs.astype(np.uint32))


def test_fpext_linear_kat():
    # baby_bear.rs:815-853
    k = KATS["fpext_linear"]
    L = O.lib()
    x, c0, c1 = O.encode(k["x"]), O.encode(k["c0"]), O.encode(k["c1"])
    out = np.zeros(4, dtype=np.uint32)
    L.orc_fpext_mul(O.ptr(out), O.ptr(x), O.ptr(c1))
    assert list(O.decode(out)) == k["x_mul_c1"]
    s = (O.decode(out).astype(np.uint64) + np.array(k["c0"], dtype=np.uint64)) % P
    assert list(s) == k["c0_plus_x_mul_c1"]
    inv = np.zeros(4, dtype=np.uint32)
    L.orc_fpext_inv(O.ptr(inv), O.ptr(x))
    L.orc_fpext_mul(O.ptr(out), O.ptr(inv), O.ptr(x))
    assert list(O.decode(out)) == [1, 0, 0, 0]


def test_rou_tables_self_check():
    # baby_bear.rs:184-199: ROU_FWD[k]^(2^k) = 1, primitive, and ROU_REV is its inverse
    L = O.lib()
    one = L.orc_fp_encode(1)
    for k in range(28):
        f, r = L.orc_rou_fwd(k), L.orc_rou_rev(k)
        assert L.orc_fp_pow(f, 1 << k) == one
        if k:
            assert L.orc_fp_pow(f, 1 << (k - 1)) != one
        assert L.orc_fp_mul(f, r) == one


def test_poseidon2_permutation_kat():
    # poseidon2/mod.rs:330-351
    k = KATS["poseidon2_perm"]
    out = O.poseidon2_mix(O.encode(k["input"]))
    assert list(O.decode(out)) == k["output"]


@pytest.mark.parametrize("name", ["poseidon2_hash32", "poseidon2_hash17"])
def test_poseidon2_hash_kats(name):
    # poseidon2/mod.rs:354-401 (aligned and unaligned sponge)
    k = KATS[name]
    d = O.hash_elems(O.POSEIDON2, O.encode(k["input"]))
    assert list(O.decode(d)) == k["digest_normal_form"]


def test_poseidon2_rng_kat():
    # prove/merkle.rs:161-172
    k = KATS["poseidon2_rng"]
    r = O.Rng(O.POSEIDON2)
    r.mix(np.zeros(8, dtype=np.uint32))
    x = int(O.decode(r.elem()))
    assert x == k["after_commit_zero"]
    r.mix(np.array([x, 2, 3, 4, 5, 6, 7, 8], dtype=np.uint32))  # Digest words are used raw
    assert int(O.decode(r.elem())) == k["after_commit_x2345678"]


def test_sha_hash_rows_kat():
    # hal/cpu.rs:726-733: 1 row x 16 zero columns
    d = O.hash_rows(O.SHA256, np.zeros(16, dtype=np.uint32), 1)
    assert d.tobytes().hex() == KATS["sha_hash_rows_1x16_zero"]


def test_sha_standard_vectors():
    # core/hash/sha/mod.rs:378-387
    out = np.zeros(8, dtype=np.uint32)
    O.lib().orc_sha_hash_bytes(O.ptr(out), b"abc", 3)
    assert out.tobytes().hex() == "ba7816bf8f01cfea414140de5dae2223b00361a396177a9cb410ff61f20015ad"
    O.lib().orc_sha_hash_bytes(O.ptr(out), b"", 0)
    assert out.tobytes().hex() == "e3b0c44298fc1c149afbf4c8996fb92427ae41e4649b934ca495991b7852b855"


def test_merkle_params_kats():
    # zkp/src/merkle.rs:73-102
    import ctypes as C
    for k in KATS["merkle_params"]:
        a, b, c = C.c_uint64(), C.c_uint64(), C.c_uint64()
        O.lib().orc_merkle_params(C.c_uint64(k["rows"]), C.c_uint64(k["cols"]), C.c_uint64(k["queries"]), C.byref(a),
                                  C.byref(b), C.byref(c))
        assert (a.value, b.value, c.value) == (k["layers"], k["top_layer"], k["top_size"])


def test_golden_seal_accepted():
    # verify/mod.rs:713-726: the stored Poseidon2 STARK for HelloCircuit at po2=12 must verify, all words consumed
    raw = np.fromfile(os.path.join(HERE, "golden", "proof.bin"), dtype="<u4")
    assert len(raw) == 22930
    roots = O.verify_hello(raw, 12)
    assert roots.shape == (5, 8)  # 3 groups + check + 1 FRI round


def test_golden_seal_tamper_rejected():
    raw = np.fromfile(os.path.join(HERE, "golden", "proof.bin"), dtype="<u4").copy()
    for pos in (5, 300, 5000, len(raw) - 3):
        bad = raw.copy()
        bad[pos] ^= 1
        with pytest.raises(RuntimeError):
            O.verify_hello(bad, 12)
    with pytest.raises(RuntimeError):
        O.verify_hello(raw[:-1], 12)
