"""GPU parity of the witness path: the device witness generator / accumulator (csrc/witgen*.cu, generated step
functions) against the REFERENCE'S OWN compiled C++ witgen on the same preflight trace, word for word; then the whole
prove_core from a preflight trace (everything on the device) against the CPU prover fed with the reference witness.
Traces come from risc0_b200.preflight (tests/test_preflight.py pins that on the CPU)."""
import numpy as np
import pytest

import oracle_lib as O
import witgen_ref as W
from risc0_b200 import B200Hal, SegmentProver, WitnessGenerator
from risc0_b200 import preflight as PF
from test_preflight import all_insn_guest, host_write_guest, modmul_guest

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not (W.have_ref() and O.have_ref()), reason="oracle/_ref not built")]


@pytest.fixture(scope="module")
def hal():
    h = B200Hal(0, "poseidon2")
    yield h
    h.close()


def segments():
    out = {"loop_po2_13": PF.execute(PF.simple_loop_kernel(200), segment_po2=14)[0],
           "all_insn": PF.execute(all_insn_guest(), segment_po2=14)[0]}
    out["user_mode"] = PF.execute(PF.user_mode_guest(30), segment_po2=14)[0]
    out["sha2"] = PF.execute(PF.sha2_guest(bytes(range(200))), segment_po2=14)[0]   # 4 blocks through the sha2 ecall
    out["bigint"] = PF.execute(modmul_guest(2)[0], segment_po2=14)[0]               # modmul_256 through the bigint ecall
    out["host_write"] = PF.execute(host_write_guest(), segment_po2=14)[0]
    out["user_sha2"] = PF.execute(PF.user_sha2_via_kernel_guest(bytes(range(70))), segment_po2=14)[0]   # user -> kernel -> sha2
    rng = np.random.default_rng(77)                                                  # guest-invoked poseidon2 ecall with state
    out["p2_ecall"] = PF.execute(PF.poseidon2_ecall_guest([int(x) for x in rng.integers(0, 1 << 32, 24)], 0,
                                                          [int(x) for x in rng.integers(0, PF.P, 8)]), segment_po2=14)[0]
    split = PF.execute(PF.simple_loop_kernel(4000), segment_po2=13)
    out["split_first"], out["split_second"] = split[0], split[1]
    return out


SEGS = None


def seg(name):
    global SEGS
    if SEGS is None:
        SEGS = segments()
    return SEGS[name]


@pytest.mark.parametrize("name", ["loop_po2_13", "all_insn", "split_first", "split_second", "user_mode", "sha2", "bigint",
                                  "p2_ecall", "host_write", "user_sha2"])
def test_device_witgen_and_accum_match_reference(hal, name):
    # (host_write: with the write record indexed as the executor wrote it, see tests/test_preflight.py::test_host_write_guest)
    pf = PF.PreflightResults(seg(name), (11, 12, 13, 14), write_record_off_by_one=name != "host_write")
    want_glob, want_data = W.ref_generate_witness(pf)
    wg = WitnessGenerator(hal, pf)
    assert np.array_equal(wg.global_.view(), want_glob)
    assert np.array_equal(wg.data.view(), want_data)
    assert not wg.code.view().any()
    rng = np.random.default_rng(5)
    mix = O.rand_elems(rng, 36)
    wg.accum(mix)
    assert np.array_equal(wg.accum_buf.view(), W.ref_accum(pf, want_glob, want_data, mix))


def test_device_witgen_reports_bad_traces(hal):
    pf = PF.PreflightResults(seg("loop_po2_13"), (1, 2, 3, 4))
    bad = PF.PreflightResults(seg("loop_po2_13"), (1, 2, 3, 4))
    bad.txns = bad.txns.copy()
    bad.txns["addr"][len(bad.txns) // 2] ^= 4        # "memory peek not in preflight"
    with pytest.raises(Exception) as ei:
        WitnessGenerator(hal, bad)
    assert "witgen" in str(ei.value)
    # (the reference's witgen throws "memory peek not in preflight" for this trace too, but from inside its thread
    # pool, which takes the process down - so it is not called here)
    WitnessGenerator(hal, pf)                        # the context is still usable afterwards


@pytest.mark.parametrize("name", ["loop_po2_13", "all_insn", "split_first", "user_mode", "sha2", "bigint"])
def test_prove_core_from_trace_bit_exact_and_valid(hal, name):
    pf = PF.PreflightResults(seg(name), (21, 22, 23, 24))
    po2 = pf.po2
    seal, roots, qpos, glob = SegmentProver(hal).prove_core(pf)
    # CPU side: reference witgen -> oracle transcript up to the mix -> reference accum from THAT mix -> oracle prover
    want_glob, data = W.ref_generate_witness(pf)
    code = np.zeros(pf.rows, dtype=np.uint32)
    mix = O.prove_rv32im_mix(po2, code, data, want_glob)
    accum = W.ref_accum(pf, want_glob, data, mix)
    want_seal, want_roots, want_qpos = O.prove_rv32im(po2, code, data, accum, want_glob)
    assert np.array_equal(glob, want_glob)
    assert np.array_equal(roots, want_roots) and np.array_equal(qpos, want_qpos)
    assert np.array_equal(seal, want_seal)
    # the restated verifier accepts it INCLUDING the constraint check check(z) * ((3z)^N - 1) == poly_ext(eval_u)
    vroots, validity_checked = O.verify_with_validity(seal)
    assert validity_checked and np.array_equal(vroots, roots)
    # and this is a witness the circuit accepts: every constraint vanishes on every row
    assert O.rv32im_check_constraints(accum, data, mix, want_glob, O.rand_ext(np.random.default_rng(1)), po2) == (0, None)


def test_two_phase_with_device_accum(hal):
    # the protocol-correct flow through the Hal-level API: begin -> mix -> step_accum on the device -> finish
    pf = PF.PreflightResults(seg("loop_po2_13"), (31, 32, 33, 34))
    wg = WitnessGenerator(hal, pf)
    glob = wg.global_.view()
    prover = SegmentProver(hal)
    h, mix = prover.begin(pf.po2, wg.code, wg.data, glob)
    wg.accum(mix)
    seal, roots, _ = prover.finish(h, wg.accum_buf)
    want = SegmentProver(hal).prove_core(pf)[0]
    assert np.array_equal(seal, want)


def test_bigint_segment_one_call_equals_two_phase(hal):
    # the one-call path (r0b200_segment_upload / r0b200_prove_segment) evaluates the mix-dependent BigIntAccum cells itself,
    # from the trace; the Hal-level two-phase route gets them from the host mirror (PreflightResults.bigint_accum_injector):
    # same seal, and the verifier accepts it including the constraint check
    pf = PF.PreflightResults(seg("bigint"), (51, 52, 53, 54))
    prover = SegmentProver(hal)
    one = prover.prove_segment(prover.upload_segment(pf))[0]
    two = prover.prove_core(pf, two_phase=True)[0]
    assert np.array_equal(one, two)
    assert O.verify_with_validity(one)[1]


def test_verifier_rejects_a_witness_that_violates_constraints(hal):
    # a synthetic (random) witness gives a seal whose Merkle / FRI / DEEP checks pass but whose constraint check fails:
    # the validity check is what distinguishes the two
    po2 = 10
    code, data, accum, glob = O.synthetic_witness(po2)
    seal, _, _ = SegmentProver(hal).prove(po2, code, data, accum, glob)
    O.verify_rv32im(seal)
    with pytest.raises(Exception) as ei:
        O.verify_with_validity(seal)
    assert "constraint" in str(ei.value)


def test_segment_upload_pipeline(hal):
    # r0b200_segment_upload / r0b200_prove_segment: two segments in flight (the upload of the second overlaps the proof
    # of the first), an uploaded segment can be proved twice, and the seals equal the one-call form's
    pfs = [PF.PreflightResults(seg(n), (41, 42, 43, 44)) for n in ("loop_po2_13", "all_insn")]
    prover = SegmentProver(hal)
    want = [prover.prove_core(pf)[0] for pf in pfs]
    s0 = prover.upload_segment(pfs[0])
    s1 = prover.upload_segment(pfs[1])
    got0 = prover.prove_segment(s0, free=False)[0]
    got1 = prover.prove_segment(s1)[0]
    again = prover.prove_segment(s0)[0]
    assert np.array_equal(got0, want[0]) and np.array_equal(got1, want[1]) and np.array_equal(again, want[0])


def test_two_contexts_prove_concurrently_on_one_gpu():
    # bench.py keeps two segments in flight per GPU: two contexts (own stream, copy stream and memory pool), one host
    # thread each, proving at the same time. Every seal must equal the one a single context produces.
    import threading
    pfs = [PF.PreflightResults(seg(n), (61, 62, 63, 64)) for n in ("loop_po2_13", "all_insn")]
    solo = B200Hal(0, "poseidon2")
    want = [SegmentProver(solo).prove_core(pf)[0] for pf in pfs]
    solo.close()
    hals = [B200Hal(0, "poseidon2") for _ in range(2)]
    got, errs = [[], []], []

    def worker(w):
        try:
            p = SegmentProver(hals[w])
            for r in range(4):
                pf = pfs[(w + r) % 2]
                up = p.upload_segment(pf)
                got[w].append(((w + r) % 2, p.prove_segment(up)[0]))
        except BaseException as e:   # noqa: BLE001
            errs.append(e)

    th = [threading.Thread(target=worker, args=(w,)) for w in range(2)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    for h in hals:
        h.close()
    assert not errs, errs
    for w in range(2):
        assert len(got[w]) == 4
        for which, seal in got[w]:
            assert np.array_equal(seal, want[which])


def test_reference_witgen_symbols(hal):
    # risc0_circuit_rv32im_cuda_witgen / _cuda_accum with the reference's structs (rv32im-sys/src/lib.rs:21-119),
    # device buffers + host trace, as rv32im/src/prove/hal/cuda.rs:60-157 calls them
    import ctypes as C
    from risc0_b200 import _lib
    from risc0_b200.hal import _trace_struct
    L = C.CDLL(_lib.LIB_PATH)
    L.risc0_circuit_rv32im_cuda_witgen.restype = C.c_char_p
    L.risc0_circuit_rv32im_cuda_accum.restype = C.c_char_p
    pf = PF.PreflightResults(seg("loop_po2_13"), (51, 52, 53, 54))
    want_glob, want_data = W.ref_generate_witness(pf)
    rows = pf.rows
    data = hal.alloc_elem_init("data", 211 * rows, 0xFFFFFFFF)
    hal.scatter(data, *pf.injector)
    glob = hal.copy_from_elem("global", pf.global_)
    hal.sync()
    st, keep = _trace_struct(pf)

    def rb(buf, r, c, checked=True):
        return W.RawBuffer(buf.ptr.value, r, c, checked)

    bufs = W.RawExecBuffers(rb(glob, 1, 90), rb(data, rows, 211))
    err = L.risc0_circuit_rv32im_cuda_witgen(C.c_uint32(0), C.byref(bufs), C.byref(st), C.c_uint32(rows))
    assert err is None, err
    hal.eltwise_zeroize_elem(glob)
    hal.eltwise_zeroize_elem(data)
    assert np.array_equal(data.view(), want_data) and np.array_equal(glob.view(), want_glob)
    mix_h = O.rand_elems(np.random.default_rng(6), 36)
    accum = hal.alloc_elem_init("accum", 103 * rows, 0xFFFFFFFF)
    mix = hal.copy_from_elem("mix", mix_h)
    hal.sync()
    ab = W.RawAccumBuffers(rb(data, rows, 211), rb(accum, rows, 103, False), rb(glob, 1, 90), rb(mix, 1, 36))
    err = L.risc0_circuit_rv32im_cuda_accum(C.byref(ab), C.byref(st), C.c_uint32(rows))
    assert err is None, err
    hal.eltwise_zeroize_elem(accum)
    assert np.array_equal(accum.view(), W.ref_accum(pf, want_glob, want_data, mix_h))


def test_scheduler_on_device(hal):
    # the r0vm-style scheduler over the real prover (one device here; tools/bench_schedule.py runs it over several):
    # three segments of one continuation, seals equal the direct proofs, results in submission order
    from risc0_b200.scheduler import b200_scheduler
    segs = PF.execute(PF.simple_loop_kernel(6000), segment_po2=13)[:3]
    assert len(segs) == 3
    rand_z = (61, 62, 63, 64)
    res = b200_scheduler([0], rand_z=rand_z).run(segs)
    prover = SegmentProver(hal)
    for s, r in zip(segs, res):
        assert np.array_equal(r.seal, prover.prove_core(PF.PreflightResults(s, rand_z))[0])
    assert [r.index for r in res] == [0, 1, 2]
