"""End to end through the public surface: build a small RV32IM guest, run it through the executor restatement
(risc0_b200.preflight: the reference's execute/ + prove/witgen/preflight.rs), and prove every segment of the session on
the GPUs of this node with the r0vm-style scheduler (risc0_b200.scheduler: preflight on host threads -> bounded queue ->
one worker per device, upload of segment s + 1 overlapping the proof of segment s). Everything below the scheduler is the
C ABI of include/r0b200.h (r0b200_segment_upload / r0b200_prove_segment).

    python examples/prove_guest.py [--guest loop|sha2|bigint] [--po2 14] [--gpus 1]

Needs a B200 (the library has no CPU fallback). The seals are the STARK seals of the reference's `prove_core`; checking
them is the verifier's business (tests/ do it with the restated verifier, including the constraint check).
"""
import argparse
import os
import sys
import time

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)

from risc0_b200 import preflight as PF  # noqa: E402
from risc0_b200.scheduler import b200_scheduler  # noqa: E402


def build_guest(kind):
    if kind == "loop":       # execute/testutil.rs kernel::simple_loop
        return PF.simple_loop_kernel(20000)
    if kind == "sha2":       # SHA-256 of a 575-byte message through the sha2 accelerator, 12 times
        return PF.sha2_guest(bytes(i & 0xff for i in range(575)), repeat=12)
    if kind == "bigint":     # a * b mod n on the secp256k1 prime through the reference's modmul_256 bigint2 program
        blob = open(os.path.join(ROOT, "tests", "golden", "bigint_modmul_256.blob"), "rb").read()
        n = 0xfffffffffffffffffffffffffffffffffffffffffffffffffffffffefffffc2f
        return PF.bigint_modmul_guest(blob, 0x1234567890abcdef << 128 | 0xfedcba9876543210, 3 << 200 | 12345, n)
    raise SystemExit("unknown guest %r" % kind)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--guest", default="loop")
    ap.add_argument("--po2", type=int, default=14, help="segment size limit (2^po2 cycles)")
    ap.add_argument("--gpus", type=int, default=1)
    a = ap.parse_args()

    t0 = time.perf_counter()
    segments = PF.execute(build_guest(a.guest), segment_po2=a.po2)
    t1 = time.perf_counter()
    print("executed: %d segment(s), %d user cycles, exit %r  (%.1f s)" % (
        len(segments), sum(s.suspend_cycle for s in segments), segments[-1].terminate_state, t1 - t0))

    results = b200_scheduler(list(range(a.gpus)), rand_z=(1, 2, 3, 4)).run(segments)
    t2 = time.perf_counter()
    for seg, r in zip(segments, results):
        print("segment %d: po2 %d, %d user cycles -> seal of %d words on device %d  (preflight %.2f s, prove %.3f s)" % (
            r.index, seg.po2, seg.suspend_cycle, len(r.seal), r.device, r.t_preflight, r.t_prove))
    print("proved the session in %.1f s of wall clock (host-side preflight is plain Python here)" % (t2 - t1))


if __name__ == "__main__":
    main()
