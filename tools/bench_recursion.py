#!/usr/bin/env python3
"""Recursion-circuit proof throughput (BASELINE config 4: every lift / join / resolve is one po2 = 18 recursion proof).
Synthetic witness, device-resident, CUDA-event timing on the prover's stream.

    python tools/bench_recursion.py [--po2 18] [--iters 5]
"""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from risc0_b200 import B200Hal, SegmentProver  # noqa: E402

P = 15 * 2**27 + 1


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--po2", type=int, default=18)
    ap.add_argument("--iters", type=int, default=5)
    a = ap.parse_args()
    hal = B200Hal(0)
    n = 1 << a.po2
    rng = np.random.Generator(np.random.PCG64(0x5EED1000 + a.po2))

    def rand(k):
        return (rng.integers(0, P, size=k, dtype=np.uint64) * (2**32 % P) % P).astype(np.uint32)

    ctrl, data, accum = (hal.copy_from_elem("w", rand(c * n)) for c in (23, 128, 12))
    glob = rand(32)
    prover = SegmentProver(hal)
    for _ in range(2):
        seal, _, _ = prover.prove(a.po2, ctrl, data, accum, glob, circuit="recursion")
    hal.profile_begin()
    hal.timer_start()
    for _ in range(a.iters):
        seal, _, _ = prover.prove(a.po2, ctrl, data, accum, glob, circuit="recursion")
    ms = hal.timer_stop() / a.iters
    ph = hal.profile_end()
    out = {"workload": "recursion proof po2=%d (lift/join shape), synthetic witness" % a.po2, "ms_per_proof": round(ms, 3),
           "proofs_per_s": round(1e3 / ms, 2), "seal_words": int(len(seal)),
           "phase_ms": {k: round(v["ms"] / a.iters, 3) for k, v in sorted(ph.items(), key=lambda kv: -kv[1]["ms"])}}
    print(json.dumps(out), flush=True)
    hal.close()


if __name__ == "__main__":
    main()
