#!/usr/bin/env python3
"""ncu target: ONE whole prove_core of a po2-sized loop-guest segment from its preflight trace (no warm-up proof, so the
launch list is exactly one step). python tools/profile_target.py [--po2 20]"""
import argparse
import os
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from risc0_b200 import B200Hal, SegmentProver  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--po2", type=int, default=20)
    a = ap.parse_args()
    pf = bench.build_segment(a.po2)
    hal = B200Hal(0)
    prover = SegmentProver(hal)
    seg = prover.upload_segment(pf)
    hal.sync()
    seal = prover.prove_segment(seg)[0]
    print("seal words", len(seal), "launches", hal.launch_count(), flush=True)
    hal.close()


if __name__ == "__main__":
    main()
