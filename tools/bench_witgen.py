#!/usr/bin/env python3
"""witgen / accum phase times of one prove_core (CUDA events around each phase):
    [R0B200_LIB=risc0_b200/lib/libr0b200_<variant>.so] python tools/bench_witgen.py [--po2 20] [--iters 3]"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench  # noqa: E402
from risc0_b200 import B200Hal, SegmentProver  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--po2", type=int, default=20)
    ap.add_argument("--iters", type=int, default=3)
    a = ap.parse_args()
    pf = bench.build_segment(a.po2)
    hal = B200Hal(0)
    prover = SegmentProver(hal)
    seg = prover.upload_segment(pf)
    prover.prove_segment(seg, free=False)
    hal.profile_begin()
    for _ in range(a.iters):
        prover.prove_segment(seg, free=False)
    ph = hal.profile_end()
    out = {k: round(ph[k]["ms"] / a.iters, 3) for k in ("witgen", "accum")}
    out["step_ms"] = round(sum(v["ms"] for k, v in ph.items() if not k.startswith("eval_check_p")) / a.iters, 2)
    out["lib"] = os.path.basename(os.environ.get("R0B200_LIB", "default"))
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
