#!/usr/bin/env python3
"""One-off evidence run: the po2 = 20 (headline size) seal from the GPU path against the CPU oracle, word for word.
The oracle needs several minutes of all host cores at this size, so this is not part of the regular test-suite.

    python tools/check_po2_20_bit_exact.py [--po2 20]  -> prints a JSON line
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ.setdefault("OMP_NUM_THREADS", str(os.cpu_count() or 1))
import oracle_lib as O  # noqa: E402
from risc0_b200 import B200Hal, SegmentProver  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--po2", type=int, default=20)
    a = ap.parse_args()
    code, data, accum, glob = O.synthetic_witness(a.po2)
    hal = B200Hal(0)
    t0 = time.perf_counter()
    seal, roots, qpos = SegmentProver(hal).prove(a.po2, code, data, accum, glob)
    t_gpu = time.perf_counter() - t0
    t0 = time.perf_counter()
    want_seal, want_roots, want_qpos = O.prove_rv32im(a.po2, code, data, accum, glob)
    t_cpu = time.perf_counter() - t0
    res = {"po2": a.po2, "seal_words": int(len(seal)), "roots_equal": bool(np.array_equal(roots, want_roots)),
           "query_positions_equal": bool(np.array_equal(qpos, want_qpos)), "seal_equal": bool(np.array_equal(seal, want_seal)),
           "gpu_seconds_first_call": round(t_gpu, 3), "cpu_oracle_seconds": round(t_cpu, 1),
           "cpu_threads": O.lib().orc_num_threads()}
    print(json.dumps(res), flush=True)
    hal.close()
    sys.exit(0 if res["seal_equal"] and res["roots_equal"] else 1)


if __name__ == "__main__":
    main()
