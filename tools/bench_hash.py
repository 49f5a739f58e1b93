#!/usr/bin/env python3
"""Poseidon2 hash_rows / Merkle fold micro-benchmark (ncu target): python tools/bench_hash.py [--lg 22] [--cols 64]"""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from risc0_b200 import B200Hal  # noqa: E402

P = 15 * 2**27 + 1


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--lg", type=int, default=22)
    ap.add_argument("--cols", type=int, default=64)
    ap.add_argument("--iters", type=int, default=3)
    a = ap.parse_args()
    hal = B200Hal(0)
    rows, c = 1 << a.lg, a.cols
    rng = np.random.default_rng(1)
    m = hal.copy_from_elem("m", (rng.integers(0, P, size=rows * c, dtype=np.uint64) * (2**32 % P) % P).astype(np.uint32))
    nodes = hal.alloc_digest("nodes", 2 * rows)
    for _ in range(2):
        hal.merkle_build(nodes, m, rows, c)
    hal.profile_begin()
    for _ in range(a.iters):
        hal.merkle_build(nodes, m, rows, c)
    ph = hal.profile_end()
    perms_rows = rows * ((c + 15) // 16)
    out = {k: round(v["ms"] / a.iters, 4) for k, v in ph.items()}
    out["rows_Gperm_s"] = round(perms_rows / (ph["hash_rows"]["ms"] / a.iters) / 1e6, 3)
    out["fold_Gperm_s"] = round((rows - 1) / (ph["hash_fold"]["ms"] / a.iters) / 1e6, 3)
    out["shape"] = "2^%d x %d" % (a.lg, c)
    print(json.dumps(out), flush=True)
    hal.close()


if __name__ == "__main__":
    main()
