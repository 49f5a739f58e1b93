#!/usr/bin/env python3
"""NTT-only micro-benchmark (ncu target): python tools/bench_ntt.py [--lg 20] [--cols 32] [--iters 5]"""
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
    ap.add_argument("--lg", type=int, default=20)
    ap.add_argument("--cols", type=int, default=32)
    ap.add_argument("--iters", type=int, default=5)
    a = ap.parse_args()
    hal = B200Hal(0)
    n, c = 1 << a.lg, a.cols
    rng = np.random.default_rng(1)
    x = hal.copy_from_elem("x", (rng.integers(0, P, size=n * c, dtype=np.uint64) * (2**32 % P) % P).astype(np.uint32))
    y = hal.alloc_elem("y", 4 * n * c)
    for _ in range(2):
        hal.batch_interpolate_ntt_zk(x, c)
        hal.batch_expand_into_evaluate_ntt(y, x, c, 2)
        hal.batch_bit_reverse(x, c)
    hal.profile_begin()
    for _ in range(a.iters):
        hal.batch_interpolate_ntt_zk(x, c)
        hal.batch_expand_into_evaluate_ntt(y, x, c, 2)
        hal.batch_bit_reverse(x, c)
    ph = hal.profile_end()
    out = {k: {"ms": round(v["ms"] / a.iters, 4), "alg_GBps": round(v["bytes"] / v["ms"] / 1e6, 1)} for k, v in ph.items()}
    out["shape"] = "2^%d x %d" % (a.lg, c)
    print(json.dumps(out), flush=True)
    hal.close()


if __name__ == "__main__":
    main()
