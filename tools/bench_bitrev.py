#!/usr/bin/env python3
"""batch_bit_reverse: TMA kernel (bitrev_tma.cu) against the register-path kernel (R0B200_BITREV_TMA=0), same process
not possible (the switch is read once): run twice. python tools/bench_bitrev.py [--shapes 20x64,22x16,24x4,16x256]"""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from risc0_b200 import B200Hal  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shapes", default="16x256,18x211,20x64,20x211,22x16,22x64,24x4,24x16")
    a = ap.parse_args()
    hal = B200Hal(0)
    out = {"tma": os.environ.get("R0B200_BITREV_TMA", "1") != "0"}
    for sh in a.shapes.split(","):
        lg, c = (int(v) for v in sh.split("x"))
        n = 1 << lg
        x = hal.copy_from_elem("x", (np.arange(n * c, dtype=np.uint64) % 2013265921).astype(np.uint32))
        for _ in range(2):
            hal.batch_bit_reverse(x, c)
        hal.sync()
        hal.timer_start()
        iters = 10
        for _ in range(iters):
            hal.batch_bit_reverse(x, c)
        ms = hal.timer_stop() / iters
        # 12 reversals in total = identity: check it, and one reversal against numpy on the first column
        ok = bool(np.array_equal(x.view()[:n], (np.arange(n, dtype=np.uint64) % 2013265921).astype(np.uint32)))
        hal.batch_bit_reverse(x, c)
        col = x.view()[(c - 1) * n:c * n]
        idx = np.arange(n, dtype=np.uint64)
        rev = np.zeros(n, dtype=np.uint64)
        for bit in range(lg):
            rev |= ((idx >> np.uint64(bit)) & np.uint64(1)) << np.uint64(lg - 1 - bit)
        ok = ok and bool(np.array_equal(col, (((c - 1) * n + rev) % 2013265921).astype(np.uint32)))
        out[sh] = {"ms": round(ms, 4), "GBps": round(8 * n * c / ms / 1e6, 1), "ok": ok}
        x.free() if hasattr(x, "free") else None
    print(json.dumps(out), flush=True)
    hal.close()


if __name__ == "__main__":
    main()
