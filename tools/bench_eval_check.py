#!/usr/bin/env python3
"""eval_check micro-benchmark: per-part device time of the generated kernels (CUDA events on the launching stream).

    [R0B200_LIB=risc0_b200/lib/libr0b200_<variant>.so] python tools/bench_eval_check.py [--po2 18] [--iters 5]
"""
import argparse
import json
import os
import sys

import numpy as np

if "R0B200_EVAL_TILED" not in os.environ:
    os.environ.setdefault("R0B200_PROFILE_PARTS", "1")

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from risc0_b200 import B200Hal  # noqa: E402

P = 15 * 2**27 + 1


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--po2", type=int, default=18)
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--tile-data", action="store_true", help="one random column repeated (fast host setup; timing only)")
    a = ap.parse_args()
    hal = B200Hal(0)
    n = 1 << a.po2
    domain = 4 * n
    rng = np.random.default_rng(5)

    def rand(k):
        return (rng.integers(0, P, size=k, dtype=np.uint64) * (2**32 % P) % P).astype(np.uint32)

    if a.tile_data:
        col = rand(domain)
        accum, data = hal.copy_from_elem("accum", np.tile(col, 103)), hal.copy_from_elem("data", np.tile(col[::-1], 211))
    else:
        accum, data = hal.copy_from_elem("accum", rand(103 * domain)), hal.copy_from_elem("data", rand(211 * domain))
    code = hal.alloc_elem_init("code", domain, 0)
    mix, out = hal.copy_from_elem("mix", rand(36)), hal.copy_from_elem("out", rand(90))
    check = hal.alloc_elem("check", 4 * domain)
    pm = rand(4)
    for _ in range(2):
        hal.eval_check_rv32im(check, [accum, code, data], [mix, out], pm, a.po2, n)
    hal.profile_begin()
    for _ in range(a.iters):
        hal.eval_check_rv32im(check, [accum, code, data], [mix, out], pm, a.po2, n)
    ph = hal.profile_end()
    res = {k: round(v["ms"] / a.iters, 3) for k, v in sorted(ph.items())}
    res["lib"] = os.environ.get("R0B200_LIB", "default")
    res["po2"] = a.po2
    res["ns_per_point"] = round(ph["eval_check"]["ms"] / a.iters * 1e6 / domain, 2)
    import zlib
    res["check_crc32"] = zlib.crc32(check.view().tobytes())
    res["mode"] = {k: os.environ[k] for k in ("R0B200_EVAL_TILED", "R0B200_EVAL_STREAMS", "R0B200_EVAL_AHEAD") if k in os.environ}
    print(json.dumps(res), flush=True)


if __name__ == "__main__":
    main()
