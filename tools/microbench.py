#!/usr/bin/env python3
"""Config-5 style microbench sweep: per-op device time (CUDA events on the launching stream, 3 warm-ups,
median of N) and achieved algorithmic GB/s. Inputs are larger than the 126 MB L2 wherever the op allows it.

    python tools/microbench.py [--quick] [--json out.json]
"""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from risc0_b200 import B200Hal  # noqa: E402

P = 15 * 2**27 + 1


def rand_dev(hal, n, seed):
    rng = np.random.default_rng(seed)
    return hal.copy_from_elem("x", (rng.integers(0, P, size=n, dtype=np.uint64) * (2**32 % P) % P).astype(np.uint32))


def timeit(hal, fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(iters):
        hal.timer_start()
        fn()
        ts.append(hal.timer_stop())
    return float(np.median(ts))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--json", default=None)
    a = ap.parse_args()
    hal = B200Hal(0)
    peak = None
    try:
        peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        peak = 6650.0
    rows = []

    def rec(op, shape, ms, bytes_):
        gbs = bytes_ / ms / 1e6
        rows.append(dict(op=op, shape=shape, ms=round(ms, 4), alg_GB=round(bytes_ / 1e9, 4), GBps=round(gbs, 1),
                         frac_of_hbm_peak=round(gbs / peak, 3)))
        print("%-34s %-18s %9.3f ms %9.1f GB/s  (%.2f of %.0f)" % (op, shape, ms, gbs, gbs / peak, peak), flush=True)

    ntt_shapes = [(16, 1), (16, 256), (18, 64), (20, 4), (20, 64), (20, 211), (22, 16), (24, 4)] if not a.quick else [(20, 32)]
    for lg, c in ntt_shapes:
        n = 1 << lg
        x = rand_dev(hal, n * c, 1)
        rec("batch_interpolate_ntt", "2^%d x %d" % (lg, c), timeit(hal, lambda: hal.batch_interpolate_ntt(x, c)), 8 * n * c)
        rec("batch_interpolate_ntt_zk", "2^%d x %d" % (lg, c), timeit(hal, lambda: hal.batch_interpolate_ntt_zk(x, c)), 8 * n * c)
        rec("batch_bit_reverse", "2^%d x %d" % (lg, c), timeit(hal, lambda: hal.batch_bit_reverse(x, c)), 8 * n * c)
        y = None
        if lg + 2 <= 24:  # 2^24 is the largest domain (MAX_CYCLES_PO2 = 22)
            y = hal.alloc_elem("y", 4 * n * c)
            rec("batch_expand_into_evaluate_ntt", "2^%d x %d" % (lg, c),
                timeit(hal, lambda: hal.batch_expand_into_evaluate_ntt(y, x, c, 2)), 20 * n * c)
        cp = hal.alloc_elem("cp", n * c)
        rec("eltwise_copy_elem", "2^%d x %d" % (lg, c), timeit(hal, lambda: hal.eltwise_copy_elem(cp, x)), 8 * n * c)
        del x, y, cp
    hash_shapes = [(16, 64), (18, 64), (20, 103), (22, 1), (22, 16), (22, 64), (22, 211), (24, 16)] if not a.quick else [(22, 16)]
    for lg, c in hash_shapes:
        r = 1 << lg
        m = rand_dev(hal, r * c, 2)
        out = hal.alloc_digest("d", r)
        ms = timeit(hal, lambda: hal.hash_rows(out, m), iters=5)
        perms = r * ((c + 15) // 16)
        rec("hash_rows(poseidon2)", "2^%d x %d" % (lg, c), ms, 4 * r * c + 32 * r)
        rows[-1]["Gperm_per_s"] = round(perms / ms / 1e6, 3)
        rows[-1]["Tmodmul_per_s"] = round(perms * 1356 / ms / 1e9, 3)
        print("    -> %.3f Gperm/s, %.2f T modmul/s" % (perms / ms / 1e6, perms * 1356 / ms / 1e9), flush=True)
        del m, out
    for lg in ([22, 18] if not a.quick else [22]):
        r = 1 << lg
        nodes = rand_dev(hal, 16 * r, 3)
        nodes = hal.copy_from_digest("nodes", nodes.view())
        m = hal.alloc_elem("m", 0)

        def fold_all():
            size = r
            while size > 1:
                hal.hash_fold(nodes, size, size // 2)
                size //= 2
        ms = timeit(hal, fold_all, iters=5)
        rec("hash_fold x levels (per-level)", "2^%d leaves" % lg, ms, 96 * (r - 1))
        rows[-1]["Gperm_per_s"] = round((r - 1) / ms / 1e6, 3)
        del nodes
    # FRI fold / mix / evaluate at po2=20 shapes
    n = 1 << 20
    x = rand_dev(hal, 4 * n, 4)
    o = hal.alloc_elem("o", 4 * n // 16)
    mix = np.array([5, 6, 7, 8], dtype=np.uint32)
    rec("fri_fold", "4 x 2^20", timeit(hal, lambda: hal.fri_fold(o, x, mix)), (256 + 16) * (n // 16))
    S = 211 if not a.quick else 32
    inp = rand_dev(hal, S * n, 5)
    combos = (np.arange(S) % 4).astype(np.uint32)
    out = hal.alloc_extelem_zeroed("mix", 5 * n)
    rec("mix_poly_coeffs", "%d x 2^20 -> 4" % S, timeit(hal, lambda: hal.mix_poly_coeffs(out, mix, mix, inp, combos, S, n), iters=5),
        4 * S * n + 2 * 16 * 4 * n)
    E = 670 if not a.quick else 64
    which = hal.copy_from_u32("w", (np.arange(E) % S).astype(np.uint32))
    xs = rand_dev(hal, 4 * E, 6)
    xs = hal.copy_from_extelem("xs", xs.view())
    ev = hal.alloc_extelem("ev", E)
    rec("batch_evaluate_any", "%d evals of 2^20" % E, timeit(hal, lambda: hal.batch_evaluate_any(inp, S, which, xs, ev), iters=5),
        4 * E * n)
    comb = hal.alloc_extelem_zeroed("c", 5 * n)
    fin = hal.alloc_elem("f", 4 * n)
    rec("eltwise_sum_extelem", "5 x 2^20", timeit(hal, lambda: hal.eltwise_sum_extelem(fin, comb)), 16 * 5 * n + 16 * n)
    z = [np.array([3, 4, 5, 6], dtype=np.uint32)]

    def div():
        try:
            hal.combos_divide(comb.slice(0, n), [(0, z)], n)
        except RuntimeError:
            pass
    rec("combos_divide (1 division)", "2^20 ext", timeit(hal, div, iters=5), 32 * n)
    print("launches so far:", hal.launch_count())
    if a.json:
        json.dump(rows, open(a.json, "w"), indent=1)
    hal.close()


if __name__ == "__main__":
    main()
