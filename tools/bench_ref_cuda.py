#!/usr/bin/env python3
"""Reported baseline: the reference's own CUDA kernels (non-sppark subset, oracle/_ref/libref_zkp_cuda.so, built by
`make -C oracle refcuda`) timed next to ours on the same device buffers at po2 = 20 shapes. The reference entry points
create a stream, launch and synchronise per call (cuda.h:77-100), so they are timed with the host clock around the
call; ours with CUDA events on the context's stream. NTT / Poseidon2 / eval_check cannot be built offline (sppark).

    python tools/bench_ref_cuda.py
"""
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, ".."))
sys.path.insert(0, HERE)
import ref_cuda as R  # noqa: E402
from risc0_b200 import B200Hal  # noqa: E402

P = 15 * 2**27 + 1


def main():
    hal = B200Hal(0, "sha-256")
    rng = np.random.default_rng(9)

    def rand(k):
        return (rng.integers(0, P, size=k, dtype=np.uint64) * (2**32 % P) % P).astype(np.uint32)

    def ours(fn, iters=5):
        for _ in range(2):
            fn()
        hal.timer_start()
        for _ in range(iters):
            fn()
        return hal.timer_stop() / iters

    def theirs(fn, iters=5):
        hal.sync()
        for _ in range(2):
            fn()
        t0 = time.perf_counter()
        for _ in range(iters):
            fn()
        return (time.perf_counter() - t0) * 1e3 / iters

    rows = []
    n = 1 << 20
    x = hal.copy_from_elem("x", rand(64 * n))
    rows.append(("batch_bit_reverse 64 x 2^20", ours(lambda: hal.batch_bit_reverse(x, 64)), theirs(lambda: R.batch_bit_reverse(x, 20, 64 * n))))
    S = 211
    inp = hal.copy_from_elem("in", rand(S * n))
    combos = (np.arange(S) % 4).astype(np.uint32)
    out = hal.alloc_extelem_zeroed("out", 5 * n)
    ms, mx = rand(4), rand(4)
    d_c, d_ms, d_mx = hal.copy_from_u32("c", combos), hal.copy_from_extelem("ms", ms), hal.copy_from_extelem("mx", mx)
    rows.append(("mix_poly_coeffs 211 x 2^20", ours(lambda: hal.mix_poly_coeffs(out, ms, mx, inp, combos, S, n)),
                 theirs(lambda: R.mix_poly_coeffs(out, inp, d_c, d_ms, d_mx, S, n), iters=2)))
    E = 670
    which = hal.copy_from_u32("w", (np.arange(E) % S).astype(np.uint32))
    xs = hal.copy_from_extelem("xs", rand(4 * E))
    ev = hal.alloc_extelem("ev", E)
    rows.append(("batch_evaluate_any 670 x 2^20", ours(lambda: hal.batch_evaluate_any(inp, S, which, xs, ev)),
                 theirs(lambda: R.batch_evaluate_any(ev, inp, which, xs, E, n), iters=2)))
    cnt = n // 16
    fin = hal.copy_from_elem("fin", rand(4 * n))
    fo = hal.alloc_elem("fo", 4 * cnt)
    rows.append(("fri_fold 4 x 2^20", ours(lambda: hal.fri_fold(fo, fin, mx)), theirs(lambda: R.fri_fold(fo, fin, d_mx, cnt))))
    comb = hal.copy_from_extelem("comb", rand(4 * 5 * n))
    so = hal.alloc_elem("so", 4 * n)
    rows.append(("eltwise_sum_extelem 5 x 2^20", ours(lambda: hal.eltwise_sum_extelem(so, comb)), theirs(lambda: R.eltwise_sum_fpext(so, comb, 5, n))))
    r = 1 << 22
    m = hal.copy_from_elem("m", rand(16 * r))
    dg = hal.alloc_digest("dg", r)
    rows.append(("sha_rows 2^22 x 16", ours(lambda: hal.hash_rows(dg, m)), theirs(lambda: R.sha_rows(dg, m, r, 16))))
    res = [{"op": k, "ours_ms": round(a, 4), "reference_cuda_ms": round(b, 4), "speedup": round(b / a, 2)} for k, a, b in rows]
    for x_ in res:
        print("%-32s ours %9.4f ms   reference CUDA %9.4f ms   x%.2f" % (x_["op"], x_["ours_ms"], x_["reference_cuda_ms"], x_["speedup"]))
    print(json.dumps(res))
    hal.close()


if __name__ == "__main__":
    main()
