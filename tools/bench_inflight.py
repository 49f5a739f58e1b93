#!/usr/bin/env python3
"""Throughput of one GPU with W segments in flight (W host threads, each its own context / stream / SegmentProver):
    python tools/bench_inflight.py [--po2 20] [--steps 12] [--workers 1,2,3] [--e2e]
A step is one whole prove_core of the same loop-guest segment. Work is handed out from a shared counter, so the workers
finish together. Prints one JSON line per W: ms per segment (CUDA events bracketing the whole run on worker 0's stream,
both sides synchronised over all contexts), and whether every seal equals the single-worker one."""
import argparse
import json
import os
import sys
import threading

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench  # noqa: E402
from risc0_b200 import B200Hal, SegmentProver  # noqa: E402


def run(pf, W, steps, e2e, ref_seal):
    hals = [B200Hal(0) for _ in range(W)]
    provers = [SegmentProver(h) for h in hals]
    resident = [None if e2e else p.upload_segment(pf) for p in provers]
    for h, p, r in zip(hals, provers, resident):      # warm-up: one proof per context
        if e2e:
            p.prove_segment(p.upload_segment(pf))
        else:
            p.prove_segment(r, free=False)
        h.sync()
    lock, nxt, seals, errs = threading.Lock(), [0], [], []

    def take():
        with lock:
            i = nxt[0]
            nxt[0] += 1
            return i if i < steps else None

    def worker(w):
        try:
            p = provers[w]
            if e2e:
                i = take()
                up = p.upload_segment(pf) if i is not None else None
                while i is not None:
                    j = take()
                    nup = p.upload_segment(pf) if j is not None else None     # next segment travels while this one is proved
                    seals.append(p.prove_segment(up)[0])
                    i, up = j, nup
            else:
                while take() is not None:
                    seals.append(p.prove_segment(resident[w], free=False)[0])
        except BaseException as e:   # noqa: BLE001
            errs.append(e)

    for h in hals:
        h.sync()
    hals[0].timer_start()
    th = [threading.Thread(target=worker, args=(w,)) for w in range(W)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    for h in hals:
        h.sync()
    ms = hals[0].timer_stop()
    if errs:
        raise errs[0]
    same = all(np.array_equal(s, seals[0]) for s in seals) and (ref_seal is None or np.array_equal(seals[0], ref_seal))
    out = {"workers": W, "steps": steps, "e2e": e2e, "ms_per_segment": round(ms / steps, 3), "seals_identical": bool(same),
           "cycles_per_s": round(pf.user_cycles * steps / (ms * 1e-3)), "peak_device_bytes": max(h.bytes_peak() for h in hals)}
    for p, r in zip(provers, resident):
        if r is not None:
            p.free_segment(r)
    for h in hals:
        h.close()
    return out, seals[0]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--po2", type=int, default=20)
    ap.add_argument("--steps", type=int, default=12)
    ap.add_argument("--workers", default="1,2,3")
    ap.add_argument("--e2e", action="store_true")
    a = ap.parse_args()
    pf = bench.build_segment(a.po2)
    if a.e2e:
        import torch
        pf = bench.PinnedSegment(pf, torch)
    ref = None
    for W in [int(x) for x in a.workers.split(",")]:
        out, seal = run(pf, W, a.steps, a.e2e, ref)
        ref = seal if ref is None else ref
        print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
