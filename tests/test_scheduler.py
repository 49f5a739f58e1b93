"""CPU: the host-side segment scheduler (risc0_b200/scheduler.py, the reference's r0vm worker model) with simulated
provers - the way the reference tests its actors (DevModeDelay, r0vm/src/actors/config.rs:62-67): ordering of results,
S segments over G devices with S not a multiple of G, dynamic balancing of uneven segments, bounded look-ahead,
upload-before-prove (depth 2) per device, error propagation."""
import threading
import time

import pytest

from risc0_b200.scheduler import SegmentScheduler


class FakeCluster:
    def __init__(self, prove_s=0.01, fail_on=None):
        self.prove_s, self.fail_on = prove_s, fail_on
        self.lock = threading.Lock()
        self.events = []          # (what, device, index)
        self.in_flight_preflight = 0
        self.max_ahead = 0
        self.proved = 0

    def preflight(self, seg):
        time.sleep(0.001)
        with self.lock:
            self.events.append(("preflight", None, seg["i"]))
        return seg

    def make_device(self, d):
        return {"d": d}

    def upload(self, ctx, pf):
        with self.lock:
            self.events.append(("upload", ctx["d"], pf["i"]))
        return pf

    def prove(self, ctx, handle):
        if self.fail_on == handle["i"]:
            raise RuntimeError("witgen: eqz failure at cycle 7")
        time.sleep(self.prove_s * handle.get("w", 1))
        with self.lock:
            self.events.append(("prove", ctx["d"], handle["i"]))
            self.proved += 1
        return ("seal%d" % handle["i"], None, None)

    def scheduler(self, devices, **kw):
        return SegmentScheduler(devices, self.preflight, self.make_device, self.upload, self.prove, **kw)


@pytest.mark.parametrize("S,G", [(1, 1), (5, 1), (8, 2), (7, 3), (3, 4), (32, 8)])
def test_results_in_order_any_S_over_G(S, G):
    c = FakeCluster()
    res = c.scheduler(list(range(G))).run([{"i": i} for i in range(S)])
    assert [r.index for r in res] == list(range(S))
    assert [r.seal for r in res] == ["seal%d" % i for i in range(S)]
    assert c.proved == S
    used = {r.device for r in res}
    assert used <= set(range(G)) and len(used) == min(S, G) or S < 2 * G


def test_uneven_segments_balance_across_devices():
    # one heavy segment (40x) and 20 light ones on 2 devices: the free device takes the light ones meanwhile
    c = FakeCluster(prove_s=0.01)
    segs = [{"i": 0, "w": 40}] + [{"i": i} for i in range(1, 21)]
    t0 = time.perf_counter()
    res = c.scheduler([0, 1]).run(segs)
    wall = time.perf_counter() - t0
    heavy_dev = res[0].device
    light_on_other = sum(1 for r in res[1:] if r.device != heavy_dev)
    assert light_on_other >= 15
    assert wall < 0.01 * (40 + 20) * 0.9        # clearly better than one device doing everything


def test_upload_runs_one_segment_ahead_of_prove_per_device():
    c = FakeCluster(prove_s=0.05)
    c.scheduler([0]).run([{"i": i} for i in range(5)])
    ev = [w for (w, d, i) in c.events if w in ("upload", "prove")]
    uploads = [k for k, w in enumerate(ev) if w == "upload"]
    proves = [k for k, w in enumerate(ev) if w == "prove"]
    assert len(uploads) == len(proves) == 5
    # depth 2 in steady state: the (n + 1)-th upload is started before the n-th proof has finished. Counted by position,
    # not by segment index - two preflight workers may hand segments over out of order - and from n = 1 on: the very first
    # segment may be proved before the second one has left preflight.
    for n in range(1, 4):
        assert uploads[n + 1] < proves[n]


def test_preflight_look_ahead_is_bounded():
    c = FakeCluster(prove_s=0.03)
    sched = c.scheduler([0], gpu_queue_depth=2, cpu_workers=4)
    sched.run([{"i": i} for i in range(10)])
    # when segment k is proved, preflight may be at most (queue depth + workers in flight + the two held by the device)
    done_pre = 0
    worst = 0
    proved = 0
    for w, d, i in c.events:
        if w == "preflight":
            done_pre += 1
        elif w == "prove":
            proved += 1
        worst = max(worst, done_pre - proved)
    assert worst <= 2 + 4 + 2


def test_prover_error_propagates_and_stops():
    c = FakeCluster(fail_on=3)
    with pytest.raises(RuntimeError) as ei:
        c.scheduler([0, 1]).run([{"i": i} for i in range(50)])
    assert "eqz failure" in str(ei.value)
    assert c.proved < 50
