"""Host-side segment scheduler: the reference's worker model for `ProveSegment` tasks, for one process driving G GPUs.

Reference (risc0/r0vm/src/actors/worker.rs:70-76,185-260,585): a worker routes `ProveSegment` to a CPU queue (preflight:
replay the segment, build PreflightResults) and forwards `ProveSegmentCore` to a GPU queue (witgen + prove), queue
depths 1 / 2 / 2, so preflight of segment s + 1 overlaps the proof of segment s; with several devices it re-execs itself
once per GPU (`CUDA_VISIBLE_DEVICES=i`, actors/mod.rs:449-456) and the factory hands tasks to whichever worker is free.
Segments are independent, so there is no GPU <-> GPU traffic: the unit on the wire is the Segment / PreflightResults,
the result a seal (~281 KB at po2 = 20).

Here: threads in ONE process, one `B200Hal` (= one device ordinal + its own streams and pool) per GPU,
ctypes releases the GIL inside every library call so the devices run concurrently.

  segments --> [cpu queue] --preflight workers--> [gpu queue, bounded] --one worker per device--> results (in order)
                                                          |-> upload segment s+1 on the copy stream while s is proved

The GPU queue is SHARED by all device workers: a device that finishes early takes the next segment, so S segments need
not be a multiple of G and uneven segment sizes balance themselves (the reference gets the same from its factory).
`prove_fn` / `preflight_fn` are injectable so the scheduling itself is testable without a GPU - the reference tests its
actors the same way with simulated provers (`DevModeDelay`, r0vm/src/actors/config.rs:62-67).
"""
import queue
import threading
import time
from concurrent.futures import ThreadPoolExecutor


class SegmentResult:
    def __init__(self, index, device, seal, roots, globals_, t_preflight, t_prove):
        self.index, self.device, self.seal, self.roots, self.globals = index, device, seal, roots, globals_
        self.t_preflight, self.t_prove = t_preflight, t_prove


class _DeviceWorker(threading.Thread):
    """one per GPU: keeps one uploaded segment in flight ahead of the one being proved (depth 2)"""

    def __init__(self, sched, device):
        super().__init__(daemon=True, name="r0b200-gpu%d" % device)
        self.sched, self.device = sched, device
        self.proved = 0
        self.busy_s = 0.0

    def run(self):
        s = self.sched
        try:
            ctx = s.make_device(self.device)
        except BaseException as e:   # noqa: BLE001 - reported to the caller of run()
            s._fail(e)
            return
        try:
            pending = None           # (index, pf, uploaded handle, t_preflight)
            done = False
            while True:
                # prefetch: take the next segment and start its upload BEFORE proving the pending one (depth 2); block
                # on the queue only when there is nothing to prove
                nxt = None
                if not done and not s._failed.is_set():
                    try:
                        item = s._gpu_q.get(block=pending is None)
                    except queue.Empty:
                        item = False
                    if item is None:
                        done = True
                        s._gpu_q.put(None)          # let the other device workers see the end marker too
                    elif item is not False:
                        index, pf, t_pre = item
                        nxt = (index, pf, s.upload_fn(ctx, pf), t_pre)
                if pending is not None:
                    index, pf, handle, t_pre = pending
                    t0 = time.perf_counter()
                    seal, roots, globals_ = s.prove_fn(ctx, handle)
                    dt = time.perf_counter() - t0
                    self.busy_s += dt
                    self.proved += 1
                    s._deliver(SegmentResult(index, self.device, seal, roots, globals_, t_pre, dt))
                pending = nxt
                if pending is None and (done or s._failed.is_set()):
                    break
        except BaseException as e:   # noqa: BLE001
            s._fail(e)
        finally:
            s.close_device(ctx)


class SegmentScheduler:
    def __init__(self, devices, preflight_fn, make_device, upload_fn, prove_fn, close_device=lambda ctx: None,
                 cpu_workers=2, gpu_queue_depth=2):
        """devices: device ordinals. preflight_fn(segment) -> pf. make_device(ordinal) -> ctx (called on the worker's own
        thread). upload_fn(ctx, pf) -> handle (returns at once). prove_fn(ctx, handle) -> (seal, roots, globals)."""
        assert devices, "no devices"
        self.devices = list(devices)
        self.preflight_fn, self.make_device = preflight_fn, make_device
        self.upload_fn, self.prove_fn, self.close_device = upload_fn, prove_fn, close_device
        self.cpu_workers = cpu_workers
        # bounded: preflight runs at most `depth` segments per device ahead of the provers (PreflightResults are big)
        self._gpu_q = queue.Queue(maxsize=max(1, gpu_queue_depth) * len(self.devices))
        self._results = {}
        self._lock = threading.Lock()
        self._failed = threading.Event()
        self._error = None
        self.workers = []

    def _fail(self, e):
        with self._lock:
            if self._error is None:
                self._error = e
        self._failed.set()
        try:
            self._gpu_q.put_nowait(None)
        except queue.Full:
            pass

    def _deliver(self, res):
        with self._lock:
            self._results[res.index] = res

    def run(self, segments):
        """prove every segment; returns the SegmentResults in submission order"""
        segments = list(segments)
        self.workers = [_DeviceWorker(self, d) for d in self.devices]
        for w in self.workers:
            w.start()

        def stage(i):
            if self._failed.is_set():
                return
            t0 = time.perf_counter()
            pf = self.preflight_fn(segments[i])
            item = (i, pf, time.perf_counter() - t0)
            while not self._failed.is_set():
                try:
                    self._gpu_q.put(item, timeout=0.1)
                    return
                except queue.Full:
                    continue

        try:
            with ThreadPoolExecutor(max_workers=self.cpu_workers) as pool:
                # submission order = proving order (up to the interleaving of the preflight workers)
                for f in [pool.submit(stage, i) for i in range(len(segments))]:
                    f.result()
        except BaseException as e:   # noqa: BLE001
            self._fail(e)
        if not self._failed.is_set():
            self._gpu_q.put(None)
        for w in self.workers:
            w.join()
        if self._error is not None:
            raise self._error
        return [self._results[i] for i in range(len(segments))]


def b200_scheduler(devices, hashfn="poseidon2", rand_z=(1, 2, 3, 4), **kw):
    """SegmentScheduler over real devices: preflight = risc0_b200.preflight.PreflightResults, prove = the one-call
    device prove_core (r0b200_segment_upload / r0b200_prove_segment)."""
    from . import preflight as PF
    from .hal import B200Hal, SegmentProver

    def make_device(ordinal):
        hal = B200Hal(ordinal, hashfn)
        return (hal, SegmentProver(hal))

    def upload(ctx, pf):
        return ctx[1].upload_segment(pf)

    def prove(ctx, handle):
        seal, roots, _qpos, glob = ctx[1].prove_segment(handle)
        return seal, roots, glob

    def close(ctx):
        ctx[0].close()

    def preflight(segment):
        return segment if isinstance(segment, PF.PreflightResults) else PF.PreflightResults(segment, rand_z)

    return SegmentScheduler(devices, preflight, make_device, upload, prove, close, **kw)
