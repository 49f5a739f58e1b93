"""N > 1 host logic on CPU: world_size-2 gloo processes exercising the segment sharding / timing reduction that
bench.py --gpus N uses (no GPU, no kernels)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from risc0_b200 import shard


def test_assign_segments_round_robin():
    assert shard.assign_segments(8, 4, 1) == [1, 5]
    assert shard.assign_segments(3, 8, 5) == []
    allseg = sorted(s for r in range(3) for s in shard.assign_segments(10, 3, r))
    assert allseg == list(range(10))
    with pytest.raises(ValueError):
        shard.assign_segments(4, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        mine = shard.assign_segments(5, world, rank)
        seconds = 1.0 + rank  # rank 1 is the slow one
        tmax = shard.max_over_ranks(seconds)
        counts = shard.gather_counts(len(mine))
        dist.barrier()
        q.put((rank, mine, tmax, counts, shard.whole_job_throughput(counts, tmax)))
    finally:
        dist.destroy_process_group()


def test_two_ranks_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][1] == [0, 2, 4] and res[1][1] == [1, 3]
    for r in res:
        assert r[2] == 2.0            # max over ranks
        assert r[3] == [3, 2]         # every rank sees every rank's count
        assert r[4] == 5 / 2.0        # whole-job throughput
