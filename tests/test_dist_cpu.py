"""The N>1 host logic on CPU with the gloo backend, world_size 2: sharding, the max-over-ranks
timing reduction and the whole-job throughput aggregation that bench.py uses."""
import os
import socket

import pytest
import torch
import torch.multiprocessing as mp

from voicesplit_b200 import dist as vdist


def test_shard_covers_batch_exactly():
    for gb in (1, 2, 7, 256, 2048):
        for world in (1, 2, 3, 8):
            spans = [vdist.shard(gb, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == gb
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, out):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    d = vdist.init("gloo")
    assert d is not None and d.get_world_size() == world
    a, b = vdist.shard(257, rank, world)
    local_ms = 100.0 + 50.0 * rank            # rank 1 is the slow one
    thr, ms = vdist.aggregate_throughput(b - a, local_ms, d)
    d.barrier()
    if rank == 0:
        out.put((thr, ms))
    d.destroy_process_group()


def test_two_rank_gloo_aggregation():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    thr, ms = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ms == pytest.approx(150.0)                  # max over ranks, not the mean
    assert thr == pytest.approx(257 / 0.150)           # all utterances / slowest rank
