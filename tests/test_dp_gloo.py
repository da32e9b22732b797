"""Host-side data-parallel logic on CPU with a world_size-2 gloo group (the engine itself needs a GPU)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from embrace_b200.dp import shard_rows, merge_step_metrics


def test_shard_rows_partition_exactly():
    for gb, world in ((8192, 8), (8192, 3), (10, 4), (7, 8)):
        ranges = [shard_rows(gb, r, world) for r in range(world)]
        assert ranges[0][0] == 0 and ranges[-1][1] == gb
        for (a, b), (c, d) in zip(ranges, ranges[1:]):
            assert b == c and b >= a
        sizes = [b - a for a, b in ranges]
        assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        # each rank holds its share of the globally-normalised loss and its own confusion counts
        recs = [dict(loss=0.25 * (rank + 1), tp=rank + 1, fp=2, fn=0, tn=10 * rank), dict(loss=0.5, tp=0, fp=0, fn=1, tn=3)]
        merged = merge_step_metrics(recs)
        # labels -> global positive count, as the bench computes it per step
        y = torch.tensor([1, 0, 0, 1, 0, 0, 0, 1])
        lo, hi = shard_rows(len(y), rank, world)
        local_pos = torch.tensor([int(y[lo:hi].sum())])
        dist.all_reduce(local_pos)
        q.put((rank, merged, int(local_pos)))
    finally:
        dist.destroy_process_group()


def test_metrics_merge_and_global_counts_world2():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, merged, npos in out:
        assert npos == 3
        assert merged[0] == dict(loss=0.75, tp=3, fp=4, fn=0, tn=10)
        assert merged[1] == dict(loss=1.0, tp=0, fp=0, fn=2, tn=6)
