"""World-size-2 gloo test of the sharded-inference host logic (row shards, final all-gather) with a stand-in engine."""
import os

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


class _FakeEngine:
    max_batch, device = 7, torch.device('cpu')

    def forward(self, x, codes, training, availabilities=None, want_probs=False):
        p = x.sum(1) * (availabilities[:, 0] if availabilities is not None else 1.0)
        return None, p.float()


def _worker(rank, world, port, n, out):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    import embrace_b200  # noqa: F401
    from embrace_b200.infer import score_regions_sharded
    x = torch.arange(n * 3, dtype=torch.float32).reshape(n, 3)
    av = torch.ones(n, 2)
    av[::4, 0] = 0
    lo, hi, s = score_regions_sharded(_FakeEngine(), x, None, av, rank=rank, world=world, batch=5, gather=True)
    out[rank] = (lo, hi, s.numpy().copy())
    dist.destroy_process_group()


def test_sharded_scores_gather_in_row_order():
    n, world = 23, 2
    out = mp.Manager().dict()
    mp.spawn(_worker, args=(world, 29641, n, out), nprocs=world, join=True)
    x = np.arange(n * 3, dtype=np.float32).reshape(n, 3)
    want = x.sum(1)
    want[::4] = 0
    for r in range(world):
        lo, hi, s = out[r]
        assert (lo, hi) == (0, n)
        np.testing.assert_allclose(s, want)
