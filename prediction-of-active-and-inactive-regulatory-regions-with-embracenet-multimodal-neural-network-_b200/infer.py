"""Sharded genome-wide inference (BASELINE config 5): score N candidate regions with a trained
EmbraceNet, rows split contiguously over the ranks, NO collective on the data path (an optional final gather of the
scores).  Missing modalities are first-class: `availabilities` [N, 2] zeroes a modality per row, and the embracement
selection then takes every dimension from the available one (EmbraceNetMultimodal.py:63-76)."""
import numpy as np
import torch

from .dp import shard_rows


def score_regions(engine, features, codes, availabilities=None, batch=None, out=None):
    """P(class 1) for every row, through the engine's eval-mode forward in chunks of `batch` (device tensors in, device
    tensor out; nothing is copied to the host)."""
    n = len(features) if features is not None else len(codes)
    batch = min(batch or engine.max_batch, engine.max_batch)
    out = out if out is not None else torch.empty(n, dtype=torch.float32, device=engine.device)
    for lo in range(0, n, batch):
        hi = min(n, lo + batch)
        av = availabilities[lo:hi] if availabilities is not None else None
        _, probs = engine.forward(features[lo:hi] if features is not None else None, codes[lo:hi] if codes is not None else None,
                                  training=False, availabilities=av, want_probs=True)
        out[lo:hi] = probs
    return out


def score_regions_sharded(engine, features, codes, availabilities=None, rank=0, world=1, batch=None, gather=False, group=None):
    """Rank r scores rows shard_rows(N, r, world); returns (lo, hi, scores[lo:hi]) or, with gather=True, all N scores on
    every rank (one all-gather of 4 bytes per region)."""
    n = len(features) if features is not None else len(codes)
    lo, hi = shard_rows(n, rank, world)
    local = score_regions(engine, features[lo:hi] if features is not None else None, codes[lo:hi] if codes is not None else None,
                          availabilities[lo:hi] if availabilities is not None else None, batch)
    if not gather or world == 1:
        return lo, hi, local
    import torch.distributed as dist
    sizes = [shard_rows(n, r, world) for r in range(world)]
    width = max(h - l for l, h in sizes)
    pad = torch.zeros(width, dtype=torch.float32, device=local.device)
    pad[:hi - lo] = local
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad, group=group)
    return 0, n, torch.cat([p[:h - l] for p, (l, h) in zip(parts, sizes)])


def synthetic_availabilities(n, seed=0, both=0.8, ffnn_only=0.1):
    """80 % both modalities, 10 % epigenomic only, 10 % sequence only."""
    u = np.random.RandomState(seed).random_sample(n)
    av = np.ones((n, 2), dtype=np.float32)
    av[(u >= both) & (u < both + ffnn_only), 1] = 0.0
    av[u >= both + ffnn_only, 0] = 0.0
    return av
