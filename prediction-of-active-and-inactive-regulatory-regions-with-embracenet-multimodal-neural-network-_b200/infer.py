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


def main(argv=None):
    """Genome-wide scoring benchmark: `python -m torch.distributed.run --nproc-per-node N -m embrace_b200.infer --regions 50000000`
    (or plain `python -m embrace_b200.infer` for one GPU).  Synthetic regions are generated on the device shard by shard;
    availability mix 80 / 10 / 10; prints one JSON line with regions/s summed over the ranks (max-over-ranks device time)."""
    import argparse
    import json
    import os
    import torch.distributed as dist
    from . import Engine, presets
    ap = argparse.ArgumentParser()
    ap.add_argument('--regions', type=int, default=50_000_000)
    ap.add_argument('--arch', default='S')
    ap.add_argument('--batch', type=int, default=65536)
    ap.add_argument('--precision', default='bf16')
    args = ap.parse_args(argv)
    rank, world, local = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1)), int(os.environ.get('LOCAL_RANK', 0))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        os.environ['NCCL_DEBUG'] = os.environ.get('EMB_NCCL_DEBUG', 'WARN')
        dist.init_process_group('nccl', device_id=dev)
    spec = presets.arch(args.arch)
    eng = Engine(spec, max_batch=args.batch, precision=args.precision, device=dev, seed=789)
    eng.init_random(789)
    lo, hi = shard_rows(args.regions, rank, world)
    n = hi - lo
    g = torch.Generator(device=dev).manual_seed(1000 + rank)
    x = torch.rand(n, spec.in_features, device=dev, generator=g)
    codes = torch.randint(0, 4, (n, 256), device=dev, dtype=torch.uint8, generator=g)
    u = torch.rand(n, device=dev, generator=g)
    av = torch.ones(n, 2, device=dev)
    av[(u >= 0.8) & (u < 0.9), 1] = 0.0
    av[u >= 0.9, 0] = 0.0
    out = torch.empty(n, dtype=torch.float32, device=dev)
    score_regions(eng, x[:args.batch * 2], codes[:args.batch * 2], av[:args.batch * 2], out=out[:args.batch * 2])      # warm-up
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = eng.launch_count
    e0.record()
    score_regions(eng, x, codes, av, out=out)
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    chk = out.double().sum().reshape(1)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(chk)
    if rank == 0:
        print(json.dumps({'metric': 'embracenet_infer_regions_per_sec', 'value': args.regions / (float(ms) * 1e-3), 'unit': 'regions/s',
                          'n_gpus': world, 'regions': args.regions, 'ms': float(ms), 'scaling': 'strong (row shards, no collective)',
                          'dtype': 'bf16' if args.precision == 'bf16' else 'f32', 'gpu_launches': int(eng.launch_count - l0),
                          'config': {'workload': f'EmbraceNet arch {args.arch} eval forward, {args.regions} synthetic regions, availability '
                                                 f'80/10/10 (both / epigenomic only / sequence only), batch {args.batch} (BASELINE configs[4])'},
                          'mean_score': float(chk) / args.regions}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
