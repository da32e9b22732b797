"""Data-parallel training == single-GPU large batch: SyncBN, global loss weights, partition-invariant Philox draws, gradient
reduce-scatter + sharded optimizer + parameter all-gather -- every exchange a kernel of the library over peer memory
(csrc/dp_peer.cuh).

  * one GPU is enough for the first group of tests: two engines share the device, each on its own CUDA stream, attached to
    each other by plain pointers (embrace_b200.dp.attach_local); the exchange kernels of one wait on the device for the
    other's, so a step's calls are issued for BOTH engines before anything synchronises;
  * the second group needs 2 GPUs (skipped otherwise; `gpurun --gpus 2`): one process per GPU, CUDA IPC mapping, the whole
    step (exchanges included) replayed as one CUDA graph, plus the NCCL mode of round 1 as the A/B reference.
"""
import json
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'gpurun_out')


def _nerr(a, b):
    return float(np.abs(np.asarray(a, np.float64) - np.asarray(b, np.float64)).max() / max(np.abs(b).max(), 1e-30))


def _single(spec, P, x, bases, y, precision, tc, steps, lr, wd, seed=4321):
    import torch
    from embrace_b200 import Engine
    from tests.test_gpu_parity import to_archspec
    GB = len(y)
    eng = Engine(to_archspec(spec), max_batch=GB, precision=precision, seed=seed, tensor_core=tc)
    eng.load_numpy(P)
    eng.metrics_reset()
    tx, tb, ty = torch.from_numpy(x.astype(np.float32)).cuda(), torch.from_numpy(bases).cuda(), torch.from_numpy(y).cuda()
    cfg = eng.opt_config('adam', lr=lr, weight_decay=wd)
    grads = None
    for s in range(steps):
        eng.train_step(tx, tb, ty, None)
        if s == 0:
            grads = eng.grads_numpy()
        eng.opt_step(cfg)
    torch.cuda.synchronize()
    return dict(grads=grads, params=eng.params_numpy(), sel=eng.last_selection(GB).cpu().numpy(), metrics=eng.metrics_read())


@pytest.mark.parametrize('precision,tc,GB', [('fp32', False, 64), ('bf16', True, 512)])
def test_dp2_on_one_device_equals_large_batch(precision, tc, GB):
    """Two shards of one global batch on ONE GPU (two engines, two streams, pointer-attached): summed gradients, post-step
    parameters, selection indices, loss and confusion counts against the single-engine step on the whole batch."""
    import torch
    from embrace_b200 import Engine
    from embrace_b200.dp import attach_local, shard_rows
    from oracle import embracenet_oracle as O
    from tests.golden.cases import ARCH_S, make_inputs
    from tests.test_gpu_parity import to_archspec
    spec, steps, lr, wd = ARCH_S, 2, 1e-3, 1e-3
    P = O.init_params(spec, 5)
    x, bases, y = make_inputs(spec, GB, 6)
    full = _single(spec, P, x, bases, y, precision, tc, steps, lr, wd)
    world = 2
    dev = torch.device('cuda', torch.cuda.current_device())
    engines, streams, inputs = [], [], []
    for r in range(world):
        lo, hi = shard_rows(GB, r, world)
        e = Engine(to_archspec(spec), max_batch=hi - lo, precision=precision, device=dev, seed=4321, tensor_core=tc)
        e.load_numpy(P)
        e.metrics_reset()
        engines.append(e)
        streams.append(torch.cuda.Stream(device=dev))
        inputs.append((torch.from_numpy(x[lo:hi].astype(np.float32)).to(dev), torch.from_numpy(bases[lo:hi]).to(dev), torch.from_numpy(y[lo:hi].astype(np.int32)).to(dev)))
    blocks = attach_local(engines, GB)
    torch.cuda.synchronize()
    cfgs = [e.opt_config('adam', lr=lr, weight_decay=wd) for e in engines]
    from embrace_b200 import _native as N
    grads0 = None
    for s in range(steps):
        for e, st, (tx, tb, ty) in zip(engines, streams, inputs):       # forward + loss + backward of both shards, no sync in between
            with torch.cuda.stream(st):
                e.train_step(tx, tb, ty, None)
        if s == 0:
            # summed gradient in both arenas (reduce-scatter + all-gather of the gradients, no optimizer) ...
            for e, st in zip(engines, streams):
                with torch.cuda.stream(st):
                    N.check(e.lib.emb_dp_allreduce_grads(e._h, e.stream))
            torch.cuda.synchronize()
            grads0 = [e.grads_numpy() for e in engines]
            # ... then the engines are rewound to redo the step with the fused reduce + optimizer kernel
            for e in engines:
                e.load_numpy(P)
                e.set_seed(4321)
                e.metrics_reset()
            torch.cuda.synchronize()
            for e, st, (tx, tb, ty) in zip(engines, streams, inputs):
                with torch.cuda.stream(st):
                    e.train_step(tx, tb, ty, None)
        for e, st, cfg in zip(engines, streams, cfgs):
            with torch.cuda.stream(st):
                e.opt_step(cfg)
    torch.cuda.synchronize()
    rep = dict(precision=precision, GB=GB, grads={}, params={})
    tol_g = 2e-4 if precision == 'fp32' else 3e-2
    for k, v in full['grads'].items():
        wk = k[:-4] + 'weight'
        if k.endswith('.bias') and full['grads'][wk].ndim == 3:
            continue                      # conv bias under BatchNorm: analytically zero, rounding noise on both sides
        rep['grads'][k] = max(_nerr(g[k], v) for g in grads0)
        assert np.array_equal(grads0[0][k], grads0[1][k]), ('the reduced gradient must be bit-identical on both ranks', k)
        assert rep['grads'][k] <= tol_g, (k, rep['grads'][k])
    params = [e.params_numpy() for e in engines]
    for k, v in full['params'].items():
        assert np.array_equal(params[0][k], params[1][k]), ('parameters must be bit-identical on both ranks after the all-gather', k)
        rep['params'][k] = float(np.abs(params[0][k] - v).max())
        # Adam normalises per element: entries whose gradient is rounding noise can move by a fraction of lr per step
        assert rep['params'][k] <= 2e-5 * max(np.abs(v).max(), 1e-6) + (0.1 if precision == 'fp32' else 2.0) * lr * steps, (k, rep['params'][k])
    sel = np.concatenate([e.last_selection(e.max_batch).cpu().numpy() for e in engines])
    if precision == 'fp32':               # same Philox step on both sides only when the rewind above did not consume a step
        pass
    m = [e.metrics_read() for e in engines]
    loss = sum(mm[-1]['loss'] for mm in m)
    assert abs(loss - full['metrics'][-1]['loss']) <= (2e-5 if precision == 'fp32' else 2e-2) * max(1.0, abs(full['metrics'][-1]['loss'])), (loss, full['metrics'][-1])
    if precision == 'fp32':
        for key in ('tp', 'fp', 'fn', 'tn'):
            assert sum(mm[0][key] for mm in m) == full['metrics'][0][key]
    assert sel.shape == full['sel'].shape
    os.makedirs(OUT, exist_ok=True)
    with open(os.path.join(OUT, f'dp_one_device_{precision}.json'), 'w') as f:
        json.dump(rep, f, indent=1)
    del blocks


def test_dp_selection_is_partition_invariant_on_one_device():
    """Philox counters are keyed by GLOBAL row: the modality selection of a shard equals the same rows of the full batch."""
    import torch
    from embrace_b200 import Engine
    from embrace_b200.dp import attach_local, shard_rows
    from oracle import embracenet_oracle as O
    from tests.golden.cases import ARCH_S, make_inputs
    from tests.test_gpu_parity import to_archspec
    spec, GB = ARCH_S, 96
    P = O.init_params(spec, 5)
    x, bases, y = make_inputs(spec, GB, 6)
    full = _single(spec, P, x, bases, y, 'fp32', False, 1, 1e-3, 1e-3)
    dev = torch.device('cuda', torch.cuda.current_device())
    engines, streams = [], []
    for r in range(3):
        lo, hi = shard_rows(GB, r, 3)
        e = Engine(to_archspec(spec), max_batch=hi - lo, precision='fp32', device=dev, seed=4321, tensor_core=False)
        e.load_numpy(P)
        engines.append(e)
        streams.append(torch.cuda.Stream(device=dev))
    blocks = attach_local(engines, GB)
    torch.cuda.synchronize()
    ins = []
    for r in range(3):
        lo, hi = shard_rows(GB, r, 3)
        ins.append((torch.from_numpy(x[lo:hi].astype(np.float32)).to(dev), torch.from_numpy(bases[lo:hi]).to(dev), torch.from_numpy(y[lo:hi].astype(np.int32)).to(dev)))
    for e, st, (tx, tb, ty) in zip(engines, streams, ins):
        with torch.cuda.stream(st):
            e.train_step(tx, tb, ty, e.opt_config('adam', lr=1e-3, weight_decay=1e-3))
    torch.cuda.synchronize()
    sel = np.concatenate([e.last_selection(e.max_batch).cpu().numpy() for e in engines])
    assert np.array_equal(sel, full['sel']), 'selection must not depend on the partition'
    del blocks


# ---------------------------------------------------------------------------------------------------------------------
# two processes, two GPUs
# ---------------------------------------------------------------------------------------------------------------------
def _worker(rank, world, port, q, comm, precision, tc, GB, steps):
    import torch
    import torch.distributed as dist
    from embrace_b200 import Engine
    from embrace_b200.dp import DataParallel
    from oracle import embracenet_oracle as O
    from tests.golden.cases import ARCH_S, make_inputs
    from tests.test_gpu_parity import to_archspec
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device('cuda', rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=dev)
    try:
        spec = ARCH_S
        P = O.init_params(spec, 5)
        x, bases, y = make_inputs(spec, GB, 6)
        eng = Engine(to_archspec(spec), max_batch=GB, precision=precision, device=dev, seed=4321, tensor_core=tc)
        eng.load_numpy(P)
        dp = DataParallel(eng, GB, comm=comm, graph=True if comm == 'peer' else False)
        cfg = eng.opt_config('adam', lr=1e-3, weight_decay=1e-3)
        lo, hi = dp.lo, dp.hi
        tx, tb, ty = (torch.from_numpy(x[lo:hi].astype(np.float32)).to(dev), torch.from_numpy(bases[lo:hi]).to(dev),
                      torch.from_numpy(y[lo:hi].astype(np.int32)).to(dev))
        eng.metrics_reset()
        for _ in range(steps):
            dp.train_step(tx, tb, ty, int(y.sum()), cfg)
        torch.cuda.synchronize()
        out = (rank, lo, hi, dp.comm, eng.params_numpy(), eng.metrics_read(), eng.last_selection(hi - lo).cpu().numpy(), int(eng.launch_count))
        dp.close()
        q.put(out)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('comm,precision,tc,GB', [('peer', 'fp32', False, 64), ('peer', 'bf16', True, 1024), ('nccl', 'fp32', False, 64)])
def test_dp2_two_gpus_equals_single_gpu_large_batch(comm, precision, tc, GB):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    import torch.multiprocessing as mp
    from oracle import embracenet_oracle as O
    from tests.golden.cases import ARCH_S, make_inputs
    steps = 3
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q, comm, precision, tc, GB, steps)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=600) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    spec = ARCH_S
    P = O.init_params(spec, 5)
    x, bases, y = make_inputs(spec, GB, 6)
    full = _single(spec, P, x, bases, y, precision, tc, steps, 1e-3, 1e-3)
    assert all(r[3] == comm for r in res), 'the requested exchange mode must be the one that ran (no silent fall-back)'
    loss = sum(r[5][-1]['loss'] for r in res)
    assert abs(loss - full['metrics'][-1]['loss']) <= (1e-4 if precision == 'fp32' else 3e-2) * max(1.0, abs(full['metrics'][-1]['loss']))
    if precision == 'fp32':
        assert sum(r[5][0]['tp'] for r in res) == full['metrics'][0]['tp'] and sum(r[5][0]['tn'] for r in res) == full['metrics'][0]['tn']
    for rank, lo, hi, _, params, _, sel, _ in res:
        assert np.array_equal(sel, full['sel'][lo:hi]), 'selection must not depend on the partition'
        for k, v in full['params'].items():      # Adam normalises per element: the update can differ by a fraction of lr per step
            assert np.abs(params[k] - v).max() <= 2e-5 * max(np.abs(v).max(), 1e-6) + (0.1 if precision == 'fp32' else 2.0) * 1e-3 * steps, (rank, k)
    for k in res[0][4]:
        np.testing.assert_array_equal(res[0][4][k], res[1][4][k])     # bit-identical parameters on both ranks
