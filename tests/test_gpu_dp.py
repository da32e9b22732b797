"""Data-parallel (2 ranks, NCCL) == single-GPU large batch: SyncBN, global loss weights, partition-invariant Philox
draws, one gradient all-reduce.  Needs 2 GPUs (skipped otherwise): run with `gpurun --gpus 2`."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, q):
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
        spec, GB = ARCH_S, 64
        P = O.init_params(spec, 5)
        x, bases, y = make_inputs(spec, GB, 6)
        eng = Engine(to_archspec(spec), max_batch=GB, precision='fp32', device=dev, seed=4321, tensor_core=False)
        eng.load_numpy(P)
        dp = DataParallel(eng, GB)
        cfg = eng.opt_config('adam', lr=1e-3, weight_decay=1e-3)
        lo, hi = dp.lo, dp.hi
        tx, tb, ty = torch.from_numpy(x[lo:hi].astype(np.float32)), torch.from_numpy(bases[lo:hi]), torch.from_numpy(y[lo:hi])
        eng.metrics_reset()
        dp.train_step(tx, tb, ty, int(y.sum()), cfg)
        torch.cuda.synchronize()
        q.put((rank, lo, hi, eng.grads_numpy(), eng.params_numpy(), eng.metrics_read(), eng.last_selection(hi - lo).cpu().numpy()))
    finally:
        dist.destroy_process_group()


def test_dp2_equals_single_gpu_large_batch():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    import torch.multiprocessing as mp
    from embrace_b200 import Engine
    from oracle import embracenet_oracle as O
    from tests.golden.cases import ARCH_S, make_inputs
    from tests.test_gpu_parity import to_archspec
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=300) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    spec, GB = ARCH_S, 64
    P = O.init_params(spec, 5)
    x, bases, y = make_inputs(spec, GB, 6)
    eng = Engine(to_archspec(spec), max_batch=GB, precision='fp32', seed=4321, tensor_core=False)
    eng.load_numpy(P)
    eng.metrics_reset()
    eng.train_step(torch.from_numpy(x.astype(np.float32)), torch.from_numpy(bases), torch.from_numpy(y), None)
    full_grads = eng.grads_numpy()
    eng.opt_step(eng.opt_config('adam', lr=1e-3, weight_decay=1e-3))
    torch.cuda.synchronize()
    full_params, full_sel, full_m = eng.params_numpy(), eng.last_selection(GB).cpu().numpy(), eng.metrics_read()[0]
    loss = sum(r[5][0]['loss'] for r in res)
    assert abs(loss - full_m['loss']) < 2e-6
    assert sum(r[5][0]['tp'] for r in res) == full_m['tp'] and sum(r[5][0]['tn'] for r in res) == full_m['tn']
    for rank, lo, hi, grads, params, _, sel in res:
        assert np.array_equal(sel, full_sel[lo:hi]), 'selection must not depend on the partition'
        for k, v in full_grads.items():
            wk = k[:-4] + 'weight'
            if k.endswith('.bias') and full_grads[wk].ndim == 3:
                continue                      # conv bias under BatchNorm: analytically zero, rounding noise on both sides
            assert np.abs(grads[k] - v).max() <= 2e-4 * max(np.abs(v).max(), 1e-6), (rank, k)
        for k, v in full_params.items():      # Adam normalises per element: the update can differ by a fraction of lr
            assert np.abs(params[k] - v).max() <= 2e-5 * max(np.abs(v).max(), 1e-6) + 0.1 * 1e-3, (rank, k)
    for k in res[0][3]:
        np.testing.assert_allclose(res[0][3][k], res[1][3][k], rtol=0, atol=0)   # identical all-reduced gradients on both ranks
