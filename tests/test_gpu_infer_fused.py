"""Inference forwards that never write the pre-pooling conv output (option `infer_fuse`):
  2  transposed conv + in-register pooling (csrc/conv_pool_tc.cuh: TMEM lane = channel, columns = positions; the first, one-hot
     layer through csrc/onehot_pool_tc.cuh: lane = (position block, channel))
  1  BatchNorm(eval) + ReLU + MaxPool1d in the row-major conv GEMM epilogue (EPI_POOL, csrc/gemm_tc.cuh)
  0  unfused (conv writes y, the pooling kernel reads it)
All share arithmetic and rounding points; 1 must equal 0 bit for bit, 2 up to the accumulation order; all are held to the oracle's
eval forward (EmbraceNetMultimodal_NoTrain.py:180-214)."""
import numpy as np
import pytest

from oracle import embracenet_oracle as O
from tests.golden.cases import ARCH_S, ARCH_M, ARCH_L, ARCH_W, espec, make_inputs
from tests.test_gpu_parity import to_archspec

pytestmark = pytest.mark.gpu


# first-layer shapes of the search space the three benchmark archs do not have (CNN_pre.py:22: out_channels_l0 in {16, 32, 64}, k in {5, 11, 15}):
# 16 channels = eight position blocks per sample in onehot_pool_tc.cuh, the last one short; 64 channels with the short tap range
ARCH_C16 = espec(48, [(64, 0.2)], [(16, 15, 0.2), (64, 11, 0.4)], 256, [(64, 0.0)], 0.5)
ARCH_C64 = espec(48, [(64, 0.2)], [(64, 5, 0.2), (32, 5, 0.4)], 256, [], 0.4)
# kernel sizes outside the search space and a 48-channel layer: the kernels' run-time (unspecialised) instantiations
ARCH_K7 = espec(48, [(64, 0.2)], [(32, 7, 0.2), (48, 9, 0.4)], 256, [], 0.4)


@pytest.mark.parametrize('arch,B', [('S', 300), ('M', 130), ('L', 77), ('S', 4096), ('W', 150), ('C16', 300), ('C64', 301), ('K7', 140), ('S', 1)])
def test_fused_pooling_epilogue_equals_unfused_eval_forward(arch, B):
    import torch
    from embrace_b200 import Engine, _native as N
    spec = {'S': ARCH_S, 'M': ARCH_M, 'L': ARCH_L, 'W': ARCH_W, 'C16': ARCH_C16, 'C64': ARCH_C64, 'K7': ARCH_K7}[arch]
    P = O.init_params(spec, 4242)
    # a trained model's BatchNorm: running statistics away from (0, 1), gamma of both signs (a negative gamma reverses the order of
    # the pre-activation values, which a fused pooling that maximises BEFORE the affine map would get wrong)
    rs = np.random.RandomState(4245)
    for k_ in list(P):
        if k_.endswith('running_mean'):
            stem = k_[:-len('running_mean')]
            n = P[k_].shape[0]
            P[k_] = 0.3 * rs.standard_normal(n)
            P[stem + 'running_var'] = 0.5 + rs.random_sample(n)
            P[stem + 'weight'] = (0.5 + rs.random_sample(n)) * np.where(rs.random_sample(n) < 0.3, -1.0, 1.0)
            P[stem + 'bias'] = 0.2 * rs.standard_normal(n)
    P = {k: (v.astype(np.float32).astype(np.float64) if v.dtype == np.float64 else v) for k, v in P.items()}
    x, bases, _ = make_inputs(spec, B, 4243)
    u = np.random.RandomState(4244).random_sample((B, spec['C']))
    av = np.ones((B, 2), dtype=np.float32)
    av[0::5, 0] = 0
    av[1::7, 1] = 0
    av[(av.sum(1) == 0), 0] = 1
    out = {}
    for fuse in (2, 1, 0):
        N.set_option('infer_fuse', fuse)
        try:
            eng = Engine(to_archspec(spec), max_batch=B, precision='bf16', tensor_core=True)
            eng.load_numpy(P)
            l0 = eng.launch_count
            logits, probs = eng.forward(torch.from_numpy(x.astype(np.float32)), torch.from_numpy(bases), training=False,
                                        draws={'embrace_u': u}, availabilities=torch.from_numpy(av), want_probs=True)
            torch.cuda.synchronize()
            out[fuse] = (logits.cpu().numpy(), probs.cpu().numpy(), eng.launch_count - l0)
        finally:
            N.set_option('infer_fuse', 2)
    for fuse in (2, 1):
        assert np.isfinite(out[fuse][0]).all()
        assert out[fuse][2] <= out[0][2], 'a fused forward must not launch more kernels'
    # pooling in the row-major epilogue shares every accumulation with the unfused forward: identical bits
    assert np.array_equal(out[1][0], out[0][0]), np.abs(out[1][0] - out[0][0]).max()
    # the transposed kernel folds the conv bias into the BatchNorm shift and does not round the conv output to bf16 on the way:
    # one rounding fewer than the unfused forward (measured 3e-3 of the logit scale)
    scale = np.abs(out[0][0]).max()
    assert np.abs(out[2][0] - out[0][0]).max() <= 1e-2 * scale, np.abs(out[2][0] - out[0][0]).max() / scale
    if B <= 300:
        ref = O.predict_proba(spec, P, x, bases, u, availabilities=av)
        for fuse in (2, 1, 0):
            assert np.abs(out[fuse][1] - ref).max() <= 5e-3


def test_predict_host_equals_device_forward_on_a_large_batch():
    """With exactly one modality available per row the eval forward is deterministic (the selection takes every dimension from that
    modality, EmbraceNetMultimodal.py:63-76), so the host entry must reproduce the device forward row for row."""
    import torch
    from embrace_b200 import Engine
    spec, B = ARCH_S, 4 * 4096 + 1234
    eng = Engine(to_archspec(spec), max_batch=B, precision='bf16', tensor_core=True)
    eng.load_numpy(O.init_params(spec, 77))
    x, bases, _ = make_inputs(spec, B, 78)
    av = np.zeros((B, 2), dtype=np.float32)
    av[np.arange(B), np.random.RandomState(79).randint(0, 2, size=B)] = 1.0
    xt, bt, at = torch.from_numpy(x.astype(np.float32)), torch.from_numpy(bases), torch.from_numpy(av)
    _, ref = eng.forward(xt, bt, training=False, availabilities=at, want_probs=True)
    got = eng.predict_host(xt.pin_memory(), bt.pin_memory(), at.pin_memory())
    assert np.isfinite(got).all()
    assert np.abs(got - ref.cpu().numpy()).max() <= 1e-6, np.abs(got - ref.cpu().numpy()).max()


def test_predict_host_pipelined_loop_returns_every_batch_in_order():
    """The software-pipelined scoring loop (emb_predict_host_pipelined): batch i's buffer is valid after call i + 1 (or the
    flush); deterministic availabilities make every batch comparable with the synchronous entry."""
    import torch
    from embrace_b200 import Engine
    spec, B, n_batches = ARCH_S, 700, 5
    eng = Engine(to_archspec(spec), max_batch=B, precision='bf16', tensor_core=True)
    eng.load_numpy(O.init_params(spec, 91))
    batches, want = [], []
    for i in range(n_batches):
        nb = B if i != 3 else 333                                     # a shorter batch in the middle of the stream
        x, bases, _ = make_inputs(spec, nb, 100 + i)
        av = np.zeros((nb, 2), dtype=np.float32)
        av[np.arange(nb), np.random.RandomState(200 + i).randint(0, 2, size=nb)] = 1.0
        t = (torch.from_numpy(x.astype(np.float32)).pin_memory(), torch.from_numpy(bases).pin_memory(), torch.from_numpy(av).pin_memory())
        batches.append(t)
        want.append(eng.predict_host(*t))
    outs = [torch.full((t[0].shape[0],), float('nan')).pin_memory() for t in batches]
    flags = [eng.predict_host_pipelined(*t, o) for t, o in zip(batches, outs)]
    assert flags == [False] + [True] * (n_batches - 1)
    for i in range(n_batches - 1):                                    # valid as soon as the following call returned
        assert np.abs(outs[i].numpy() - want[i]).max() <= 1e-6, i
    eng.predict_host_flush()
    assert np.abs(outs[-1].numpy() - want[-1]).max() <= 1e-6
