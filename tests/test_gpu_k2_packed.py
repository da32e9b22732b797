"""The packed forward pooling kernel (k2_fwd_packed_kernel: 8 channels per thread, max-pooling on the stored bf16 values,
BatchNorm applied to the pooled maximum only) against the streaming kernel it replaces (one thread per channel pair, fp32
BatchNorm on every element).  Same Philox draws, deterministic reductions: logits must be bit-identical (the pooled values
are), and the gradients -- which go through the arg-max codes -- must agree to the last bit as well, except where two
different bf16 values collapse to the same fp32 BatchNorm output (the packed kernel then routes to the larger input, the
streaming kernel to the earlier position; the reference's fp64 arithmetic sides with the packed kernel)."""
import numpy as np
import pytest

from oracle import embracenet_oracle as O
from tests.golden.cases import ARCH_S, ARCH_M, ARCH_L, make_inputs
from tests.test_gpu_parity import to_archspec

pytestmark = pytest.mark.gpu


def _step(spec, B, wide, draws=None):
    import torch
    from embrace_b200 import Engine, _native as N
    N.set_option('k2_wide', wide)
    N.set_option('deterministic', 1)
    try:
        P = O.init_params(spec, 31)
        # make a few BatchNorm scales negative: the packed kernel flips the sign of those channels before pooling
        for k in P:
            if k.endswith('.weight') and P[k].ndim == 1:
                P[k][::3] *= -1.0
        x, bases, y = make_inputs(spec, B, 32)
        eng = Engine(to_archspec(spec), max_batch=B, precision='bf16', seed=77, tensor_core=True)
        eng.load_numpy(P)
        logits = eng.forward(torch.from_numpy(x.astype(np.float32)), torch.from_numpy(bases), training=True, draws=draws)
        eng.backward(eng.loss(logits, torch.from_numpy(y)))
        torch.cuda.synchronize()
        return logits.cpu().numpy(), eng.grads_numpy()
    finally:
        N.set_option('k2_wide', 1)
        N.set_option('deterministic', 0)


@pytest.mark.parametrize('arch,B,replay', [('S', 130, False), ('M', 48, False), ('L', 40, False), ('S', 64, True), ('L', 600, False)])
def test_packed_pooling_kernel_equals_streaming_kernel(arch, B, replay):
    spec = {'S': ARCH_S, 'M': ARCH_M, 'L': ARCH_L}[arch]
    draws = O.make_draws(spec, B, 55, force_modal=False) if replay else None
    l1, g1 = _step(spec, B, 1, draws)
    l0, g0 = _step(spec, B, 0, draws)
    assert np.isfinite(l1).all()
    assert np.array_equal(l1, l0), np.abs(l1 - l0).max()
    for k in g0:
        d = np.abs(g1[k] - g0[k]).max()
        assert d <= 1e-3 * max(np.abs(g0[k]).max(), 1e-12), (k, d)
