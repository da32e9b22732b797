"""Deterministic-reduction mode (the reference's CPU path is run-to-run reproducible for a fixed
seed; a GPU step that accumulates weight gradients with fp32 atomics is not).

`emb_set_option("deterministic", 1)` gives every fp32 accumulation exactly one contributor in a fixed order: no split-K, one
CTA per weight-gradient tile, one block per bias / head-gradient column group, one CTA for the first-layer weight gradient
(slower; a verification mode).  The BatchNorm sums keep their fp64 atomics: their order changes the result below fp32
resolution.  Conv bias gradients are exactly zero by construction in every mode (engine.cu, cnn_backward_t).
With the mode on, two runs of the same training steps must agree BIT FOR BIT; with it off they drift apart (measured r2: up
to 2e-2 of a tensor's maximum after 12 Adam steps in fp32 -- Adam's per-element normalisation amplifies last-bit noise)."""
import numpy as np
import pytest

from oracle import embracenet_oracle as O
from tests.golden.cases import ARCH_S, make_inputs
from tests.test_gpu_parity import to_archspec

pytestmark = pytest.mark.gpu


def _train(precision, tc, steps, B=96):
    import torch
    from embrace_b200 import Engine
    spec = ARCH_S
    P = O.init_params(spec, 5)
    x, bases, y = make_inputs(spec, B, 6)
    eng = Engine(to_archspec(spec), max_batch=B, precision=precision, seed=99, tensor_core=tc)
    eng.load_numpy(P)
    cfg = eng.opt_config('adam', lr=3e-3, weight_decay=1e-4)
    tx, tb, ty = torch.from_numpy(x.astype(np.float32)).cuda(), torch.from_numpy(bases).cuda(), torch.from_numpy(y.astype(np.int32)).cuda()
    for _ in range(steps):
        eng.train_step(tx, tb, ty, cfg)          # Philox draws: same seed, same step counter -> same masks and selection
    torch.cuda.synchronize()
    return eng.params_numpy()


@pytest.mark.parametrize('precision,tc', [('fp32', False), ('bf16', True)])
def test_deterministic_mode_is_bitwise_reproducible(precision, tc):
    from embrace_b200 import _native as N
    N.set_option('deterministic', 1)
    try:
        a, b = _train(precision, tc, 6), _train(precision, tc, 6)
    finally:
        N.set_option('deterministic', 0)
    for k in a:
        assert np.array_equal(a[k], b[k]), k
    c = _train(precision, tc, 6)                  # the default mode still trains to the same place within its noise
    worst = max(float(np.abs(c[k] - a[k]).max() / max(np.abs(a[k]).max(), 1e-30)) for k in a)
    print(precision, 'default vs deterministic mode, worst tensor max-norm distance after 6 steps:', worst)
    assert np.isfinite(worst)
