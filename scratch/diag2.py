import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import embracenet_oracle as O
from tests.golden.cases import CASES, ARCH_M, make_inputs
from tests.test_gpu_parity import to_archspec, nerr
from embrace_b200 import Engine

def run(spec, B, seed, force, precision):
    P = O.init_params(spec, seed)
    x, bases, y = make_inputs(spec, B, seed + 1)
    draws = O.make_draws(spec, B, seed + 100, force_modal=force)
    if precision == 'bf16':
        with O.quantized(O.bf16_round):
            ref = O.train_step(spec, {k: v.copy() for k, v in P.items()}, x, bases, y, draws)
    else:
        ref = O.train_step(spec, {k: v.copy() for k, v in P.items()}, x, bases, y, draws)
    eng = Engine(to_archspec(spec), max_batch=B, precision=precision, tensor_core=False)
    eng.load_numpy(P)
    lg = eng.forward(torch.from_numpy(x.astype(np.float32)), torch.from_numpy(bases), training=True, draws=draws)
    dl = eng.loss(lg, torch.from_numpy(y))
    eng.backward(dl)
    g = eng.grads_numpy()
    print(precision, 'force', force, 'logits %.2e' % nerr(lg.cpu().numpy(), ref['logits']), ' '.join('%s=%.1e' % (k.split('.')[-2][-2:] + k[-1], nerr(g[k], v)) for k, v in ref['grads'].items() if np.abs(v).max() > 1e-12 and ('docking_1' in k or 'CNN' in k and 'weight' in k)))

for force in (True, False):
    for prec in ('fp32', 'bf16'):
        run(ARCH_M, 48, 91, force, prec)
spec = dict(ARCH_M); spec['cnn_dropout'] = [0.0, 0.0, 0.0, 0.0]
print('no cnn dropout'); run(spec, 48, 91, True, 'bf16')
spec = dict(ARCH_M); spec['cnn_dropout'] = [0.2, 0.4, 0.4, 0.5]
print('all cnn dropout'); run(spec, 48, 91, True, 'bf16')
