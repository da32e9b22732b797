import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import embracenet_oracle as O
from tests.golden.cases import CASES, ARCH_M, make_inputs
from tests.test_gpu_parity import to_archspec, nerr
from embrace_b200 import Engine

def run(spec, B, seed, force):
    P = O.init_params(spec, seed)
    x, bases, y = make_inputs(spec, B, seed + 1)
    draws = O.make_draws(spec, B, seed + 100, force_modal=force)
    with O.quantized(O.bf16_round):
        ref = O.train_step(spec, {k: v.copy() for k, v in P.items()}, x, bases, y, draws)
    out = {}
    for tc in (False, True):
        eng = Engine(to_archspec(spec), max_batch=B, precision='bf16', tensor_core=tc)
        eng.load_numpy(P)
        lg = eng.forward(torch.from_numpy(x.astype(np.float32)), torch.from_numpy(bases), training=True, draws=draws)
        dl = eng.loss(lg, torch.from_numpy(y))
        eng.backward(dl)
        out[tc] = (lg.cpu().numpy(), eng.grads_numpy())
    print('logits simt/ref %.2e tc/ref %.2e tc/simt %.2e' % (nerr(out[False][0], ref['logits']), nerr(out[True][0], ref['logits']), nerr(out[True][0], out[False][0])))
    for k, g in ref['grads'].items():
        if np.abs(g).max() < 1e-12: continue
        print('%-32s simt/ref %.2e  tc/ref %.2e  tc/simt %.2e' % (k, nerr(out[False][1][k], g), nerr(out[True][1][k], g), nerr(out[True][1][k], out[False][1][k])))

print('== archM B=48'); run(ARCH_M, 48, 91, True)
print('== deep4 B=37'); run(CASES['deep4']['spec'], 37, 31, None)
