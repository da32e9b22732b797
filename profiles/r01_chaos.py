import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import embracenet_oracle as O
from tests.golden.cases import CASES, ARCH_M, make_inputs
def nerr(a,b): return np.abs(a-b).max()/max(np.abs(b).max(),1e-30)
spec, B, seed = ARCH_M, 48, 91
P = O.init_params(spec, seed); x, bases, y = make_inputs(spec, B, seed+1)
draws = O.make_draws(spec, B, seed+100, force_modal=True)
def run(P, quant):
    P = {k: v.copy() for k, v in P.items()}
    if quant:
        with O.quantized(O.bf16_round): return O.train_step(spec, P, x, bases, y, draws)
    return O.train_step(spec, P, x, bases, y, draws)
r0 = run(P, True)
rs = np.random.RandomState(0)
P2 = {k: (v * (1 + 1e-6 * rs.standard_normal(v.shape)) if v.dtype == np.float64 and v.ndim >= 1 else v) for k, v in P.items()}
r1 = run(P2, True)
rp = run(P, False); rp2 = run(P2, False)
print('quantised oracle, weights perturbed by 1e-6: logits', nerr(r1['logits'], r0['logits']))
for k in r0['grads']:
    if 'docking_1' in k or ('CNN' in k and 'weight' in k):
        print(f'{k:32s} quant-perturb {nerr(r1["grads"][k], r0["grads"][k]):.2e}   fp64-perturb {nerr(rp2["grads"][k], rp["grads"][k]):.2e}  quant-vs-fp64 {nerr(r0["grads"][k], rp["grads"][k]):.2e}')
