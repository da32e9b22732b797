import os, sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
from oracle import embracenet_oracle as O
from tests.golden.cases import ARCH_S, make_inputs
from tests.test_gpu_parity import to_archspec
from embrace_b200 import Engine, _native as N
spec, B = ARCH_S, 128
P = O.init_params(spec, 5)
x, bases, y = make_inputs(spec, B, 6)
def run(fork, steps=12, prec='fp32', tc=False):
    N.set_option('fork', fork)
    eng = Engine(to_archspec(spec), max_batch=B, precision=prec, seed=99, tensor_core=tc)
    eng.load_numpy(P)
    cfg = eng.opt_config('adam', lr=3e-3, weight_decay=1e-4)
    tx, tb, ty = torch.from_numpy(x.astype(np.float32)).cuda(), torch.from_numpy(bases).cuda(), torch.from_numpy(y.astype(np.int32)).cuda()
    for s in range(steps):
        d = O.make_draws(spec, B, 5000 + s)
        eng.train_step(tx, tb, ty, cfg, draws=d)
    torch.cuda.synchronize()
    return eng.params_numpy()
for prec, tc in (('fp32', False), ('bf16', True)):
    a = run(0, prec=prec, tc=tc); b = run(0, prec=prec, tc=tc); c = run(1, prec=prec, tc=tc); d = run(1, prec=prec, tc=tc)
    def dist(p, q): return max(float(np.abs(p[k] - q[k]).max() / max(np.abs(q[k]).max(), 1e-30)) for k in p)
    print(prec, 'nofork vs nofork', dist(a, b), 'fork vs fork', dist(c, d), 'fork vs nofork', dist(c, a), flush=True)
