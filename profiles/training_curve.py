"""Training curves of the same planted-signal task on three arithmetic paths with the same replayed draws:
  ref   oracle/torch_port.py, fp64 CPU (the reference's arithmetic)
  refp  the same, its initial weights perturbed by 1e-6 relative: the reference's OWN sensitivity (training is chaotic)
  fp32  the engine's verification mode (SIMT kernels, fp32 activations)
  bf16  the engine's production mode (bf16 storage, tcgen05 GEMMs)
Writes gpurun_out/training_curve.json: per-step loss of each, and ranking / hard AUPRC on held-out rows every 10 steps.
Run on a B200:  python profiles/training_curve.py [steps]      (test infrastructure: uses oracle/ as the checker)"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sklearn.metrics import average_precision_score   # noqa: E402
from oracle import embracenet_oracle as O, torch_port as TP   # noqa: E402
from tests.golden.cases import ARCH_S   # noqa: E402
from tests.test_gpu_parity import to_archspec   # noqa: E402
from embrace_b200 import Engine   # noqa: E402


def main(steps=260, decay_at=200, B=1024, N=8192):
    spec = ARCH_S
    rs = np.random.RandomState(3)
    F = spec['F']

    def data(n):
        x = rs.random_sample((n, F)).astype(np.float32).astype(np.float64)
        bases = rs.randint(0, 4, size=(n, 256)).astype(np.uint8)
        motif = np.array([0, 2, 2, 1, 3, 0, 3, 1], dtype=np.uint8)
        z = 8.0 * (x[:, :4].mean(1) - 0.5) * np.sqrt(48.0) - 1.0
        y = (rs.random_sample(n) < 1 / (1 + np.exp(-z))).astype(np.int64)
        for i in np.nonzero(y)[0]:
            if rs.random_sample() < 0.7:
                for _ in range(3):
                    pos = rs.randint(0, 248)
                    bases[i, pos:pos + 8] = motif
        return x, bases, y
    xtr, btr, ytr = data(N)
    xte, bte, yte = data(8192)
    P = O.init_params(spec, 17)
    P = {k: (v.astype(np.float32).astype(np.float64) if v.dtype == np.float64 else v) for k, v in P.items()}
    torch.set_num_threads(os.cpu_count() or 1)
    st = TP.TrainState(spec, {k: v.copy() for k, v in P.items()}, 'adam', lr=3e-3, wd=1e-4)
    prs = np.random.RandomState(1)
    Pp = {k: (v * (1 + 1e-6 * prs.standard_normal(v.shape)) if (v.dtype == np.float64 and v.ndim >= 1) else v.copy()) for k, v in P.items()}
    stp = TP.TrainState(spec, Pp, 'adam', lr=3e-3, wd=1e-4)
    engs = {'fp32': Engine(to_archspec(spec), max_batch=8192, precision='fp32', tensor_core=False),
            'bf16': Engine(to_archspec(spec), max_batch=8192, precision='bf16', tensor_core=True)}
    for e in engs.values():
        e.load_numpy(P)
        e.metrics_reset()
    u = np.random.RandomState(9).random_sample((len(yte), spec['C']))
    out = {'loss': {'ref': [], 'fp32': [], 'bf16': []}, 'eval': []}

    def evaluate(step):
        with torch.no_grad():
            Tt = {k: v.detach() for k, v in st.T.items()}
            lg, _ = TP.forward(spec, Tt, torch.from_numpy(xte), torch.from_numpy(O.onehot_from_bases(bte)), {'embrace_u': u}, training=False)
            lgp, _ = TP.forward(spec, {k: v.detach() for k, v in stp.T.items()}, torch.from_numpy(xte), torch.from_numpy(O.onehot_from_bases(bte)),
                                {'embrace_u': u}, training=False)
        row = {'step': step}
        logits = {'ref': lg.numpy(), 'refp': lgp.numpy()}
        for k, e in engs.items():
            logits[k] = e.forward(torch.from_numpy(xte.astype(np.float32)), torch.from_numpy(bte), training=False,
                                  draws={'embrace_u': u, 'modal_u0': 0.0}).cpu().numpy()
        for k, l in logits.items():
            row[k] = dict(rank=float(average_precision_score(yte, l[:, 1] - l[:, 0])), hard=float(O.auprc_hard(l, yte)))
        out['eval'].append(row)
        print(row, flush=True)

    for s in range(steps):
        lr = 3e-3 if s < decay_at else 3e-4
        for g in list(st.opt.param_groups) + list(stp.opt.param_groups):
            g['lr'] = lr
        lo = (s * B) % N
        xb, bb, yb = xtr[lo:lo + B], btr[lo:lo + B], ytr[lo:lo + B]
        draws = O.make_draws(spec, B, 5000 + s)
        loss, _, _ = st.step(torch.from_numpy(xb), torch.from_numpy(O.onehot_from_bases(bb)), yb, draws)
        out['loss']['ref'].append(float(loss))
        stp.step(torch.from_numpy(xb), torch.from_numpy(O.onehot_from_bases(bb)), yb, draws)
        for k, e in engs.items():
            e.train_step(torch.from_numpy(xb.astype(np.float32)), torch.from_numpy(bb), torch.from_numpy(yb),
                         e.opt_config('adam', lr=lr, weight_decay=1e-4), draws=draws)
        if (s + 1) % 20 == 0:
            for k, e in engs.items():
                m = e.metrics_read()
                out['loss'][k] = [r['loss'] for r in m]
            evaluate(s + 1)
    os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
    with open(os.path.join(ROOT, 'gpurun_out', 'training_curve.json'), 'w') as f:
        json.dump(out, f)


if __name__ == '__main__':
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 260
    main(n, decay_at=int(n * 10 / 13))
