import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import embracenet_oracle as O
from oracle import torch_port as TP
from tests.test_gpu_dropin import build
from embrace_b200.BIOINF_tesi.models.utils.training_models_multimodal import lift_optimizer
spec = dict(kind='embracenet', F=16, ffnn_units=[32, 16], ffnn_dropout=[0.2, 0.0], cnn_channels=[16, 32], cnn_kernels=[5, 5],
            cnn_dropout=[0.2, 0.0], C=64, post_units=[32], post_dropout=[0.2], p_ffnn=0.5)
rs = np.random.RandomState(3)
N, B, steps = 2048, 128, 40
def data(n):
    x = rs.random_sample((n, spec['F'])).astype(np.float32).astype(np.float64)
    bases = rs.randint(0, 4, size=(n, 256)).astype(np.uint8)
    motif = np.array([0, 2, 2, 1, 3, 0], dtype=np.uint8)
    y = (rs.random_sample(n) < 1 / (1 + np.exp(-(6 * (x[:, :4].mean(1) - 0.5) - 1.0)))).astype(np.int64)
    for i in np.nonzero(y)[0]:
        if rs.random_sample() < 0.7:
            pos = rs.randint(0, 250); bases[i, pos:pos + 6] = motif
    return x, bases, y
xtr, btr, ytr = data(N)
P = O.init_params(spec, 17)
st = TP.TrainState(spec, {k: v.copy() for k, v in P.items()}, 'adam', lr=3e-3, wd=1e-4)
m = build(spec, P, precision='fp32')
cfg = lift_optimizer(torch.optim.Adam(m.parameters(), lr=3e-3, weight_decay=1e-4))
m.train()
for s in range(steps):
    lo = (s * B) % N
    xb, bb, yb = xtr[lo:lo + B], btr[lo:lo + B], ytr[lo:lo + B]
    draws = O.make_draws(spec, B, 5000 + s)
    st.step(torch.from_numpy(xb), torch.from_numpy(O.onehot_from_bases(bb)), yb, draws)
    m.train_batch(torch.from_numpy(xb), torch.from_numpy(bb), torch.from_numpy(yb), cfg, draws=draws)
    if s < 4 or s % 5 == 4:
        sd = m.state_dict()
        worst = max(((np.abs(sd[k].cpu().numpy() - st.T[k].detach().numpy()).max() / max(np.abs(st.T[k].detach().numpy()).max(), 1e-12)), k) for k in st.T if st.T[k].ndim >= 1)
        print('step', s, 'worst rel param diff', worst)
