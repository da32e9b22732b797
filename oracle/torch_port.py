"""PyTorch-CPU port of the reference's train step -- TEST / BASELINE INFRASTRUCTURE ONLY.

The reference's arithmetic is delegated to PyTorch library calls on fp64 CPU tensors
(`model.double()`, training_models_multimodal.py:115).  This file issues the SAME library calls in the
same order (nn.Linear/Conv1d/BatchNorm1d/MaxPool1d functional forms, autograd backward,
nn.CrossEntropyLoss on fp32 logits, torch.optim.Adam), so that timing it on the GPU box's host cores is
timing the reference's own CPU path (`cpu_baseline.kind = "port"`: /root/reference cannot travel to
the GPU box).  It is pinned by tests/test_oracle_golden.py::test_torch_port_* against the golden vectors.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may import it.
Reference lines: EmbraceNetMultimodal.py:34-90,159-193; FFNN_pre.py:10-49; CNN_pre.py:12-76;
training_models_multimodal.py:132-162; utils.py:121-140.
"""
import numpy as np
import torch
import torch.nn.functional as F

from . import embracenet_oracle as O


def params_to_torch(P, dtype=torch.float64, requires_grad=True):
    out = {}
    for k, v in P.items():
        t = torch.from_numpy(np.asarray(v).copy())
        if not O.is_buffer(k):
            t = t.to(dtype).requires_grad_(requires_grad)
        elif not k.endswith('num_batches_tracked'):
            t = t.to(dtype)
        out[k] = t
    return out


# ---- bf16 STORAGE emulation (what EMB_PREC_BF16 does: every stored activation / activation gradient and every GEMM weight
# operand is rounded to bfloat16, arithmetic and accumulation stay wide) -- the same rounding points as
# embracenet_oracle.quantized(bf16_round), expressed as autograd functions so that the fp64 port can play the role of
# "the engine with exact arithmetic" at batch sizes the numpy oracle cannot reach.
def _bf16(x):
    return x.to(torch.float32).to(torch.bfloat16).to(x.dtype)


class _QF(torch.autograd.Function):       # round the value, pass the gradient
    @staticmethod
    def forward(ctx, x):
        return _bf16(x)

    @staticmethod
    def backward(ctx, g):
        return g


class _QB(torch.autograd.Function):       # pass the value, round the gradient
    @staticmethod
    def forward(ctx, x):
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        return _bf16(g)


class _QFB(torch.autograd.Function):      # a tensor the engine stores together with its gradient
    @staticmethod
    def forward(ctx, x):
        return _bf16(x)

    @staticmethod
    def backward(ctx, g):
        return _bf16(g)


class _BNTrainQ(torch.autograd.Function):
    """Training-mode BatchNorm1d whose backward rounds dz ONLY inside the dy formula (the engine's apply pass reads the
    routed gradient as bf16, while its reduction pass sums it unrounded): embracenet_oracle.bn_train_bwd under quantized()."""
    @staticmethod
    def forward(ctx, y, gamma, beta):
        mu = y.mean(dim=(0, 2))
        var = y.var(dim=(0, 2), unbiased=False)
        rstd = 1.0 / torch.sqrt(var + 1e-5)
        xhat = (y - mu[None, :, None]) * rstd[None, :, None]
        ctx.save_for_backward(xhat, rstd, gamma)
        return gamma[None, :, None] * xhat + beta[None, :, None]

    @staticmethod
    def backward(ctx, dz):
        xhat, rstd, gamma = ctx.saved_tensors
        n = dz.shape[0] * dz.shape[2]
        dbeta = dz.sum(dim=(0, 2))
        dgamma = (dz * xhat).sum(dim=(0, 2))
        dy = (gamma * rstd)[None, :, None] * (_bf16(dz) - dbeta[None, :, None] / n - xhat * dgamma[None, :, None] / n)
        return dy, dgamma, dbeta


def _ident(x):
    return x


def _dropout(x, p, training, u):
    if (not training) or p == 0:
        return x
    if u is None:
        return F.dropout(x, p, True)                      # the reference's own RNG path (timing runs)
    keep = (torch.as_tensor(u, dtype=x.dtype) >= p).to(x.dtype)
    return x * keep / (1.0 - p)


def forward(spec, T, x_ffnn, x_onehot, draws=None, training=True, availabilities=None, emulate_bf16=False):
    """T: params_to_torch(...) dict. x_onehot: [B,4,256] fp64 (what the reference feeds Conv1d).
    emulate_bf16: round at the engine's bf16 storage points (see above); False = the reference's arithmetic."""
    draws = draws or {}
    qf, qb, qfb = (_QF.apply, _QB.apply, _QFB.apply) if emulate_bf16 else (_ident, _ident, _ident)
    kind = spec.get('kind', 'embracenet')
    pre_f = 'FFNN.model.' if kind == 'embracenet' else 'model.'
    pre_c = 'CNN.CNN_model.' if kind == 'embracenet' else 'CNN_model.'
    xf = xc = None
    if kind != 'cnn':
        h = qf(x_ffnn)
        for i, p in enumerate(spec['ffnn_dropout']):
            h = F.relu(qb(F.linear(h, qf(T[f'{pre_f}{3*i}.weight']), T[f'{pre_f}{3*i}.bias'])))
            h = qf(_dropout(h, p, training, draws.get('ffnn_drop', [None] * 4)[i] if draws else None))
        xf = h
    if kind != 'ffnn':
        h = x_onehot
        for i, (k, p) in enumerate(zip(spec['cnn_kernels'], spec['cnn_dropout'])):
            w = T[f'{pre_c}{5*i}.weight']
            h = qfb(F.conv1d(h, w if i == 0 else qf(w), T[f'{pre_c}{5*i}.bias'], stride=1, padding=int((k - 1) / 2)))
            if emulate_bf16 and training:
                with torch.no_grad():         # running statistics exactly as F.batch_norm updates them
                    F.batch_norm(h.detach(), T[f'{pre_c}{5*i+1}.running_mean'], T[f'{pre_c}{5*i+1}.running_var'], None, None, True, 0.1, 1e-5)
                h = _BNTrainQ.apply(h, T[f'{pre_c}{5*i+1}.weight'], T[f'{pre_c}{5*i+1}.bias'])
            else:
                h = F.batch_norm(h, T[f'{pre_c}{5*i+1}.running_mean'], T[f'{pre_c}{5*i+1}.running_var'],
                                 T[f'{pre_c}{5*i+1}.weight'], T[f'{pre_c}{5*i+1}.bias'], training, 0.1, 1e-5)
            h = F.max_pool1d(F.relu(h), kernel_size=O.POOL_K, stride=O.POOL_S)
            h = qfb(_dropout(h, p, training, draws.get('cnn_drop', [None] * 4)[i] if draws else None))
        xc = h.reshape(h.size(0), -1)
    if kind == 'ffnn':
        n = len(spec['ffnn_units'])
        return F.linear(xf, qf(T[f'{pre_f}{3*n}.weight']), T[f'{pre_f}{3*n}.bias']), None
    if kind == 'cnn':
        h = qfb(F.linear(xc, qf(T['last_layer1.weight']), T['last_layer1.bias']))
        h = qfb(F.linear(h, qf(T['last_layer2.weight']), T['last_layer2.bias']))
        return F.linear(h, qf(T['last_output.weight']), T['last_output.bias']), None
    B = xf.shape[0]
    if training:
        u0 = draws['modal_u0'] if 'modal_u0' in draws else torch.rand(1)[0].item()
        if u0 >= 0.5:
            rows = torch.as_tensor(draws['modal_rows']) if 'modal_rows' in draws else torch.rand([B])
            availabilities = F.one_hot(torch.round(rows).to(torch.int64), num_classes=2).float()
    sel = torch.tensor([spec['p_ffnn'], 1.0 - spec['p_ffnn']]).repeat(B, 1)
    av = torch.ones(B, 2) if availabilities is None else torch.as_tensor(availabilities).float()
    p = sel.float() * av
    p = p / p.sum(dim=-1, keepdim=True)
    C = spec['C']
    if 'embrace_u' in draws:
        pd = p.double()
        idx = (torch.as_tensor(draws['embrace_u'], dtype=torch.float64) > (pd[:, 0] / (pd[:, 0] + pd[:, 1]))[:, None]).long()
    else:
        idx = torch.multinomial(p, num_samples=C, replacement=True)
    d0 = F.relu(qb(F.linear(xf, qf(T['embracenet.docking_0.weight']), T['embracenet.docking_0.bias'])))
    d1 = F.relu(qb(F.linear(xc, qf(T['embracenet.docking_1.weight']), T['embracenet.docking_1.bias'])))
    stack = torch.stack([d0, d1], dim=-1)
    toggles = F.one_hot(idx, num_classes=2).to(stack.dtype)
    h = qfb((stack * toggles).sum(dim=-1))
    for i, pdrop in enumerate(spec['post_dropout']):
        h = F.relu(qb(F.linear(h, qf(T[f'post.{3*i}.weight']), T[f'post.{3*i}.bias'])))
        h = qf(_dropout(h, pdrop, training, draws.get('post_drop', [None] * 2)[i] if draws else None))
    n = len(spec['post_units'])
    return F.linear(h, qf(T[f'post.{3*n}.weight']), T[f'post.{3*n}.bias']), idx


class TrainState:
    """Parameters + torch.optim optimizer, as the reference's objective()/fit_multimodal() hold them."""

    def __init__(self, spec, P, opt='adam', lr=1e-3, wd=0.0, dtype=torch.float64, emulate_bf16=False):
        self.spec, self.dtype, self.emulate_bf16 = spec, dtype, emulate_bf16
        self.T = params_to_torch(P, dtype)
        plist = [v for k, v in self.T.items() if not O.is_buffer(k)]
        if opt == 'adam':
            self.opt = torch.optim.Adam(plist, lr=lr, weight_decay=wd)
        elif opt == 'rmsprop':
            self.opt = torch.optim.RMSprop(plist, lr=lr, weight_decay=wd)
        else:
            self.opt = torch.optim.NAdam(plist, lr=lr, weight_decay=wd, momentum_decay=4e-3)

    def step(self, x_ffnn, x_onehot, y, draws=None):
        """Loop body of training_models_multimodal.py:132-162. Returns (loss, logits, idx)."""
        y = torch.as_tensor(y).reshape(-1, 1)
        w_pos, w_neg = O.loss_weights_from_labels(y.numpy())
        crit = torch.nn.CrossEntropyLoss(weight=torch.tensor([w_neg, w_pos]))
        self.opt.zero_grad()
        out, idx = forward(self.spec, self.T, x_ffnn, x_onehot, draws, training=True, emulate_bf16=self.emulate_bf16)
        loss = crit.float()(out.float(), y.squeeze(1))
        loss.backward()
        self.opt.step()
        lv = loss.item()
        O.auprc_hard(out.detach().numpy(), y.numpy())      # the reference computes the metric every batch
        return lv, out.detach(), idx


def time_train(spec, B, steps, warmup, seed=0, dtype=torch.float64, threads=None):
    """Wall-clock samples/s of the port on synthetic data (bench.py cpu_baseline / --impl reference)."""
    import time
    if threads:
        torch.set_num_threads(threads)
    from tests.golden.cases import make_inputs
    P = O.init_params(spec, seed)
    x, bases, y = make_inputs(spec, B, seed + 1)
    st = TrainState(spec, P, 'adam', lr=1e-3, wd=1e-4, dtype=dtype)
    x1 = torch.from_numpy(x).to(dtype)
    x2 = torch.from_numpy(O.onehot_from_bases(bases)).to(dtype)
    for _ in range(warmup):
        st.step(x1, x2, y)
    t0 = time.perf_counter()
    for _ in range(steps):
        st.step(x1, x2, y)
    dt = time.perf_counter() - t0
    return dict(samples_per_s=B * steps / dt, seconds=dt, threads=torch.get_num_threads(), B=B, steps=steps)
