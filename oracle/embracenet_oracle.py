"""CPU oracle for the EmbraceNet hot path -- TEST INFRASTRUCTURE ONLY.

This file is a plain numpy (fp64) restatement of what the reference's PyTorch
calls compute on the path named by BASELINE.json `north_star`.  It is the
checker for the CUDA engine; nothing in the product package imports it.  Only
`tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / reference
leg may import it.

Parity status: PINNED against the reference itself.  The reference ships no
tests or golden vectors, so the pins are outputs of the
unmodified reference modules imported from /root/reference in the build
container by `tests/golden/make_golden.py` (committed) and stored as
`tests/golden/*.npz`; `tests/test_oracle_golden.py` checks every function here
against them.

Reference lines restated (all under /root/reference/BIOINF_tesi/models):
  FFNN_pre.py:10-49                     -> ffnn_forward / ffnn_backward
  CNN_pre.py:12-76                      -> cnn_forward / cnn_backward
  EmbraceNetMultimodal.py:34-90         -> embrace_probabilities, embrace_select,
                                           docking + embrace in forward()/backward()
  EmbraceNetMultimodal.py:159-193       -> forward() (modality dropout, post head)
  utils/utils.py:80-153                 -> auprc_hard, f1_precision_recall,
                                           loss_weights_from_labels, size_out_convolution
  utils/training_models_multimodal.py:132-162 -> weighted_ce, train_step
The arithmetic itself lives in PyTorch (third-party, un-vendored, version
unpinned upstream; torch 2.11 in this image): Conv1d, BatchNorm1d, MaxPool1d,
Dropout, Linear, multinomial, CrossEntropyLoss, optim.Adam/RMSprop/NAdam.
Their published definitions are what is written out below.

Random draws are always explicit (`draws` dict), never taken from a generator:
  draws['ffnn_drop'][i]  uniforms [B, units_i]          keep = (u >= p)
  draws['cnn_drop'][i]   uniforms [B, Cout_i, Lpool_i]  keep = (u >= p)
  draws['modal_u0']      scalar  (modality-dropout coin, fp32)
  draws['modal_rows']    [B]     (per-row modality coin, fp32; round-half-even)
  draws['embrace_u']     [B, C]  fp64 uniforms, idx = (u > cum0)
  draws['post_drop'][i]  uniforms [B, units_i]
"""
import numpy as np

POOL_K = 10      # CNN_pre.py:18
POOL_S = 2       # CNN_pre.py:20
SEQ_LEN = 256    # CNN_pre.py:21
N_BASES = 4      # CNN_pre.py:22
BN_EPS = 1e-5
BN_MOMENTUM = 0.1


# --------------------------------------------------------------------------
# optional storage-precision emulation.  The reference is fp64 throughout; the CUDA engine's bf16
# precision rounds every STORED activation / gradient and every GEMM weight operand to bf16 (fp32
# accumulation).  `with quantized(bf16_round):` makes the oracle round at exactly those points, so the
# bf16 kernels can be checked tightly; without it the oracle is the plain fp64 restatement.
# --------------------------------------------------------------------------
import contextlib

_Q = None


def bf16_round(x):
    """Round-to-nearest-even to bfloat16, returned as float64."""
    x32 = np.ascontiguousarray(np.asarray(x, dtype=np.float32))
    b = x32.view(np.uint32)
    r = ((b >> np.uint32(16)) & np.uint32(1)) + np.uint32(0x7FFF)
    return ((b + r) & np.uint32(0xFFFF0000)).view(np.float32).astype(np.float64)


@contextlib.contextmanager
def quantized(fn):
    global _Q
    old, _Q = _Q, fn
    try:
        yield
    finally:
        _Q = old


def _q(x):
    return x if _Q is None else _Q(x)


# --------------------------------------------------------------------------
# shape helpers (utils.py:143-153, :178-202)
# --------------------------------------------------------------------------
def size_out_convolution(input_size, kernel, padding, stride):
    return int(((input_size + 2 * padding - kernel) / stride) + 1)


def cnn_lengths(kernels):
    """[(L_conv, L_pool)] per layer: 256 -> 124 -> 58 -> 25 -> 8 for odd k."""
    out, L = [], SEQ_LEN
    for k in kernels:
        p = int((k - 1) / 2)
        Lc = size_out_convolution(L, k, p, 1)
        Lp = size_out_convolution(Lc, POOL_K, 0, POOL_S)
        out.append((Lc, Lp))
        L = Lp
    return out


def param_shapes(spec):
    """Ordered {state_dict key: shape} exactly as the reference nn.Sequential
    builds them (checkpoint row)."""
    kind = spec.get('kind', 'embracenet')
    pre_f = 'FFNN.model.' if kind in ('embracenet', 'concatnet') else 'model.'
    pre_c = 'CNN.CNN_model.' if kind in ('embracenet', 'concatnet') else 'CNN_model.'
    shp = {}
    fin = spec.get('F', 0)
    if kind in ('embracenet', 'concatnet', 'ffnn'):
        for i, u in enumerate(spec['ffnn_units']):
            shp[f'{pre_f}{3*i}.weight'] = (u, fin)
            shp[f'{pre_f}{3*i}.bias'] = (u,)
            fin = u
        if kind == 'ffnn':
            n = len(spec['ffnn_units'])
            shp[f'{pre_f}{3*n}.weight'] = (2, fin)
            shp[f'{pre_f}{3*n}.bias'] = (2,)
    if kind in ('embracenet', 'concatnet', 'cnn'):
        cin = N_BASES
        for i, (co, k) in enumerate(zip(spec['cnn_channels'], spec['cnn_kernels'])):
            shp[f'{pre_c}{5*i}.weight'] = (co, cin, k)
            shp[f'{pre_c}{5*i}.bias'] = (co,)
            shp[f'{pre_c}{5*i+1}.weight'] = (co,)
            shp[f'{pre_c}{5*i+1}.bias'] = (co,)
            shp[f'{pre_c}{5*i+1}.running_mean'] = (co,)
            shp[f'{pre_c}{5*i+1}.running_var'] = (co,)
            shp[f'{pre_c}{5*i+1}.num_batches_tracked'] = ()
            cin = co
        cnn_out = spec['cnn_channels'][-1] * cnn_lengths(spec['cnn_kernels'])[-1][1]
        if kind == 'cnn':     # CNN_net.py:71-73
            shp['last_layer1.weight'] = (1000, cnn_out)
            shp['last_layer1.bias'] = (1000,)
            shp['last_layer2.weight'] = (64, 1000)
            shp['last_layer2.bias'] = (64,)
            shp['last_output.weight'] = (2, 64)
            shp['last_output.bias'] = (2,)
    if kind == 'concatnet':      # ConcatNetMultimodal.py:38-62: post over cat(FFNN out, CNN out)
        fin = spec['ffnn_units'][-1] + cnn_out
        for i, u in enumerate(spec['post_units']):
            shp[f'post.{3*i}.weight'] = (u, fin)
            shp[f'post.{3*i}.bias'] = (u,)
            fin = u
        n = len(spec['post_units'])
        shp[f'post.{3*n}.weight'] = (2, fin)
        shp[f'post.{3*n}.bias'] = (2,)
    if kind == 'embracenet':
        C = spec['C']
        shp['embracenet.docking_0.weight'] = (C, spec['ffnn_units'][-1])
        shp['embracenet.docking_0.bias'] = (C,)
        shp['embracenet.docking_1.weight'] = (C, cnn_out)
        shp['embracenet.docking_1.bias'] = (C,)
        fin = C
        for i, u in enumerate(spec['post_units']):
            shp[f'post.{3*i}.weight'] = (u, fin)
            shp[f'post.{3*i}.bias'] = (u,)
            fin = u
        n = len(spec['post_units'])
        shp[f'post.{3*n}.weight'] = (2, fin)
        shp[f'post.{3*n}.bias'] = (2,)
    return shp


def is_buffer(key):
    return key.endswith(('running_mean', 'running_var', 'num_batches_tracked'))


def init_params(spec, seed):
    """Deterministic synthetic weights (numpy legacy RandomState: stable across
    numpy versions).  Scale follows nn.Linear/Conv1d's U(-1/sqrt(fan_in), ..)."""
    rs = np.random.RandomState(seed)
    shapes = param_shapes(spec)
    P = {}
    for key, shape in shapes.items():
        module = key.rsplit('.', 1)[0]
        if key.endswith('num_batches_tracked'):
            P[key] = np.array(0, dtype=np.int64)
        elif key.endswith('running_mean'):
            P[key] = rs.uniform(-0.1, 0.1, shape)
        elif key.endswith('running_var'):
            P[key] = rs.uniform(0.5, 1.5, shape)
        elif module + '.running_mean' in shapes:
            # BatchNorm affine: gamma near 1, beta near 0
            P[key] = rs.uniform(0.5, 1.5, shape) if key.endswith('weight') else rs.uniform(-0.2, 0.2, shape)
        else:
            wshape = shapes[module + '.weight']
            fan_in = int(np.prod(wshape[1:]))
            bound = 1.0 / np.sqrt(max(fan_in, 1))
            P[key] = rs.uniform(-bound, bound, shape)
    return P


# --------------------------------------------------------------------------
# primitive ops
# --------------------------------------------------------------------------
def linear_fwd(x, W, b):
    return x @ W.T + b


def onehot_from_bases(bases):
    """uint8 [B,256] codes 0..3 (a,c,g,t = sklearn's sorted categories,
    data_pipe/utils.py:269-276) -> fp64 [B,4,256]."""
    B, L = bases.shape
    x = np.zeros((B, N_BASES, L))
    x[np.arange(B)[:, None], bases.astype(np.int64), np.arange(L)[None, :]] = 1.0
    return x


def conv1d_fwd(x, W, b):
    """nn.Conv1d(stride 1, padding (k-1)/2) (CNN_pre.py:37-39).
    y[b,o,l] = bias[o] + sum_{c,k} W[o,c,k] x[b,c,l+k-p]."""
    k = W.shape[2]
    p = int((k - 1) / 2)
    xp = np.pad(x, ((0, 0), (0, 0), (p, p)))
    win = np.lib.stride_tricks.sliding_window_view(xp, k, axis=2)   # [B,C,L,k]
    return np.einsum('bclk,ock->bol', win, W, optimize=True) + b[None, :, None]


def conv1d_bwd(x, W, g):
    """returns (dx, dW, db) for conv1d_fwd."""
    k = W.shape[2]
    p = int((k - 1) / 2)
    xp = np.pad(x, ((0, 0), (0, 0), (p, p)))
    win = np.lib.stride_tricks.sliding_window_view(xp, k, axis=2)   # [B,C,L,k]
    dW = np.einsum('bol,bclk->ock', g, win, optimize=True)
    db = g.sum(axis=(0, 2))
    gp = np.pad(g, ((0, 0), (0, 0), (p, p)))
    gwin = np.lib.stride_tricks.sliding_window_view(gp, k, axis=2)  # [B,O,L,k] : g[l'+j-p]
    # dx[b,c,l'] = sum_{o,k} g[b,o,l'-k+p] W[o,c,k]  ; with j = 2p-k: l'-k+p = l'+j-p
    dx = np.einsum('bolj,ocj->bcl', gwin, W[:, :, ::-1], optimize=True)
    return dx, dW, db


def onehot_conv_fwd(bases, W, b):
    """Layer-0 conv over one-hot input as a table gather-sum:
    out[b,o,l] = bias[o] + sum_{k: 0<=l+k-p<256} W[o, base[b,l+k-p], k]."""
    B, L = bases.shape
    O, _, k = W.shape
    p = int((k - 1) / 2)
    out = np.tile(b[None, :, None], (B, 1, L)).astype(np.float64)
    T = np.transpose(W, (2, 1, 0))                  # [k, 4, O]
    for kk in range(k):
        lo, hi = max(0, p - kk), min(L, L + p - kk)  # output positions with valid source
        src = bases[:, lo + kk - p: hi + kk - p].astype(np.int64)   # [B, n]
        out[:, :, lo:hi] += np.transpose(T[kk][src], (0, 2, 1))
    return out


def onehot_conv_bwd(bases, g, k):
    """dW[o,c,k] = sum_{b,l} g[b,o,l] [base[b,l+k-p]==c] (histogram-add); db = sum g."""
    B, L = bases.shape
    O = g.shape[1]
    p = int((k - 1) / 2)
    dW = np.zeros((O, N_BASES, k))
    for kk in range(k):
        lo, hi = max(0, p - kk), min(L, L + p - kk)
        src = bases[:, lo + kk - p: hi + kk - p]
        gs = g[:, :, lo:hi]
        for c in range(N_BASES):
            m = (src == c)[:, None, :]
            dW[:, c, kk] = (gs * m).sum(axis=(0, 2))
    return dW, g.sum(axis=(0, 2))


def bn_train_fwd(y, gamma, beta, rm, rv):
    """nn.BatchNorm1d in training mode over (B,L) (CNN_pre.py:41)."""
    n = y.shape[0] * y.shape[2]
    mu = y.mean(axis=(0, 2))
    var = ((y - mu[None, :, None]) ** 2).mean(axis=(0, 2))      # biased
    rstd = 1.0 / np.sqrt(var + BN_EPS)
    xhat = (y - mu[None, :, None]) * rstd[None, :, None]
    z = gamma[None, :, None] * xhat + beta[None, :, None]
    new_rm = (1 - BN_MOMENTUM) * rm + BN_MOMENTUM * mu
    new_rv = (1 - BN_MOMENTUM) * rv + BN_MOMENTUM * var * n / max(n - 1, 1)
    return z, xhat, rstd, new_rm, new_rv


def bn_eval_fwd(y, gamma, beta, rm, rv):
    return gamma[None, :, None] * (y - rm[None, :, None]) / np.sqrt(rv[None, :, None] + BN_EPS) \
        + beta[None, :, None]


def bn_train_bwd(dz, xhat, rstd, gamma):
    n = dz.shape[0] * dz.shape[2]
    dbeta = dz.sum(axis=(0, 2))
    dgamma = (dz * xhat).sum(axis=(0, 2))
    dy = (gamma * rstd)[None, :, None] * (_q(dz) - dbeta[None, :, None] / n - xhat * dgamma[None, :, None] / n)
    return dy, dgamma, dbeta


def relu_maxpool_fwd(z):
    """ReLU then MaxPool1d(10, stride 2) (CNN_pre.py:42-44). Returns pooled and
    the first arg-max position (absolute index along L) of each window."""
    r = np.maximum(z, 0.0)
    win = np.lib.stride_tricks.sliding_window_view(r, POOL_K, axis=2)[:, :, ::POOL_S, :]  # [B,C,Lp,10]
    am = win.argmax(axis=3)                            # first max
    pooled = np.take_along_axis(win, am[..., None], axis=3)[..., 0]
    pos = am + (np.arange(win.shape[2]) * POOL_S)[None, None, :]
    return pooled, pos


def relu_maxpool_bwd(z, pos, gpool):
    dz = np.zeros_like(z)
    B, C, Lp = gpool.shape
    bi = np.arange(B)[:, None, None].repeat(C, 1).repeat(Lp, 2)
    ci = np.arange(C)[None, :, None].repeat(B, 0).repeat(Lp, 2)
    np.add.at(dz, (bi, ci, pos), gpool)
    return dz * (z > 0)


def dropout_fwd(x, u, p, training):
    """nn.Dropout with explicit uniforms: keep = (u >= p), scale 1/(1-p).
    p == 0 or eval: identity, consumes no draw."""
    if (not training) or p == 0:
        return x, None
    keep = (u >= p).astype(np.float64)
    return x * keep / (1.0 - p), keep


# --------------------------------------------------------------------------
# embracement (EmbraceNetMultimodal.py:63-90, :178-187)
# --------------------------------------------------------------------------
def modality_dropout_availabilities(B, u0, u_rows):
    """EmbraceNetMultimodal.py:178-182. Returns availabilities [B,2] or None."""
    if np.float32(u0) >= np.float32(0.5):
        target = np.round(np.asarray(u_rows, dtype=np.float32)).astype(np.int64)  # half-to-even, as torch.round
        av = np.zeros((B, 2), dtype=np.float32)
        av[np.arange(B), target] = 1.0
        return av
    return None


def embrace_probabilities(p_ffnn, B, availabilities=None):
    """EmbraceNetMultimodal.py:157,184 and :63-76 -- fp32 arithmetic, then the
    fp64 normalised cumulative threshold torch.multinomial's CPU path uses
    (SURVEY 8 a7): cum0 = double(p0) / (double(p0) + double(p1))."""
    sel = np.array([p_ffnn, 1.0 - p_ffnn], dtype=np.float32)[None, :].repeat(B, 0)
    av = np.ones((B, 2), dtype=np.float32) if availabilities is None else np.asarray(availabilities, dtype=np.float32)
    p = sel * av
    with np.errstate(invalid='ignore', divide='ignore'):
        p = p / p.sum(axis=-1, keepdims=True, dtype=np.float32)
        p0 = p[:, 0].astype(np.float64)
        p1 = p[:, 1].astype(np.float64)
        cum0 = p0 / (p0 + p1)
    return p, cum0


def embrace_select(u, cum0):
    """idx[b,c] = 1 if u[b,c] > cum0[b] else 0  (u == cum0 selects modality 0)."""
    return (np.asarray(u, dtype=np.float64) > cum0[:, None]).astype(np.int64)


# --------------------------------------------------------------------------
# loss / metrics (utils.py:80-140, training_models_multimodal.py:140-154)
# --------------------------------------------------------------------------
def loss_weights_from_labels(y):
    y = np.asarray(y).reshape(-1)
    pos = int((y == 1).sum())
    neg = int((y == 0).sum())
    pos_inv = 1 / pos if pos != 0 else 0
    neg_inv = 1 / neg if neg != 0 else 0
    return pos_inv / (neg_inv + pos_inv), neg_inv / (neg_inv + pos_inv)   # (w_pos, w_neg)


def weighted_ce(logits, y):
    """nn.CrossEntropyLoss(weight=[w_neg,w_pos]) on output.float(): fp32.
    Returns (loss fp32, dlogits fp64 [B,2])."""
    y = np.asarray(y).reshape(-1).astype(np.int64)
    w_pos, w_neg = loss_weights_from_labels(y)
    cw = np.array([w_neg, w_pos], dtype=np.float32)
    z = np.asarray(logits, dtype=np.float32)
    m = z.max(axis=1, keepdims=True)
    e = np.exp(z - m)
    s = e / e.sum(axis=1, keepdims=True, dtype=np.float32)
    lse = np.log(e.sum(axis=1, dtype=np.float32)) + m[:, 0]
    nll = lse - z[np.arange(len(y)), y]
    w = cw[y]
    wsum = w.sum(dtype=np.float32)
    loss = (w * nll).sum(dtype=np.float32) / wsum
    oh = np.zeros_like(z)
    oh[np.arange(len(y)), y] = 1.0
    dz = (w[:, None] * (s - oh) / wsum).astype(np.float32)
    return np.float32(loss), dz.astype(np.float64)


def confusion_counts(logits, y):
    pred = np.argmax(np.asarray(logits), axis=1)     # ties -> class 0, as torch.argmax
    y = np.asarray(y).reshape(-1)
    tp = int(((pred == 1) & (y == 1)).sum())
    fp = int(((pred == 1) & (y == 0)).sum())
    fn = int(((pred == 0) & (y == 1)).sum())
    tn = int(((pred == 0) & (y == 0)).sum())
    return tp, fp, fn, tn


def auprc_hard_from_counts(tp, fp, fn, tn):
    """Closed form of sklearn.average_precision_score(target, hard_pred)
    (utils.py:80-86). No positive targets -> NaN -> 0."""
    n = tp + fp + fn + tn
    npos = tp + fn
    if npos == 0:
        return 0.0
    if tp + fp == 0:
        return npos / n
    return (tp / npos) * (tp / (tp + fp)) + (fn / npos) * (npos / n)


def auprc_hard(logits, y):
    return auprc_hard_from_counts(*confusion_counts(logits, y))


def f1_precision_recall_from_counts(tp, fp, fn, tn):
    """Macro (precision, recall, F1) over the labels present in y or pred,
    zero_division=0 (utils.py:89-94)."""
    stats = []
    for (t, f_p, f_n) in ((tn, fn, fp), (tp, fp, fn)):   # class 0, class 1
        present = (t + f_p + f_n) > 0
        prec = t / (t + f_p) if (t + f_p) > 0 else 0.0
        rec = t / (t + f_n) if (t + f_n) > 0 else 0.0
        f1 = 2 * prec * rec / (prec + rec) if (prec + rec) > 0 else 0.0
        if present:
            stats.append((prec, rec, f1))
    return np.mean(np.array(stats), axis=0)


class EarlyStopping:
    """utils.py:23-67 (score must improve by >= delta; '<' counts as worse)."""

    def __init__(self, patience=4, delta=0):
        self.patience, self.delta = patience, delta
        self.counter, self.best_score, self.early_stop = 0, None, False

    def __call__(self, score):
        if self.best_score is None:
            self.best_score = score
        elif score < self.best_score + self.delta:
            self.counter += 1
            if self.counter >= self.patience:
                self.early_stop = True
        else:
            self.best_score = score
            self.counter = 0


# --------------------------------------------------------------------------
# whole-model forward / backward
# --------------------------------------------------------------------------
def _keys(spec):
    kind = spec.get('kind', 'embracenet')
    pre_f = 'FFNN.model.' if kind in ('embracenet', 'concatnet') else 'model.'
    pre_c = 'CNN.CNN_model.' if kind in ('embracenet', 'concatnet') else 'CNN_model.'
    return kind, pre_f, pre_c


def ffnn_forward(spec, P, x, draws, training, cache):
    _, pre_f, _ = _keys(spec)
    h = _q(np.asarray(x, dtype=np.float64))
    cache['ffnn'] = []
    for i, (u, p) in enumerate(zip(spec['ffnn_units'], spec['ffnn_dropout'])):
        pre = linear_fwd(h, _q(P[f'{pre_f}{3*i}.weight']), P[f'{pre_f}{3*i}.bias'])
        r = np.maximum(pre, 0.0)
        out, keep = dropout_fwd(r, draws['ffnn_drop'][i] if (training and p > 0) else None, p, training)
        cache['ffnn'].append((h, pre, keep, p))
        h = _q(out)
    return h


def ffnn_backward(spec, P, g, cache, G):
    _, pre_f, _ = _keys(spec)
    for i in reversed(range(len(spec['ffnn_units']))):
        h, pre, keep, p = cache['ffnn'][i]
        if keep is not None:
            g = g * keep / (1.0 - p)
        g = _q(g * (pre > 0))
        G[f'{pre_f}{3*i}.weight'] = g.T @ h
        G[f'{pre_f}{3*i}.bias'] = g.sum(axis=0)
        g = g @ _q(P[f'{pre_f}{3*i}.weight'])
    return g


def cnn_forward(spec, P, bases, draws, training, cache, new_buffers):
    _, _, pre_c = _keys(spec)
    cache['cnn'] = []
    x = None
    for i, (co, k, p) in enumerate(zip(spec['cnn_channels'], spec['cnn_kernels'], spec['cnn_dropout'])):
        W, b = P[f'{pre_c}{5*i}.weight'], P[f'{pre_c}{5*i}.bias']
        gamma, beta = P[f'{pre_c}{5*i+1}.weight'], P[f'{pre_c}{5*i+1}.bias']
        rm, rv = P[f'{pre_c}{5*i+1}.running_mean'], P[f'{pre_c}{5*i+1}.running_var']
        y = _q(onehot_conv_fwd(bases, W, b) if i == 0 else conv1d_fwd(x, _q(W), b))
        if training:
            z, xhat, rstd, nrm, nrv = bn_train_fwd(y, gamma, beta, rm, rv)
            new_buffers[f'{pre_c}{5*i+1}.running_mean'] = nrm
            new_buffers[f'{pre_c}{5*i+1}.running_var'] = nrv
            new_buffers[f'{pre_c}{5*i+1}.num_batches_tracked'] = \
                np.array(int(P[f'{pre_c}{5*i+1}.num_batches_tracked']) + 1, dtype=np.int64)
        else:
            z, xhat, rstd = bn_eval_fwd(y, gamma, beta, rm, rv), None, None
        pooled, pos = relu_maxpool_fwd(z)
        out, keep = dropout_fwd(pooled, draws['cnn_drop'][i] if (training and p > 0) else None, p, training)
        cache['cnn'].append((x, z, xhat, rstd, pos, keep, p))
        x = _q(out)
    return x.reshape(x.shape[0], -1)                  # channel-major flatten: c*L_last + l


def cnn_backward(spec, P, g_flat, bases, cache, G):
    _, _, pre_c = _keys(spec)
    n = len(spec['cnn_channels'])
    Lp_last = cnn_lengths(spec['cnn_kernels'])[-1][1]
    g = g_flat.reshape(g_flat.shape[0], spec['cnn_channels'][-1], Lp_last)
    for i in reversed(range(n)):
        x, z, xhat, rstd, pos, keep, p = cache['cnn'][i]
        if keep is not None:
            g = g * keep / (1.0 - p)
        dz = relu_maxpool_bwd(z, pos, g)
        dy, dgamma, dbeta = bn_train_bwd(dz, xhat, rstd, P[f'{pre_c}{5*i+1}.weight'])
        G[f'{pre_c}{5*i+1}.weight'] = dgamma
        G[f'{pre_c}{5*i+1}.bias'] = dbeta
        db = dy.sum(axis=(0, 2))
        dy = _q(dy)
        W = P[f'{pre_c}{5*i}.weight']
        if i == 0:
            dW, _ = onehot_conv_bwd(bases, dy, W.shape[2])
            g = None
        else:
            g, dW, _ = conv1d_bwd(x, _q(W), dy)
            g = _q(g)
        G[f'{pre_c}{5*i}.weight'] = dW
        G[f'{pre_c}{5*i}.bias'] = db


def forward(spec, P, x_ffnn, bases, draws=None, training=False, availabilities=None,
            embracenet_dropout=True):
    """EmbraceNetMultimodal.forward (:159-193) / FFNN.forward / CNN.forward.
    Returns (logits [B,2] fp64, cache). cache['idx'] is the modality index map,
    cache['new_buffers'] the updated BatchNorm buffers (training)."""
    kind, pre_f, pre_c = _keys(spec)
    draws = draws or {}
    cache = {'new_buffers': {}}
    if kind == 'ffnn':
        h = ffnn_forward(spec, P, x_ffnn, draws, training, cache)
        n = len(spec['ffnn_units'])
        cache['head_in'] = h
        return linear_fwd(h, _q(P[f'{pre_f}{3*n}.weight']), P[f'{pre_f}{3*n}.bias']), cache
    if kind == 'cnn':
        f = cnn_forward(spec, P, bases, draws, training, cache, cache['new_buffers'])
        h1 = _q(linear_fwd(f, _q(P['last_layer1.weight']), P['last_layer1.bias']))
        h2 = _q(linear_fwd(h1, _q(P['last_layer2.weight']), P['last_layer2.bias']))
        cache['head'] = (f, h1, h2)
        return linear_fwd(h2, _q(P['last_output.weight']), P['last_output.bias']), cache

    B = x_ffnn.shape[0]
    xf = ffnn_forward(spec, P, x_ffnn, draws, training, cache)
    xc = cnn_forward(spec, P, bases, draws, training, cache, cache['new_buffers'])
    if kind == 'concatnet':      # ConcatNetMultimodal.forward (:65-82): post(cat(FFNN(x1), CNN(x2)))
        h = np.concatenate([xf, xc], axis=1)
        cache.update(xf=xf, xc=xc)
        cache['post'] = []
        for i, (u, p) in enumerate(zip(spec['post_units'], spec['post_dropout'])):
            pre = linear_fwd(h, _q(P[f'post.{3*i}.weight']), P[f'post.{3*i}.bias'])
            r = np.maximum(pre, 0.0)
            out, keep = dropout_fwd(r, draws['post_drop'][i] if (training and p > 0) else None, p, training)
            cache['post'].append((h, pre, keep, p))
            h = _q(out)
        n = len(spec['post_units'])
        cache['head_in'] = h
        return linear_fwd(h, _q(P[f'post.{3*n}.weight']), P[f'post.{3*n}.bias']), cache
    if training and embracenet_dropout:
        av = modality_dropout_availabilities(B, draws['modal_u0'], draws.get('modal_rows'))
        if av is not None:
            availabilities = av
    p32, cum0 = embrace_probabilities(spec['p_ffnn'], B, availabilities)
    idx = embrace_select(draws['embrace_u'], cum0)
    pre0 = linear_fwd(xf, _q(P['embracenet.docking_0.weight']), P['embracenet.docking_0.bias'])
    pre1 = linear_fwd(xc, _q(P['embracenet.docking_1.weight']), P['embracenet.docking_1.bias'])
    d0, d1 = np.maximum(pre0, 0.0), np.maximum(pre1, 0.0)
    e = _q(np.where(idx == 1, d1, d0))
    cache.update(xf=xf, xc=xc, pre0=pre0, pre1=pre1, idx=idx, cum0=cum0, p32=p32, e=e,
                 availabilities=availabilities)
    h = e
    cache['post'] = []
    for i, (u, p) in enumerate(zip(spec['post_units'], spec['post_dropout'])):
        pre = linear_fwd(h, _q(P[f'post.{3*i}.weight']), P[f'post.{3*i}.bias'])
        r = np.maximum(pre, 0.0)
        out, keep = dropout_fwd(r, draws['post_drop'][i] if (training and p > 0) else None, p, training)
        cache['post'].append((h, pre, keep, p))
        h = _q(out)
    n = len(spec['post_units'])
    cache['head_in'] = h
    logits = linear_fwd(h, _q(P[f'post.{3*n}.weight']), P[f'post.{3*n}.bias'])
    return logits, cache


def backward(spec, P, dlogits, bases, cache):
    """Gradients of every parameter given dL/dlogits [B,2] (training forward)."""
    kind, pre_f, pre_c = _keys(spec)
    G = {}
    g = np.asarray(dlogits, dtype=np.float64)
    if kind == 'ffnn':
        n = len(spec['ffnn_units'])
        G[f'{pre_f}{3*n}.weight'] = g.T @ cache['head_in']
        G[f'{pre_f}{3*n}.bias'] = g.sum(axis=0)
        ffnn_backward(spec, P, g @ _q(P[f'{pre_f}{3*n}.weight']), cache, G)
        return G
    if kind == 'cnn':
        f, h1, h2 = cache['head']
        for name, inp in (('last_output', h2), ('last_layer2', h1), ('last_layer1', f)):
            G[f'{name}.weight'] = g.T @ inp
            G[f'{name}.bias'] = g.sum(axis=0)
            g = _q(g @ _q(P[f'{name}.weight']))
        cnn_backward(spec, P, g, bases, cache, G)
        return G
    n = len(spec['post_units'])
    G[f'post.{3*n}.weight'] = g.T @ cache['head_in']
    G[f'post.{3*n}.bias'] = g.sum(axis=0)
    g = g @ _q(P[f'post.{3*n}.weight'])
    for i in reversed(range(n)):
        h, pre, keep, p = cache['post'][i]
        if keep is not None:
            g = g * keep / (1.0 - p)
        g = _q(g * (pre > 0))
        G[f'post.{3*i}.weight'] = g.T @ h
        G[f'post.{3*i}.bias'] = g.sum(axis=0)
        g = g @ _q(P[f'post.{3*i}.weight'])
    if kind == 'concatnet':
        g = _q(g)
        fo = cache['xf'].shape[1]
        ffnn_backward(spec, P, g[:, :fo], cache, G)
        cnn_backward(spec, P, g[:, fo:], bases, cache, G)
        return G
    idx = cache['idx']
    dd0 = _q(g * (idx == 0) * (cache['pre0'] > 0))
    dd1 = _q(g * (idx == 1) * (cache['pre1'] > 0))
    G['embracenet.docking_0.weight'] = dd0.T @ cache['xf']
    G['embracenet.docking_0.bias'] = dd0.sum(axis=0)
    G['embracenet.docking_1.weight'] = dd1.T @ cache['xc']
    G['embracenet.docking_1.bias'] = dd1.sum(axis=0)
    ffnn_backward(spec, P, dd0 @ _q(P['embracenet.docking_0.weight']), cache, G)
    cnn_backward(spec, P, _q(dd1 @ _q(P['embracenet.docking_1.weight'])), bases, cache, G)
    return G


# --------------------------------------------------------------------------
# optimizers (training_models_multimodal.py:318-325; torch/optim/{adam,rmsprop,nadam}.py;
# timm.optim.Nadam is absent from the image -- torch NAdam(momentum_decay=4e-3) is the
# same published rule, see SURVEY 8 c)
# --------------------------------------------------------------------------
def opt_init(P, kind):
    st = {'t': 0, 'mu_product': 1.0, 'kind': kind, 'm': {}, 'v': {}}
    for k, v in P.items():
        if not is_buffer(k):
            st['m'][k] = np.zeros_like(v, dtype=np.float64)
            st['v'][k] = np.zeros_like(v, dtype=np.float64)
    return st


def opt_step(P, G, st, lr, wd, decoupled=False, b1=0.9, b2=0.999, eps=1e-8, alpha=0.99, psi=4e-3):
    """In-place update of P. kind in {'adam','rmsprop','nadam'}; `decoupled`
    turns Adam into AdamW (north_star's variant; not what the reference runs)."""
    st['t'] += 1
    t = st['t']
    kind = st['kind']
    if kind == 'nadam':
        mu = b1 * (1.0 - 0.5 * (0.96 ** (t * psi)))
        mu_next = b1 * (1.0 - 0.5 * (0.96 ** ((t + 1) * psi)))
        st['mu_product'] *= mu
        mp = st['mu_product']
    for k in st['m']:
        p, g = P[k], G[k]
        if decoupled:
            p *= (1.0 - lr * wd)
        else:
            g = g + wd * p
        if kind == 'rmsprop':
            st['v'][k] = alpha * st['v'][k] + (1 - alpha) * g * g
            p -= lr * g / (np.sqrt(st['v'][k]) + eps)
            continue
        st['m'][k] = b1 * st['m'][k] + (1 - b1) * g
        st['v'][k] = b2 * st['v'][k] + (1 - b2) * g * g
        m, v = st['m'][k], st['v'][k]
        if kind == 'adam':
            denom = np.sqrt(v) / np.sqrt(1 - b2 ** t) + eps
            p -= (lr / (1 - b1 ** t)) * m / denom
        elif kind == 'nadam':
            denom = np.sqrt(v / (1 - b2 ** t)) + eps
            p -= lr * (1 - mu) / (1 - mp) * g / denom
            p -= lr * mu_next / (1 - mp * mu_next) * m / denom
        else:
            raise ValueError(kind)


def train_step(spec, P, x_ffnn, bases, y, draws, opt_state=None, lr=1e-3, wd=0.0, decoupled=False):
    """One iteration of the loop body at training_models_multimodal.py:132-162.
    Mutates P (and BN buffers). Returns dict(logits, loss, grads, idx, auprc)."""
    logits, cache = forward(spec, P, x_ffnn, bases, draws, training=True)
    loss, dlogits = weighted_ce(logits, y)
    G = backward(spec, P, dlogits, bases, cache)
    for k, v in cache['new_buffers'].items():
        P[k] = v
    if opt_state is not None:
        opt_step(P, G, opt_state, lr, wd, decoupled)
    return dict(logits=logits, loss=loss, grads=G, idx=cache.get('idx'),
                auprc=auprc_hard(logits, y), counts=confusion_counts(logits, y))


def predict_proba(spec, P, x_ffnn, bases, embrace_u, availabilities=None):
    """EmbraceNetMultimodal_NoTrain.forward (:180-214): eval forward + softmax;
    returns P(class 1) per row (the caller takes element [1], visual.py:290-293)."""
    logits, _ = forward(spec, P, x_ffnn, bases, {'embrace_u': embrace_u}, training=False,
                        availabilities=availabilities)
    m = logits.max(axis=1, keepdims=True)
    e = np.exp(logits - m)
    return (e / e.sum(axis=1, keepdims=True))[:, 1]


def make_draws(spec, B, seed, force_modal=None):
    """Synthetic explicit draws for one training forward (order documented in SURVEY quirk 8)."""
    rs = np.random.RandomState(seed)
    kind = spec.get('kind', 'embracenet')
    d = {'ffnn_drop': [], 'cnn_drop': [], 'post_drop': []}
    if kind in ('embracenet', 'concatnet', 'ffnn'):
        for u in spec['ffnn_units']:
            d['ffnn_drop'].append(rs.random_sample((B, u)).astype(np.float32))
    if kind in ('embracenet', 'concatnet', 'cnn'):
        for co, (_, Lp) in zip(spec['cnn_channels'], cnn_lengths(spec['cnn_kernels'])):
            d['cnn_drop'].append(rs.random_sample((B, co, Lp)).astype(np.float32))
    if kind == 'concatnet':
        for u in spec['post_units']:
            d['post_drop'].append(rs.random_sample((B, u)).astype(np.float32))
    if kind == 'embracenet':
        u0 = np.float32(rs.random_sample())
        if force_modal is not None:
            u0 = np.float32(0.75 if force_modal else 0.25)
        d['modal_u0'] = u0
        d['modal_rows'] = rs.random_sample(B).astype(np.float32)
        d['embrace_u'] = rs.random_sample((B, spec['C']))
        for u in spec['post_units']:
            d['post_drop'].append(rs.random_sample((B, u)).astype(np.float32))
    return d
