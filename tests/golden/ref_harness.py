"""Import the UNMODIFIED reference (read-only at /root/reference) in the build
container and run it with explicit, replayed random draws.

Used only by make_golden.py (fixture generation) and by tests that are skipped
when /root/reference is absent (it does not exist on the GPU box).

The reference imports seaborn / matplotlib / optuna / timm / botorch /
sqlalchemy / miceforest / imblearn at module import time; none is installed
here and none is on the hot path, so permissive stub modules are registered
first ("Import recipe").
"""
import os
import sys
import types
import contextlib

import numpy as np
import torch
import torch.nn.functional as F

REF_ROOT = os.environ.get('EMB_REFERENCE_ROOT', '/root/reference')
_STUBS = ['seaborn', 'matplotlib', 'matplotlib.pylab', 'matplotlib.pyplot', 'optuna', 'optuna.integration',
          'optuna.samplers', 'optuna.pruners', 'optuna.trial', 'timm', 'timm.optim', 'botorch', 'sqlalchemy',
          'miceforest', 'imblearn', 'imblearn.over_sampling', 'imblearn.under_sampling', 'imblearn.pipeline',
          'barplots', 'tqdm.notebook', 'statannot', 'statannotations', 'statannotations.Annotator']


class _Stub(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith('__') and name.endswith('__'):
            raise AttributeError(name)
        child = _Stub(f'{self.__name__}.{name}')
        setattr(self, name, child)
        return child

    def __call__(self, *a, **k):
        return _Stub(self.__name__ + '()')

    def __iter__(self):
        return iter(())


def reference_available():
    return os.path.isdir(os.path.join(REF_ROOT, 'BIOINF_tesi'))


def import_reference():
    """Returns the reference's BIOINF_tesi.models package."""
    for name in _STUBS:
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                sys.modules[name] = _Stub(name)
    # the repo also ships a drop-in package called BIOINF_tesi: make sure the reference wins here
    for k in [k for k in sys.modules if k == 'BIOINF_tesi' or k.startswith('BIOINF_tesi.')]:
        mod = sys.modules[k]
        if REF_ROOT not in (getattr(mod, '__file__', '') or ''):
            del sys.modules[k]
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    import BIOINF_tesi.models as M
    assert REF_ROOT in M.__file__, M.__file__
    return M


class FixedTrial:
    """Duck-typed optuna trial returning fixed values."""

    def __init__(self, params):
        self.params = dict(params)
        self.number = 0

    def suggest_int(self, name, lo, hi):
        return int(self.params[name])

    def suggest_categorical(self, name, choices):
        return self.params[name]

    def suggest_float(self, name, lo, hi, **kw):
        return float(self.params[name])

    suggest_loguniform = suggest_float

    def report(self, *a, **k):
        pass

    def should_prune(self):
        return False


def spec_to_trial_params(spec):
    """oracle spec dict -> the reference's trial parameter names."""
    kind = spec.get('kind', 'embracenet')
    pf = 'FFNN_' if kind in ('embracenet', 'concatnet') else ''
    pc = 'CNN_' if kind in ('embracenet', 'concatnet') else ''
    tp = {}
    if kind in ('embracenet', 'concatnet', 'ffnn'):
        tp[f'{pf}n_layers'] = len(spec['ffnn_units'])
        for i, (u, p) in enumerate(zip(spec['ffnn_units'], spec['ffnn_dropout'])):
            tp[f'{pf}n_units_l{i}'] = u
            tp[f'{pf}dropout_l{i}'] = p
    if kind in ('embracenet', 'concatnet', 'cnn'):
        tp[f'{pc}n_layers'] = len(spec['cnn_channels'])
        for i, (c, k, p) in enumerate(zip(spec['cnn_channels'], spec['cnn_kernels'], spec['cnn_dropout'])):
            tp[f'{pc}out_channels_l{i}'] = c
            tp[f'{pc}kernel_size_l{i}'] = k
            tp[f'{pc}dropout_l{i}'] = p
    if kind == 'embracenet':
        tp['EMBRACENET_embracement_size'] = spec['C']
        tp['n_post_layers'] = len(spec['post_units'])
        for i, (u, p) in enumerate(zip(spec['post_units'], spec['post_dropout'])):
            tp[f'EMBRACENET_n_units_l{i}'] = u
            tp[f'EMBRACENET_dropout_l{i}'] = p
        tp['selection_probabilities_FFNN'] = spec['p_ffnn']
    if kind == 'concatnet':
        tp['CONCATNET_n_post_layers'] = len(spec['post_units'])
        for i, (u, p) in enumerate(zip(spec['post_units'], spec['post_dropout'])):
            tp[f'CONCATNET_n_units_l{i}'] = u
            tp[f'CONCATNET_dropout_l{i}'] = p
    return tp


def build_reference_model(M, spec, P):
    """Instantiate the reference class for `spec` and load oracle-format params P."""
    kind = spec.get('kind', 'embracenet')
    trial = FixedTrial(spec_to_trial_params(spec))
    if kind == 'embracenet':
        model = M.EmbraceNetMultimodal(trial, cell_line='A549', task='active_E_vs_inactive_E', device='cpu',
                                       in_features_FFNN=spec['F'])
    elif kind == 'concatnet':
        import importlib
        CN = importlib.import_module('BIOINF_tesi.models.ConcatNetMultimodal')
        model = CN.ConcatNetMultimodal(trial, cell_line='A549', task='active_E_vs_inactive_E', in_features_FFNN=spec['F'], device='cpu')
    elif kind == 'ffnn':
        model = M.FFNN(trial, spec['F'], device='cpu')
    else:
        model = M.CNN(trial, device='cpu')
    sd = {k: torch.from_numpy(np.asarray(v).copy()) for k, v in P.items()}
    model = model.double()          # cast first: loading fp64 values into fp32 params would round them
    model.load_state_dict(sd, strict=True)
    return model


@contextlib.contextmanager
def replay_draws(queue):
    """Patch F.dropout / torch.multinomial / torch.rand so that every random draw the
    reference makes is taken, in order, from `queue` (list of (tag, array))."""
    log = []
    orig_dropout, orig_multi, orig_rand = F.dropout, torch.multinomial, torch.rand

    def pop(tag, shape):
        assert queue, f'draw queue exhausted at {tag} {shape}'
        qtag, arr = queue.pop(0)
        assert qtag == tag, f'draw order mismatch: reference asks {tag}{tuple(shape)}, queue has {qtag}'
        arr = np.asarray(arr)
        assert tuple(arr.shape) == tuple(shape), (tag, arr.shape, tuple(shape))
        log.append((tag, tuple(shape)))
        return arr

    def dropout(input, p=0.5, training=True, inplace=False):
        if (not training) or p == 0:
            return input
        u = torch.from_numpy(pop('dropout', input.shape).astype(np.float64))
        return input * (u >= p).to(input.dtype) / (1.0 - p)

    def multinomial(probs, num_samples, replacement=False, *, generator=None, out=None):
        assert replacement
        u = torch.from_numpy(pop('multinomial', (probs.shape[0], num_samples)).astype(np.float64))
        p = probs.double()
        cum0 = p[:, 0] / (p[:, 0] + p[:, 1])          # rule verified against the real op in make_golden.py
        return (u > cum0[:, None]).long()

    def rand(*size, **kw):
        shape = tuple(size[0]) if len(size) == 1 and isinstance(size[0], (list, tuple, torch.Size)) else tuple(size)
        return torch.from_numpy(pop('rand', shape).astype(np.float32))

    F.dropout, torch.multinomial, torch.rand = dropout, multinomial, rand
    torch.nn.functional.dropout = dropout
    try:
        yield log
    finally:
        F.dropout, torch.multinomial, torch.rand = orig_dropout, orig_multi, orig_rand
        torch.nn.functional.dropout = orig_dropout


def draws_to_queue(spec, draws, training=True):
    """Order in which one reference forward consumes draws."""
    kind = spec.get('kind', 'embracenet')
    q = []
    if training and kind in ('embracenet', 'concatnet', 'ffnn'):
        for i, p in enumerate(spec['ffnn_dropout']):
            if p > 0:
                q.append(('dropout', draws['ffnn_drop'][i]))
    if training and kind in ('embracenet', 'concatnet', 'cnn'):
        for i, p in enumerate(spec['cnn_dropout']):
            if p > 0:
                q.append(('dropout', draws['cnn_drop'][i]))
    if kind == 'concatnet' and training:
        for i, p in enumerate(spec['post_dropout']):
            if p > 0:
                q.append(('dropout', draws['post_drop'][i]))
    if kind == 'embracenet':
        if training:
            q.append(('rand', np.asarray([draws['modal_u0']], dtype=np.float32)))
            if np.float32(draws['modal_u0']) >= np.float32(0.5):
                q.append(('rand', draws['modal_rows']))
        q.append(('multinomial', draws['embrace_u']))
        if training:
            for i, p in enumerate(spec['post_dropout']):
                if p > 0:
                    q.append(('dropout', draws['post_drop'][i]))
    return q
