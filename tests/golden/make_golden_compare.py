"""Golden fixture for SURVEY.md 8 (f4): checkpoint interop + the Compare_Models_Result predict loop.

Runs the UNMODIFIED reference (visual/visual.py:250-295) in the build container:
  * each of the five model names of `Compare_Models_Result.models_dict` is built by the reference's own TRAINING class,
    its `state_dict()` is written the way `Kfold_CV*` writes the per-fold files (`{'model_state_dict', 'model_params'}`,
    training_models_multimodal.py:639-642) -> `tests/golden/ckpt/*.pt` (committed: they are "real reference .pt files");
  * `Compare_Models_Result.get_model_predictions` (the batch-1 Python loop through the `_NoTrain` twins) scores N regions
    with the multinomial draws replayed from a logged tensor;
  * `scipy.stats.wilcoxon` p-values between the base models and the comparison models, as `__call__` computes them.

python -m tests.golden.make_golden_compare
"""
import os
import shutil
import sys
import tempfile

import numpy as np
import pandas as pd
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import embracenet_oracle as O                      # noqa: E402
from tests.golden import ref_harness as RH                     # noqa: E402
from tests.golden.cases import CASES, make_inputs              # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
CELL, TASK, FOLD, N, F = 'A549', 'active_E_vs_inactive_E', 1, 24, 12


def compare_specs():
    emb = dict(CASES['small2']['spec'])                                       # F = 12
    emb_aug = dict(emb)                  # same architecture, weights = base + small noise: a non-degenerate p-value
    concat = dict(CASES['concat_small']['spec'], F=F)
    ffnn = dict(kind='ffnn', F=F, ffnn_units=[16, 8], ffnn_dropout=[0.2, 0.0])
    cnn = dict(kind='cnn', cnn_channels=[8, 8, 8, 4], cnn_kernels=[11, 5, 5, 5], cnn_dropout=[0.2, 0.4, 0.0, 0.0])   # CNN_out = 32
    return {'EmbraceNetMultimodal': (emb, 5101), 'EmbraceNetMultimodal_augmentation': (emb_aug, 5102),
            'ConcatNetMultimodal': (concat, 5103), 'FFNN': (ffnn, 5104), 'CNN': (cnn, 5105)}


def main():
    M = RH.import_reference()
    import BIOINF_tesi.visual.visual as V
    specs = compare_specs()
    x1, bases, _ = make_inputs(dict(F=F), N, 5100)
    u = {name: np.random.RandomState(seed + 50).random_sample((N, spec['C'])) for name, (spec, seed) in specs.items() if spec.get('kind', 'embracenet') == 'embracenet'}
    ckdir = os.path.join(HERE, 'ckpt')
    os.makedirs(ckdir, exist_ok=True)
    cwd = os.getcwd()
    preds = {}
    with tempfile.TemporaryDirectory() as td:
        os.chdir(td)
        try:
            for name, (spec, seed) in specs.items():
                P = O.init_params(spec, 5101 if name.endswith('augmentation') else seed)
                if name.endswith('augmentation'):
                    rs = np.random.RandomState(seed)
                    P = {k: (v + 0.02 * rs.standard_normal(np.shape(v)) if k.startswith('post') and np.ndim(v) else v) for k, v in P.items()}
                model = RH.build_reference_model(M, spec, P)
                fn = f'{CELL}_{name}_{TASK}_{FOLD}_test_.pt'
                torch.save({'model_state_dict': model.state_dict(), 'model_params': RH.spec_to_trial_params(spec)}, fn)
                shutil.copy(fn, os.path.join(ckdir, fn))
            cmp_ = V.Compare_Models_Result()
            cmp_.X_1 = pd.Series([torch.from_numpy(x1[i:i + 1]).double() for i in range(N)])
            cmp_.X_2 = pd.Series([torch.from_numpy(O.onehot_from_bases(bases[i:i + 1])).double() for i in range(N)])
            for name in specs:
                queue = [('multinomial', u[name][i:i + 1]) for i in range(N)] if name in u else []
                with RH.replay_draws(queue):
                    preds[name] = cmp_.get_model_predictions(CELL, TASK, name, FOLD, 'cpu').numpy().astype(np.float64)
        finally:
            os.chdir(cwd)
    from scipy.stats import wilcoxon
    out = {f'pred_{k}': v for k, v in preds.items()}
    out.update({f'u_{k}': v for k, v in u.items()})
    for b in ('EmbraceNetMultimodal', 'EmbraceNetMultimodal_augmentation'):
        for c in specs:
            if c != b:
                out[f'pval_{b}__{c}'] = np.array(wilcoxon(preds[b], preds[c])[1])
    np.savez_compressed(os.path.join(HERE, 'compare_models.npz'), x1=x1, bases=bases, **out)
    for k, v in preds.items():
        print(k, v[:4])
    print({k: float(v) for k, v in out.items() if k.startswith('pval_')})


if __name__ == '__main__':
    main()
