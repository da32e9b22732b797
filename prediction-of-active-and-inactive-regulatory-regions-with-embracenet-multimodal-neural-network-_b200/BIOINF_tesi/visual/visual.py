"""Mirror of `Compare_Models_Result` (BIOINF_tesi/visual/visual.py:250-404): per fold, score every region with the base
model(s) and the comparison models rebuilt from the reference's per-fold checkpoints
(`{cell}_{model}_{task}_{fold}_test_.pt` in the working directory, keys `model_state_dict` / `model_params`), and compare
the paired predictions with a Wilcoxon signed-rank test.

What changes: the batch-1 Python loop of `get_model_predictions` (visual.py:282-293, up to 163 k forward calls per
model) becomes one batched engine forward per 65 536 regions (`EngineModule.predict_scores`).  What stays: class and
method names, the constructor-by-model-name table, file naming, `load_state_dict` of the reference's state dict,
the `[1]` element convention (P(class 1) for the softmax twins, logit 1 for ConcatNetMultimodal_NoTrain whose softmax the
reference loses to a typo), the `pval_results_dict_{task}.pickle` cache, scipy's `wilcoxon`, the printed verdicts.
The plots and tables of the rest of visual.py stay with the reference (CPU pandas/matplotlib, not on the hot path).

Data access: the reference pulls `X_1` / `X_2` out of `Build_DataLoader_Pipeline(path_name=f'{task}.pickle')`
(visual.py:359-368), a pandas/sklearn pipeline over pickles that are not part of the repository.  Here `__call__` takes
`data(task, cell_line) -> (X_1, X_2)` instead: `X_1` [n, F] features, `X_2` [n, 4, 256] one-hot or [n, 256] base codes
(`data_pipe.wire.pack_sequences`); `set_data(X_1, X_2)` serves direct use of `get_model_predictions`."""
import os
import pickle
import warnings
from collections import OrderedDict, defaultdict

import numpy as np
import torch
from scipy.stats import wilcoxon

from ..models import (CNN_NoTrain, ConcatNetMultimodal_NoTrain, EmbraceNetMultimodal_NoTrain, FFNN_NoTrain)

TASKS = ['active_E_vs_inactive_E', 'active_P_vs_inactive_P', 'active_E_vs_active_P', 'inactive_E_vs_inactive_P',
         'active_EP_vs_inactive_rest']
CELL_LINES = ['A549', 'GM12878', 'H1', 'HEK293', 'HEPG2', 'K562', 'MCF7']
UNIMODAL_NETWORKS_SEQ = ('CNN')          # plain strings, as in the reference: `startswith` is a prefix test
UNIMODAL_NETWORKS_NOSEQ = ('FFNN')
MULTIMODAL_NETWORKS = ('EmbraceNetMultimodal', 'ConcatNetMultimodal')


def dd():
    return defaultdict(dict)


class Compare_Models_Result():

    def __init__(self, precision=None, batch_size=65536):
        self.models_dict = {'EmbraceNetMultimodal': EmbraceNetMultimodal_NoTrain,
                            'EmbraceNetMultimodal_augmentation': EmbraceNetMultimodal_NoTrain,
                            'ConcatNetMultimodal': ConcatNetMultimodal_NoTrain,
                            'FFNN': FFNN_NoTrain,
                            'CNN': CNN_NoTrain}
        self.pval_dict = defaultdict(dd)
        self.precision, self.batch_size = precision, batch_size
        self.X_1 = self.X_2 = None
        warnings.filterwarnings("ignore")

    def set_data(self, X_1, X_2):
        """X_1: [n, F] array / tensor (or a sequence of [1, F] tensors, the reference's pandas Series of rows);
        X_2: [n, 4, 256] one-hot, [n, 256] base codes, or a sequence of [1, 4, 256] tensors."""
        def stack(x):
            if isinstance(x, (np.ndarray, torch.Tensor)):
                return torch.as_tensor(x)
            return torch.cat([torch.as_tensor(r) for r in x], dim=0)
        self.X_1, self.X_2 = stack(X_1), stack(X_2)
        if len(self.X_1) != len(self.X_2):
            raise ValueError(f'X_1 has {len(self.X_1)} rows, X_2 {len(self.X_2)}')

    def get_model_predictions(self, cell_line, task, model, n_iteration, device, draws=None):
        """visual.py:263-295.  Returns a CPU fp32 tensor [n]: element [1] of the model's output for every region."""
        model_ = self.models_dict[model]
        kw = dict(precision=self.precision)
        if model == 'CNN':
            model_ = model_(cell_line, task, n_iteration, device, **kw)
        elif model.endswith('augmentation'):
            model_ = model_(cell_line, task, n_iteration, self.X_1.shape[1], device=device, augmentation=True, **kw)
        else:
            model_ = model_(cell_line, task, n_iteration, self.X_1.shape[1], device=device, **kw)

        state_dict = torch.load(f'{cell_line}_{model}_{task}_{n_iteration}_test_.pt', map_location='cpu', weights_only=False)
        model_.load_state_dict(state_dict['model_state_dict'])
        model_.double().to(device)
        for p in model_.parameters():
            p.requires_grad = False
        model_.eval()

        if model.startswith(UNIMODAL_NETWORKS_NOSEQ):
            output = model_.predict_scores(self.X_1, None, batch_size=self.batch_size)
        elif model.startswith(UNIMODAL_NETWORKS_SEQ):
            output = model_.predict_scores(None, self.X_2, batch_size=self.batch_size)
        elif model.startswith(MULTIMODAL_NETWORKS):
            column = 'logit' if model.startswith('ConcatNetMultimodal') else 'prob'
            output = model_.predict_scores(self.X_1, self.X_2, batch_size=self.batch_size, draws=draws, column=column)
        else:
            raise ValueError(f"unknown model name {model!r}: use one among {list(self.models_dict)}")
        return output.float().cpu()

    def print_model_difference(self, p_val=0.05):
        """visual.py:298-325: a comparison model counts as different from a base model when p < p_val in at least two folds.
        Fills `counter_dict[task][cell_line][base][comparison]` (number of significant folds) and prints the verdicts."""
        counts = {}
        for task, cells in self.pval_dict.items():
            for cell, folds in cells.items():
                for per_base in folds.values():
                    for base, others in per_base.items():
                        slot = counts.setdefault(task, {}).setdefault(cell, {}).setdefault(base, {})
                        for other, p in others.items():
                            slot[other] = slot.get(other, 0) + (0 if p >= p_val else 1)     # as the reference: a NaN p-value counts as different
        self.counter_dict = counts
        for task, cells in counts.items():
            print(f'\n\n================ TASK: {task} ================')
            for cell, bases in cells.items():
                print(f'\n\n{cell}')
                for base, others in bases.items():
                    print(f'\n\nBASE MODEL: {base}\n')
                    for other, n_sig in others.items():
                        print(f'{other} ===> different: {n_sig >= 2}')

    def __call__(self, device, base_model='EmbraceNetMultimodal', comparison_models=['FFNN', 'CNN', 'ConcatNetMultimodal'],
                 augmentation_base_model=True, n_folds=3, cell_lines=CELL_LINES, tasks=TASKS, pval_dict=None, data=None):
        if pval_dict:
            self.pval_dict = pval_dict
        else:
            if data is None:
                raise ValueError("data=callable(task, cell_line) -> (X_1, X_2) is required: the reference's "
                                 "Build_DataLoader_Pipeline pickles are not part of this package")
            base_model = [base_model] if isinstance(base_model, str) else list(base_model)
            comparison_models = [comparison_models] if isinstance(comparison_models, str) else list(comparison_models)
            tasks = [tasks] if isinstance(tasks, str) else tasks
            cell_lines = [cell_lines] if isinstance(cell_lines, str) else cell_lines
            MODELS = comparison_models + base_model
            if augmentation_base_model:
                MODELS += [f'{base_model[0]}_augmentation']
                base_model += [f'{base_model[0]}_augmentation']
            for task in tasks:
                if os.path.exists(f'pval_results_dict_{task}.pickle'):
                    with open(f'pval_results_dict_{task}.pickle', 'rb') as fin:
                        self.pval_dict = defaultdict(lambda: defaultdict(dict), pickle.load(fin))
                for cell_line in cell_lines:
                    self.set_data(*data(task, cell_line))
                    for i in range(1, n_folds + 1):
                        self.pval_dict[task][cell_line][str(i)] = defaultdict(dd)
                        preds = {}                    # every model is scored once per fold (the reference re-scores the
                        for b_model in base_model:    # comparison models for each base model)
                            for m in [b_model] + [c for c in MODELS if c != b_model]:
                                if m not in preds:
                                    preds[m] = self.get_model_predictions(cell_line, task, m, i, device).double().numpy()
                            for c_model in MODELS:
                                if c_model != b_model:
                                    self.pval_dict[task][cell_line][str(i)][b_model][c_model] = wilcoxon(preds[b_model], preds[c_model])[1]
                    self.X_1 = self.X_2 = None
                    with open(f'pval_results_dict_{task}.pickle', 'wb') as fout:
                        pickle.dump(_plain(self.pval_dict), fout)
        self.print_model_difference()
        return self.pval_dict


def _plain(d):
    """defaultdicts built from lambdas do not pickle: store ordinary (ordered) dicts, as the reference's OrderedDict(...) does
    at the top level."""
    return OrderedDict((k, _plain(v)) for k, v in d.items()) if isinstance(d, dict) else d
