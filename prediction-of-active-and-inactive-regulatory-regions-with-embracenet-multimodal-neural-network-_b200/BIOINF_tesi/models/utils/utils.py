"""Mirror of the helpers of BIOINF_tesi/models/utils/utils.py that sit on the hot path."""
import re
from collections import defaultdict

import numpy as np
import torch
import torch.nn as nn


class EarlyStopping:
    """utils.py:23-67: stop when the score has not improved by `delta` for `patience` calls (equal counts as improved)."""

    def __init__(self, patience=4, verbose=False, delta=0, trace_func=print):
        self.patience, self.verbose, self.delta, self.trace_func = patience, verbose, delta, trace_func
        self.counter, self.best_score, self.early_stop = 0, None, False

    def __call__(self, score):
        if self.best_score is None:
            self.best_score = score
        elif score < self.best_score + self.delta:
            self.counter += 1
            self.trace_func(f'EarlyStopping counter: {self.counter} out of {self.patience}')
            if self.counter >= self.patience:
                self.early_stop = True
        else:
            self.best_score = score
            self.counter = 0


def accuracy(output, target):
    """utils.py:71-77."""
    return (torch.argmax(output, dim=1) == target.reshape(-1).to(output.device)).float().mean()


def confusion_counts(output, target):
    pred = torch.argmax(output, dim=1).reshape(-1)
    t = torch.as_tensor(target).reshape(-1).to(pred.device)
    tp = int(((pred == 1) & (t == 1)).sum())
    fp = int(((pred == 1) & (t != 1)).sum())
    fn = int(((pred != 1) & (t == 1)).sum())
    tn = int(((pred != 1) & (t != 1)).sum())
    return tp, fp, fn, tn


def auprc_from_counts(tp, fp, fn, tn):
    """sklearn.average_precision_score(target, hard_prediction) in closed form (utils.py:80-86): a hard 0/1 score has
    a two-threshold PR curve.  No positive target -> NaN -> 0, as the reference maps it."""
    n, npos = tp + fp + fn + tn, tp + fn
    if npos == 0:
        return 0.0
    if tp + fp == 0:
        return npos / n
    return (tp / npos) * (tp / (tp + fp)) + (fn / npos) * (npos / n)


def f1_precision_recall_from_counts(tp, fp, fn, tn):
    """precision_recall_fscore_support(average='macro', zero_division=0)[:3] (utils.py:89-94): macro over the labels
    that occur in target or prediction; returns (precision, recall, F1) in that order."""
    stats = []
    for t, f_p, f_n in ((tn, fn, fp), (tp, fp, fn)):
        if t + f_p + f_n == 0:
            continue
        prec = t / (t + f_p) if t + f_p else 0.0
        rec = t / (t + f_n) if t + f_n else 0.0
        stats.append((prec, rec, 2 * prec * rec / (prec + rec) if prec + rec else 0.0))
    return np.mean(np.array(stats), axis=0)


def AUPRC(output, target):
    return auprc_from_counts(*confusion_counts(output, target))


def F1_precision_recall(output, target):
    return f1_precision_recall_from_counts(*confusion_counts(output, target))


def get_loss_weights_from_labels(label):
    """utils.py:121-140: (w_pos, w_neg) by inverse number of samples."""
    label = torch.as_tensor(np.asarray(label.cpu() if torch.is_tensor(label) else label)).reshape(-1)
    pos, neg = int((label == 1).sum()), int((label == 0).sum())
    pos_inv = 1 / pos if pos != 0 else 0
    neg_inv = 1 / neg if neg != 0 else 0
    return pos_inv / (neg_inv + pos_inv), neg_inv / (neg_inv + pos_inv)


def get_loss_weights_from_dataloader(dataloader):
    pos = tot = 0
    for _, j in dataloader:
        pos += int(torch.as_tensor(j).sum())
        tot += len(j)
    neg = tot - pos
    pos_inv = 1 / pos if pos != 0 else 0
    neg_inv = 1 / neg if neg != 0 else 0
    return pos_inv / (neg_inv + pos_inv), neg_inv / (neg_inv + pos_inv)


def size_out_convolution(input_size, kernel, padding, stride):
    """utils.py:143-153."""
    return int(((input_size + 2 * padding - kernel) / stride) + 1)


def weight_reset(x):
    """utils.py:155-163 (Conv1d / Linear only: BatchNorm keeps its affine and running stats, quirk 4).  The mirrors'
    parameters are views into the engine arena, so the in-place re-initialisation lands there."""
    if isinstance(x, (nn.Conv1d, nn.Linear, nn.LSTM)):
        x.reset_parameters()


def get_input_size(data_loader):
    """utils.py:165-167.  Device loaders answer from their tensor shape; anything else is iterated like the reference does."""
    size = getattr(data_loader, 'input_size', None)
    if size is not None:
        return size
    for d, _ in data_loader:
        return d.shape[1]


def output_size_from_model_params(model_params):
    """utils.py:178-202."""
    input_size = 256
    for i in range(model_params['n_layers']):
        k = model_params[f'kernel_size_l{i}']
        output_size = size_out_convolution(input_size, k, int((k - 1) / 2), 1)
        output_size = size_out_convolution(output_size, 10, 0, 2)
        input_size = output_size
        out_channels = model_params[f'out_channels_l{i}']
    return output_size * out_channels


def get_single_model_params(model_params, models=['CNN', 'FFNN']):
    """utils.py:360-375: split the trial-parameter dict by its model prefix."""
    ddict = defaultdict(lambda: defaultdict(dict))
    if isinstance(models, str):
        models = [models]
    for model in models:
        for key in [k for k in model_params.keys() if k.startswith(model)]:
            start = re.search('_', key).span()[1]
            ddict[model][key[start:]] = model_params[key]
    return ddict
