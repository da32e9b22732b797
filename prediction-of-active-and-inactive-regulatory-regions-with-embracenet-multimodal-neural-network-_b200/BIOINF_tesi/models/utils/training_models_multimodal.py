"""Mirror of fit_multimodal (BIOINF_tesi/models/utils/training_models_multimodal.py:40-226).

Same signature, same loader convention, same return lists, same checkpoint keys.  The loop body (forward, weighted CE,
backward, optimizer step, per-batch metric) is ONE fused engine call per batch; loss and confusion counts accumulate on
the device and are read back once per epoch instead of twice per batch."""
import os

import numpy as np
import torch

from .utils import EarlyStopping, auprc_from_counts, f1_precision_recall_from_counts
from .... import _native as N

CELL_LINES = ['A549', 'GM12878', 'H1', 'HEK293', 'HEPG2', 'K562', 'MCF7']
TASKS = ['active_E_vs_inactive_E', 'active_P_vs_inactive_P', 'active_E_vs_active_P', 'inactive_E_vs_inactive_P',
         'active_EP_vs_inactive_rest']


def lift_optimizer(optimizer):
    """torch.optim.Adam / AdamW / RMSprop / NAdam (or timm's Nadam) instance -> EmbOptConfig for the fused kernel.
    The reference builds these with lr / weight_decay from the trial and library defaults otherwise (:318-325)."""
    from ....engine import Engine
    if optimizer is None:
        raise ValueError('fit needs an optimizer (the reference calls optimizer.zero_grad() unconditionally)')
    if len(optimizer.param_groups) != 1:
        raise ValueError('one parameter group expected (the reference passes model.parameters())')
    g = optimizer.param_groups[0]
    name = type(optimizer).__name__.lower()
    if name in ('adam', 'adamw'):
        if g.get('amsgrad', False) or g.get('maximize', False):
            raise ValueError('amsgrad / maximize are not implemented by the fused optimizer')
        return Engine.opt_config(name, lr=g['lr'], weight_decay=g['weight_decay'], betas=tuple(g['betas']), eps=g['eps'])
    if name == 'rmsprop':
        if g.get('momentum', 0) != 0 or g.get('centered', False):
            raise ValueError('RMSprop momentum / centered are not implemented by the fused optimizer')
        return Engine.opt_config('rmsprop', lr=g['lr'], weight_decay=g['weight_decay'], alpha=g['alpha'], eps=g['eps'])
    if name == 'nadam':
        md = g.get('momentum_decay', g.get('schedule_decay', 4e-3))
        return Engine.opt_config('nadam', lr=g['lr'], weight_decay=g['weight_decay'], betas=tuple(g['betas']), eps=g['eps'], momentum_decay=md)
    raise ValueError(f'optimizer {type(optimizer).__name__} is not supported (Adam, AdamW, RMSprop, Nadam)')


def _epoch_metrics(records, n_batches):
    """Per-batch records -> (loss sum, mean AUPRC over len(loader), mean F1/precision/recall) as the reference does."""
    loss = sum(r['loss'] for r in records)
    auprc = sum(auprc_from_counts(r['tp'], r['fp'], r['fn'], r['tn']) for r in records) / n_batches
    f1 = np.zeros(3)
    for r in records:
        f1 += f1_precision_recall_from_counts(r['tp'], r['fp'], r['fn'], r['tn'])
    return loss, auprc, f1 / n_batches


def run_epochs(model, batches_train, batches_test, n_train, n_test, cfg, num_epochs, patience, delta, verbose, draws_train=None,
               draws_test=None):
    """Shared by fit() and fit_multimodal().  batches_*: callables returning an iterator of (x_ffnn|None, x_cnn|None, target)."""
    AUPRC_train_scores, AUPRC_test_scores, F1_scores = [], [], []
    early_stopping = EarlyStopping(patience=patience, delta=delta, verbose=True)
    for epoch in range(1, num_epochs + 1):
        model.train()
        started = False
        for bi, (x1, x2, target) in enumerate(batches_train()):
            d = draws_train(epoch - 1, bi) if draws_train else None
            model.train_batch(x1, x2, target, cfg, draws=d, reset_metrics=not started)
            started = True
        train_loss, AUPRC_train, _ = _epoch_metrics(model.engine.metrics_read() if started else [], n_train)
        model.eval()
        started = False
        for bi, (x1, x2, target) in enumerate(batches_test()):
            d = draws_test(epoch - 1, bi) if draws_test else None
            model.eval_batch(x1, x2, target, draws=d, reset_metrics=not started)
            started = True
        test_loss, AUPRC_test, F1_test = _epoch_metrics(model.engine.metrics_read() if started else [], n_test)
        AUPRC_train_scores.append(AUPRC_train)
        AUPRC_test_scores.append(AUPRC_test)
        F1_scores.append(F1_test)
        if verbose:
            print('Epoch: {} \tTraining AUPRC score: {:.4f} \tTest AUPRC score: {:.4f} \tTraining Loss: {:.4f} \tTest Loss: {:.4f}'.format(
                epoch, AUPRC_train, AUPRC_test, train_loss, test_loss))
        early_stopping(AUPRC_test)
        if early_stopping.early_stop:
            print('Early stopping the training')
            break
    return AUPRC_train_scores, AUPRC_test_scores, F1_scores


def fit_multimodal(model, train_loader, test_loader, device, cell_line, task, optimizer=None, num_epochs=100, patience=4, delta=0,
                   verbose=True, checkpoint_path=None, draws_train=None, draws_test=None):
    """Lists of AUPRC_train_scores, AUPRC_test_scores, F1_precision_recall_test_scores per epoch.
    train_loader / test_loader: {'FFNN': iterable of (x[B,F], y[B,1]), 'CNN': iterable of (x[B,4,256] one-hot | [B,256] codes, y)},
    both supporting len().  draws_*: optional (epoch, batch) -> replayed draws dict (parity tests)."""
    if cell_line not in CELL_LINES:
        raise ValueError(f"Argument 'cell_line' has an incorrect value: use one among {CELL_LINES}")
    if task not in TASKS:
        raise ValueError(f"Argument 'task' has an incorrect value: use one among {TASKS} ")
    if os.path.exists(checkpoint_path):          # checkpoint_path=None raises here, exactly like the reference
        checkpoint = torch.load(checkpoint_path, map_location='cpu', weights_only=False)
        model.load_state_dict(checkpoint['model_state_dict'])
        return checkpoint['AUPRC_train_scores'], checkpoint['AUPRC_test_scores'], checkpoint['F1_precision_recall_test_scores']
    cfg = lift_optimizer(optimizer)
    model = model.double().to(device)            # accepted and ignored: fp32 master weights on the GPU

    def pairs(loader):
        def it():
            for (x_1, target), (x_2, t2) in zip(loader['FFNN'], loader['CNN']):
                assert len(x_1) == len(x_2)
                if not (torch.is_tensor(target) and target.is_cuda):      # the reference's assert, without forcing a device sync
                    assert torch.equal(torch.as_tensor(target).reshape(-1), torch.as_tensor(t2).reshape(-1).cpu())
                yield x_1, x_2, target
        return it
    scores = run_epochs(model, pairs(train_loader), pairs(test_loader), len(train_loader['FFNN']), len(test_loader['FFNN']), cfg,
                        num_epochs, patience, delta, verbose, draws_train, draws_test)
    if checkpoint_path:
        torch.save({'model_state_dict': model.state_dict(), 'AUPRC_train_scores': scores[0], 'AUPRC_test_scores': scores[1],
                    'F1_precision_recall_test_scores': scores[2]}, checkpoint_path)
    return scores
