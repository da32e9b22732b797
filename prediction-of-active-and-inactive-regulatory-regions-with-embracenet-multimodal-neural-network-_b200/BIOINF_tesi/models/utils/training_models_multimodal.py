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
        train_loss, AUPRC_train, _ = _epoch_metrics(model.metrics_read() if started else [], n_train)
        model.eval()
        started = False
        for bi, (x1, x2, target) in enumerate(batches_test()):
            d = draws_test(epoch - 1, bi) if draws_test else None
            model.eval_batch(x1, x2, target, draws=d, reset_metrics=not started)
            started = True
        test_loss, AUPRC_test, F1_test = _epoch_metrics(model.metrics_read() if started else [], n_test)
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


def paired_batches(loader):
    """zip(loader['FFNN'], loader['CNN']) with the reference's two asserts (training_models_multimodal.py:136-137): equal
    batch lengths and `torch.eq(target, _).all()`.  For device-resident targets the label comparison is accumulated in a
    device flag and read ONCE when the epoch's iteration ends (no per-batch sync); host targets are compared at once."""
    flag = None
    for (x_1, target), (x_2, t2) in zip(loader['FFNN'], loader['CNN']):
        assert len(x_1) == len(x_2)
        if torch.is_tensor(target) and target.is_cuda:
            bad = (target.reshape(-1) != torch.as_tensor(t2).reshape(-1).to(target.device)).any()
            flag = bad if flag is None else (flag | bad)
        else:
            assert torch.equal(torch.as_tensor(target).reshape(-1), torch.as_tensor(t2).reshape(-1).cpu()), \
                'FFNN and CNN loaders disagree on the labels of a batch'
        yield x_1, x_2, target
    if flag is not None and bool(flag):
        raise AssertionError('FFNN and CNN loaders disagree on the labels of a batch (loaders out of step?)')


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
        return lambda: paired_batches(loader)
    scores = run_epochs(model, pairs(train_loader), pairs(test_loader), len(train_loader['FFNN']), len(test_loader['FFNN']), cfg,
                        num_epochs, patience, delta, verbose, draws_train, draws_test)
    if checkpoint_path:
        torch.save({'model_state_dict': model.state_dict(), 'AUPRC_train_scores': scores[0], 'AUPRC_test_scores': scores[1],
                    'F1_precision_recall_test_scores': scores[2]}, checkpoint_path)
    return scores


# =====================================================================================================
# Hyper-parameter search and K-fold cross-validation: callers of the hot path
# =====================================================================================================
from collections import defaultdict

from . import hpo
from . import optim as _optim
from .utils import get_input_size, weight_reset


def dd():
    return defaultdict(list)


def path_augmentation(augmentation):
    return '_augmentation' if augmentation else ''


def _make_optimizer(name, params, lr, weight_decay):
    """Nadam (timm in the reference) / Adam / RMSprop with library defaults but lr and weight_decay (:318-325, :585-591)."""
    if name not in ('Nadam', 'Adam', 'RMSprop'):
        raise ValueError(f'unknown optimizer {name!r}')
    return getattr(_optim, name)(params, lr=lr, weight_decay=weight_decay)


class _FixedParams:
    """Trial stand-in that replays a finished trial's parameters (rebuilding the best model from study.best_params)."""

    def __init__(self, params, number=0):
        self.params, self.number = dict(params), number

    def _get(self, name, *a, **k):
        return self.params[name]
    suggest_int = suggest_categorical = suggest_float = suggest_loguniform = suggest_uniform = _get

    def report(self, value, step):
        pass

    def should_prune(self):
        return False


class Param_Search_Multimodal():
    """Hyper-parameter tuning of one fold (training_models_multimodal.py:232-462): `n_trials` trials of a study, each
    building `model(trial, ...)` from the trial's suggestions, training it with per-epoch pruning reports and early
    stopping (patience 4) on the test AUPRC.  Same constructor, attributes (`best_model`, `best_params`) and study /
    checkpoint naming; the Optuna pieces come from `hpo` and the loop body is the fused engine step.
    Deviation: trials save {'model_state_dict', 'model_params'} instead of pickling the whole module (:413)."""

    def __init__(self, model, train_loader, test_loader, num_epochs, study_name, device, cell_line, task, sampler='TPE',
                 n_trials=3, storage='BIOINF_optuna_tuning.db', seed=None):
        self.model_ = model                        # the CLASS, as in the notebooks (copy.deepcopy of a class is the class)
        self.train_loader, self.test_loader = train_loader, test_loader
        self.num_epochs, self.study_name, self.device = num_epochs, study_name, device
        self.cell_line, self.task, self.n_trials, self.storage = cell_line, task, n_trials, storage
        if cell_line not in CELL_LINES:
            raise ValueError(f"Argument 'cell_line' has an incorrect value: use one among {CELL_LINES}")
        if task not in TASKS:
            raise ValueError(f"Argument 'task' has an incorrect value: use one among {TASKS} ")
        self.model_name = model.__name__
        if sampler == 'BO':
            raise ValueError("sampler 'BO' (BoTorch) is not available: use 'TPE' or 'random'")
        elif sampler == 'TPE':
            self.sampler = hpo.TPESampler(seed=seed)
        elif sampler == 'random':
            self.sampler = hpo.RandomSampler(seed=seed)
        else:
            raise ValueError(f'unknown sampler {sampler!r}')

    def _build(self, trial):
        in_features_FFNN = get_input_size(self.train_loader['FFNN'])
        return self.model_(trial, cell_line=self.cell_line, task=self.task, device=self.device, in_features_FFNN=in_features_FFNN)

    def objective(self, trial):
        """AUPRC test score of one trial (:307-416)."""
        self.model = self._build(trial)
        optimizer_name = trial.suggest_categorical("optimizer", ["Nadam", "Adam", "RMSprop"])
        lr = trial.suggest_loguniform("lr", 1e-5, 1e-1)
        weight_decay = trial.suggest_loguniform("weight_decay", 1e-4, 1e-1)
        cfg = lift_optimizer(_make_optimizer(optimizer_name, self.model.parameters(), lr, weight_decay))
        self.model = self.model.double().to(self.device)
        early_stopping = EarlyStopping(patience=4, verbose=True)
        n_test = len(self.test_loader['FFNN'])
        AUPRC_test = 0.0
        for epoch in range(1, self.num_epochs + 1):
            self.model.train()
            started = False
            for x_1, x_2, target in paired_batches(self.train_loader):
                self.model.train_batch(x_1, x_2, target, cfg, reset_metrics=not started)
                started = True
            self.model.eval()
            started = False
            for x_1, x_2, target in paired_batches(self.test_loader):
                self.model.eval_batch(x_1, x_2, target, reset_metrics=not started)
                started = True
            _, AUPRC_test, _ = _epoch_metrics(self.model.metrics_read() if started else [], n_test)
            trial.report(AUPRC_test, epoch)
            if trial.should_prune():
                raise hpo.TrialPruned()
            early_stopping(AUPRC_test)
            if early_stopping.early_stop:
                print('Early stopping the training')
                break
        torch.save({'model_state_dict': self.model.state_dict(), 'model_params': dict(trial.params)}, f'{self.study_name}{trial.number}.pt')
        return AUPRC_test

    def run_trial(self):
        """Runs (or resumes) the study and stores the best model / parameters (:420-460)."""
        study = hpo.create_study(study_name=self.study_name, direction="maximize",
                                 pruner=hpo.PatientPruner(hpo.MedianPruner(), patience=2),
                                 storage=f'sqlite:///{self.storage}' if self.storage else None, load_if_exists=True, sampler=self.sampler)
        complete_trials = [t for t in study.trials if t.state == hpo.TrialState.COMPLETE]
        if len(complete_trials) < self.n_trials:
            study.optimize(self.objective, n_trials=self.n_trials - len(complete_trials))
        trials = study.trials
        pruned_trials = [t for t in trials if t.state == hpo.TrialState.PRUNED]
        complete_trials = [t for t in trials if t.state == hpo.TrialState.COMPLETE]
        best = study.best_trial
        ck = torch.load(f'{self.study_name}{best.number}.pt', map_location='cpu', weights_only=False)
        best_model = self._build(_FixedParams(best.params, best.number))
        best_model.load_state_dict(ck['model_state_dict'])
        self.best_model = best_model
        self.study = study
        print("Study statistics: ")
        print("  Number of finished trials: ", len(trials))
        print("  Number of pruned trials: ", len(pruned_trials))
        print("  Number of complete trials: ", len(complete_trials))
        print("Best trial:")
        self.best_params = dict(best.params)
        print("  Value: ", best.value)
        print("  Params: ")
        for key, value in best.params.items():
            print("    {}: {}".format(key, value))


class ArrayDataClass:
    """Minimal stand-in for Data_Prepare's CV interface (dataprepare.py:264-306) over in-memory arrays:
    return_index_data_for_cv -> (KFold(n_folds, shuffle=True, random_state), X, y)."""

    def __init__(self, features, sequences, labels):
        self.features, self.sequences, self.labels = np.asarray(features), sequences, np.asarray(labels).reshape(-1)

    def return_index_data_for_cv(self, cell_line=None, sequence=False, n_folds=3, random_state=789):
        from sklearn.model_selection import KFold
        return KFold(n_splits=n_folds, shuffle=True, random_state=random_state), (self.sequences if sequence else self.features), self.labels


class ArrayPipeline:
    """`build_dataloader_pipeline` stand-in: only `.data_class` is used by Kfold_CV_Multimodal.__call__ (:719-722)."""

    def __init__(self, features, sequences, labels):
        self.data_class = ArrayDataClass(features, sequences, labels)


def _take(X, idx):
    return X.iloc[idx] if hasattr(X, 'iloc') else (X[idx] if isinstance(X, np.ndarray) else [X[i] for i in idx])


def _to_codes(seqs):
    from ...data_pipe import encode_sequences
    if isinstance(seqs, np.ndarray) and seqs.dtype == np.uint8:
        return seqs
    if hasattr(seqs, 'tolist'):
        seqs = seqs.tolist()
    return encode_sequences(list(seqs))


class Kfold_CV_Multimodal():
    """K-fold cross-validation (training_models_multimodal.py:475-798): per fold a hyper-parameter study on a
    train/validation split, then a final fit of the best architecture (weights reset, BatchNorm kept: utils.weight_reset)
    on train+validation, scored on the held-out fold.  Same call signature, scores_dict layout, study names
    (`{study_name}_{fold}`) and checkpoint names.  Data enters through `build_dataloader_pipeline.data_class
    .return_index_data_for_cv` (pandas objects or arrays); the packed device-resident loaders of data_pipe.wire replace
    Dataset_Wrap / DataLoader.  Augmentation and rebalancing (imblearn) belong to the out-of-scope preprocessing."""

    def __init__(self):
        self.scores_dict = defaultdict(dd)
        self.scores_dict['final_test_AUPRC_scores'] = []
        self.scores_dict['final_train_AUPRC_scores'] = []
        self.model_ = []
        self.optimizer = []
        self.best_params = defaultdict(dict)

    def build_dataloader_forCV(self, X, y, sequences, batch_size=100, training=True, augmentation=False):
        """X: features, sequences: strings or uint8 codes (lists of parts are concatenated, :513-522) -> {'FFNN','CNN'} loaders."""
        from ...data_pipe import PackedDataset, build_loaders
        if augmentation:
            raise NotImplementedError('data augmentation / rebalancing is part of the out-of-scope preprocessing (data_pipe/utils.py:327-685)')
        if isinstance(X, list):
            X = np.concatenate([np.asarray(x) for x in X])
            y = np.concatenate([np.asarray(v).reshape(-1) for v in y])
            sequences = np.concatenate([_to_codes(s) for s in sequences])
        data = PackedDataset(np.asarray(X), _to_codes(sequences), y, device=self.device)
        return build_loaders(data, batch_size=batch_size, training=training, random_state=self.random_state)

    def hyper_tuning(self, train_loader, test_loader, num_epochs, cell_line, task, study_name, device, sampler):
        param_search = Param_Search_Multimodal(model=self.model_, train_loader=train_loader, test_loader=test_loader,
                                               num_epochs=num_epochs, cell_line=cell_line, task=task, device=device, sampler=sampler,
                                               n_trials=self.n_trials, study_name=study_name, storage=self.storage, seed=self.sampler_seed)
        param_search.run_trial()
        best_params = param_search.best_params
        self.model_ = param_search.best_model
        self.best_params[self.i] = best_params
        self.model_.apply(weight_reset)
        self.optimizer = _make_optimizer(best_params['optimizer'], self.model_.parameters(), best_params['lr'], best_params['weight_decay'])

    def model_testing(self, train_loader, test_loader, num_epochs, test_model_path, device, cell_line, task, checkpoint_path=None):
        AUPRC_train, AUPRC_test, other_scores = fit_multimodal(model=self.model_, train_loader=train_loader, test_loader=test_loader,
                                                               device=device, cell_line=cell_line, task=task, optimizer=self.optimizer,
                                                               num_epochs=num_epochs, patience=4, verbose=False,
                                                               checkpoint_path=f'{checkpoint_path}.pt')
        self.scores_dict[f'iteration_n_{self.i}']['AUPRC_train'] = AUPRC_train
        self.scores_dict[f'iteration_n_{self.i}']['AUPRC_test'] = AUPRC_test
        self.scores_dict[f'iteration_n_{self.i}']['F1_precision_recall'] = other_scores
        final_test_AUPRC_score = AUPRC_test[-1]
        self.scores_dict['final_test_AUPRC_scores'].append(final_test_AUPRC_score)
        self.scores_dict['final_train_AUPRC_scores'].append(AUPRC_train[-1])
        print(f'AUPRC test score: {final_test_AUPRC_score}\n\n')
        self.avg_score.append(final_test_AUPRC_score)
        if final_test_AUPRC_score == max(self.avg_score) and test_model_path:
            os.makedirs('models_', exist_ok=True)
            torch.save({'model_state_dict': self.model_.state_dict(), 'model_params': self.best_params[self.i]}, f'models_/{test_model_path}.pt')

    def run_fold(self, i, train_index, test_index, X_1, X_2, y, model, cell_line, task, num_epochs, batch_size, study_name, sampler,
                 test_model_path):
        """One CV iteration (:725-792); the unit of work of the trial-parallel sweep (one fold per GPU job)."""
        from sklearn.model_selection import train_test_split
        self.i = i
        STUDY_NAME = f'{study_name}_{str(self.i)}'
        print(f'>>> ITERATION N. {self.i}')
        X_train_1, X_test_1 = _take(X_1, train_index), _take(X_1, test_index)
        X_train_2, X_test_2 = _take(X_2, train_index), _take(X_2, test_index)
        y_train, y_test = _take(y, train_index), _take(y, test_index)
        X_train_1, X_val_1, _, _ = train_test_split(X_train_1, y_train, test_size=1 / self.n_folds, random_state=self.random_state, shuffle=True)
        X_train_2, X_val_2, y_train, y_val = train_test_split(X_train_2, y_train, test_size=1 / self.n_folds, random_state=self.random_state,
                                                              shuffle=True)
        self.model_ = model
        print('\n===============> HYPERPARAMETERS TUNING')
        train_loader = self.build_dataloader_forCV(X_train_1, y_train, X_train_2, batch_size=batch_size, training=True, augmentation=self.augmentation)
        test_loader = self.build_dataloader_forCV(X_val_1, y_val, X_val_2, batch_size=batch_size, training=False)
        self.hyper_tuning(train_loader, test_loader, num_epochs, cell_line, task, STUDY_NAME, self.device, sampler)
        print('\n===============> MODEL TESTING')
        train_loader = self.build_dataloader_forCV([X_train_1, X_val_1], [y_train, y_val], [X_train_2, X_val_2], batch_size=batch_size,
                                                   training=True, augmentation=self.augmentation)
        test_loader = self.build_dataloader_forCV(X_test_1, y_test, X_test_2, batch_size=batch_size, training=False)
        self.model_testing(train_loader, test_loader, num_epochs, test_model_path, self.device, cell_line, task,
                           checkpoint_path=f'{cell_line}_{model.__name__}{path_augmentation(self.augmentation)}_{task}_{self.i}_test_')

    def __call__(self, build_dataloader_pipeline, cell_line, device, task=None, model=None, augmentation=False, rebalance_threshold=0.1,
                 random_state=789, n_folds=3, num_epochs=100, batch_size=100, study_name=None, sampler='TPE', test_model_path=None,
                 n_trials=3, storage='BIOINF_optuna_tuning.db', sampler_seed=None, folds=None):
        """Reference signature (:645-660) plus: n_trials (hard-coded 3 there, :573), storage, sampler_seed, and `folds`
        (iterable of 1-based fold numbers to run: the sweep driver gives every GPU worker its own folds)."""
        self.n_folds, self.augmentation, self.rebalance_threshold = n_folds, augmentation, rebalance_threshold
        self.random_state, self.device = random_state, device
        self.n_trials, self.storage, self.sampler_seed = n_trials, storage, sampler_seed
        self.avg_score = []
        self.hp_score = []
        if cell_line not in CELL_LINES:
            raise ValueError(f"Argument 'cell_line' has an incorrect value: use one among {CELL_LINES}")
        if task not in TASKS:
            raise ValueError(f"Argument 'task' has an incorrect value: use one among {TASKS} ")
        data_class = build_dataloader_pipeline.data_class
        kf, X_1, y = data_class.return_index_data_for_cv(cell_line=cell_line, sequence=False, n_folds=n_folds, random_state=self.random_state)
        _, X_2, _ = data_class.return_index_data_for_cv(cell_line=cell_line, sequence=True, n_folds=n_folds, random_state=self.random_state)
        for i, (train_index, test_index) in enumerate(kf.split(X_1)):
            if folds is not None and (i + 1) not in folds:
                continue
            self.run_fold(i + 1, train_index, test_index, X_1, X_2, y, model, cell_line, task, num_epochs, batch_size, study_name, sampler,
                          test_model_path)
        avg_CV_AUPRC = np.round(sum(self.avg_score) / max(len(self.avg_score), 1), 5)
        self.scores_dict['average_CV_AUPRC'] = avg_CV_AUPRC
        print(f'\n{n_folds}-FOLD CROSS-VALIDATION AUPRC TEST SCORE: {avg_CV_AUPRC}')
        return self.scores_dict
