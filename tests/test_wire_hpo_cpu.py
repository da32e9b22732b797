"""CPU tests of the callers / data format either side of the hot path: the batch wire format against
golden index batches of the reference's own BalancePos_BatchSampler, and the Optuna-compatible sweep substrate."""
import math
import os
import random

import numpy as np
import pytest
import torch

import embrace_b200  # noqa: F401
from embrace_b200.BIOINF_tesi.data_pipe import (encode_sequences, decode_onehot, BalancePos_BatchSampler, PackedDataset, build_loaders)
from embrace_b200.BIOINF_tesi.models.utils import hpo

GOLD = os.path.join(os.path.dirname(__file__), 'golden')


# ---- wire format ---------------------------------------------------------------------------------------
def test_balanced_batches_equal_the_reference_sampler():
    g = np.load(os.path.join(GOLD, 'balance_sampler.npz'))
    for case in range(3):
        y, bs = g[f'y{case}'], int(g[f'bs{case}'])
        s = BalancePos_BatchSampler(y, bs)
        assert len(s) == int(g[f'len{case}'])
        for epoch in range(2):
            batches = list(iter(s))
            assert [len(b) for b in batches] == g[f'n{case}_{epoch}'].tolist()
            assert len(batches) == len(s) + 1                                      # reference quirk 11
            assert np.concatenate(batches).tolist() == g[f'idx{case}_{epoch}'].tolist()


def test_encode_sequences_follows_process_sequence():
    rs = random.Random(5)
    seqs = [''.join(rs.choice('ACGTacgtnN') for _ in range(256)) for _ in range(20)]
    random.seed(99)
    codes = encode_sequences(seqs)
    # restatement of data_pipe/utils.py:268-276 with the same generator state
    random.seed(99)
    for s, row in zip(seqs, codes):
        bp = random.choice(['a', 'c', 'g', 't'])
        want = ['acgt'.index(bp if ch == 'n' else ch) for ch in s.lower()]
        assert row.tolist() == want
    onehot = torch.zeros(20, 4, 256)
    onehot[torch.arange(20)[:, None], torch.from_numpy(codes.astype(np.int64)), torch.arange(256)[None, :]] = 1
    assert torch.equal(decode_onehot(onehot), torch.from_numpy(codes))
    with pytest.raises(ValueError):
        encode_sequences(['acgt'])
    with pytest.raises(ValueError):
        encode_sequences(['x' * 256])


def test_loaders_stay_in_lock_step_and_cover_the_data():
    rs = np.random.RandomState(0)
    n = 530
    x, codes, y = rs.random_sample((n, 7)).astype(np.float32), rs.randint(0, 4, (n, 256)).astype(np.uint8), (rs.random_sample(n) < 0.2).astype(int)
    data = PackedDataset(x, codes, y, device='cpu')
    for training in (True, False):
        L = build_loaders(data, batch_size=100, training=training)
        for epoch in range(2):
            seen = []
            for (x1, t1), (x2, t2) in zip(L['FFNN'], L['CNN']):
                assert x1.shape[0] == x2.shape[0] == t1.shape[0] and t1.shape[1] == 1
                assert torch.equal(t1, t2)
                # rows can be identified by their (unique) feature vector
                for r in x1:
                    seen.append(int(np.nonzero((x == r.numpy()).all(1))[0][0]))
            assert sorted(seen) == list(range(n))
        assert len(L['FFNN']) == (6 if training else 3)
        if training:   # every batch gets its share of the positives
            pos = [int(t.sum()) for _, t in L['FFNN']]
            assert max(pos) - min(pos) <= 1


def test_loaders_stay_paired_after_a_solo_iteration():
    """Param_Search_Multimodal._build calls get_input_size(train_loader['FFNN']) once per trial
    (training_models_multimodal.py:313 in the reference), which iterates ONE loader on its own.  The two modality loaders
    must still hand out the same rows afterwards, in every later epoch, whichever of them is started first."""
    from embrace_b200.BIOINF_tesi.models.utils.utils import get_input_size
    from embrace_b200.BIOINF_tesi.models.utils.training_models_multimodal import paired_batches
    rs = np.random.RandomState(1)
    n = 1000
    x = rs.random_sample((n, 5)).astype(np.float32)
    codes = rs.randint(0, 4, (n, 256)).astype(np.uint8)
    x[:, 0] = np.arange(n)                                  # row identity travels in the features ...
    codes[:, 0] = np.arange(n) % 4                          # ... and in the sequences
    codes[:, 1] = (np.arange(n) // 4) % 4
    codes[:, 2] = (np.arange(n) // 16) % 4
    codes[:, 3] = (np.arange(n) // 64) % 4
    codes[:, 4] = (np.arange(n) // 256) % 4
    y = (rs.random_sample(n) < 0.15).astype(int)
    data = PackedDataset(x, codes, y, device='cpu')

    def row_of_codes(c):
        c = c.numpy().astype(np.int64)
        return c[:, 0] + 4 * c[:, 1] + 16 * c[:, 2] + 64 * c[:, 3] + 256 * c[:, 4]
    for training in (True, False):
        L = build_loaders(data, batch_size=100, training=training)
        assert get_input_size(L['FFNN']) == 5                # answered from the tensor shape: no iteration is started
        for _, _ in L['FFNN']:                               # a genuine solo iteration (what the reference's DataLoader would see)
            break
        for order in (('FFNN', 'CNN'), ('CNN', 'FFNN'), ('FFNN', 'CNN')):
            seen = 0
            for (a, ta), (b, tb) in zip(L[order[0]], L[order[1]]):
                xa, cb = (a, b) if order[0] == 'FFNN' else (b, a)
                assert torch.equal(ta, tb)
                assert np.array_equal(xa[:, 0].numpy().astype(np.int64), row_of_codes(cb)), 'features and sequences of different samples were paired'
                seen += len(xa)
            assert seen == n
        list(L['CNN'])                                       # a solo pass over the other modality
        for x1, x2, t in paired_batches(L):
            assert np.array_equal(x1[:, 0].numpy().astype(np.int64), row_of_codes(x2))


def test_paired_batches_raises_when_labels_disagree():
    from embrace_b200.BIOINF_tesi.models.utils.training_models_multimodal import paired_batches
    a = [(torch.zeros(3, 2), torch.tensor([[1], [0], [0]]))]
    b = [(torch.zeros(3, 4, 256), torch.tensor([[1], [1], [0]]))]
    with pytest.raises(AssertionError):
        list(paired_batches({'FFNN': a, 'CNN': b}))
    ok = list(paired_batches({'FFNN': a, 'CNN': [(b[0][0], a[0][1])]}))
    assert len(ok) == 1


# ---- sweep substrate -----------------------------------------------------------------------------------
def _space(trial):
    a = trial.suggest_int('n', 1, 4)
    b = trial.suggest_categorical('u', [32, 64, 128, 256])
    c = trial.suggest_loguniform('lr', 1e-5, 1e-1)
    d = trial.suggest_float('p', 0.0, 1.0)
    assert trial.suggest_int('n', 1, 4) == a          # repeated suggestions return the stored value
    return a, b, c, d


def test_samplers_respect_the_distributions_and_seeds():
    for make in (lambda s: hpo.RandomSampler(seed=s), lambda s: hpo.TPESampler(seed=s, n_startup_trials=4)):
        runs = []
        for rep in range(2):
            study = hpo.create_study(direction='maximize', sampler=make(7))
            vals = []
            study.optimize(lambda t: vals.append(_space(t)) or -abs(math.log10(vals[-1][2]) + 3), n_trials=12)
            runs.append(vals)
            for a, b, c, d in vals:
                assert 1 <= a <= 4 and b in (32, 64, 128, 256) and 1e-5 <= c <= 1e-1 and 0 <= d <= 1
        assert runs[0] == runs[1]
    # TPE concentrates on the good region after its start-up phase
    study = hpo.create_study(direction='maximize', sampler=hpo.TPESampler(seed=3, n_startup_trials=10))
    study.optimize(lambda t: -abs(math.log10(t.suggest_loguniform('lr', 1e-5, 1e-1)) + 3), n_trials=60)
    late = [abs(math.log10(t.params['lr']) + 3) for t in study.trials[30:]]
    early = [abs(math.log10(t.params['lr']) + 3) for t in study.trials[:10]]
    assert np.mean(late) < np.mean(early)
    assert abs(math.log10(study.best_params['lr']) + 3) < 0.3


def test_pruners():
    study = hpo.create_study(direction='maximize', pruner=hpo.MedianPruner(n_startup_trials=2))
    curves = {0: [0.5, 0.6, 0.7], 1: [0.4, 0.5, 0.6], 2: [0.1, 0.1, 0.1], 3: [0.9, 0.9, 0.9]}
    states = {}

    def objective(t):
        for step, v in enumerate(curves[t.number], 1):
            t.report(v, step)
            if t.should_prune():
                raise hpo.TrialPruned()
        return curves[t.number][-1]
    study.optimize(objective, n_trials=4)
    states = {t.number: t.state for t in study.trials}
    assert states == {0: 'COMPLETE', 1: 'COMPLETE', 2: 'PRUNED', 3: 'COMPLETE'}
    assert study.best_trial.number == 3
    # the reference's pruner: PatientPruner(MedianPruner(), patience=2) never fires within 3 trials (5 start-up trials)
    ref = hpo.create_study(direction='maximize', pruner=hpo.PatientPruner(hpo.MedianPruner(), patience=2))
    ref.optimize(lambda t: [t.report(0.1, s) or t.should_prune() for s in range(1, 8)] and 0.1, n_trials=3)
    assert all(t.state == 'COMPLETE' for t in ref.trials)
    # patience: no deferral while the curve still improves
    pp = hpo.PatientPruner(type('Always', (), {'prune': lambda self, s, t: True})(), patience=2)
    st = hpo.create_study(direction='maximize', pruner=pp)
    fired = []

    def obj2(t):
        for step, v in enumerate([0.1, 0.2, 0.3, 0.3, 0.3, 0.3], 1):
            t.report(v, step)
            fired.append(t.should_prune())
        return 0.3
    st.optimize(obj2, n_trials=1)
    assert fired == [False, False, False, False, False, True]


def test_jsonl_storage_resumes_like_load_if_exists(tmp_path):
    path = str(tmp_path / 'studies.db')
    st = hpo.create_study(study_name='A549_x_1', direction='maximize', storage=f'sqlite:///{path}', load_if_exists=True,
                          sampler=hpo.RandomSampler(seed=1))
    st.optimize(lambda t: t.suggest_float('p', 0, 1), n_trials=2)
    assert os.path.exists(str(tmp_path / 'studies.jsonl'))
    again = hpo.create_study(study_name='A549_x_1', direction='maximize', storage=f'sqlite:///{path}', load_if_exists=True)
    assert len([t for t in again.trials if t.state == 'COMPLETE']) == 2
    again.optimize(lambda t: t.suggest_float('p', 0, 1), n_trials=1)
    assert [t.number for t in again.trials] == [0, 1, 2]
    assert again.best_value == max(t.value for t in again.trials)
    with pytest.raises(ValueError):
        hpo.create_study(study_name='A549_x_1', storage=f'sqlite:///{path}')
    other = hpo.create_study(study_name='A549_x_2', storage=f'sqlite:///{path}', load_if_exists=True)
    assert other.trials == []
    with pytest.raises(RuntimeError):
        other.optimize(lambda t: (_ for _ in ()).throw(RuntimeError('boom')), n_trials=1)
    assert other.trials[0].state == 'FAIL'


def test_sweep_job_list_and_inference_sharding():
    from embrace_b200.sweep import make_jobs, synthetic_dataset, CELL_F
    jobs = make_jobs(14, 3)
    assert len(jobs) == 42 and len({(j['cell_line'], j['task']) for j in jobs}) == 14
    x, codes, y = synthetic_dataset('HEPG2', 'active_P_vs_inactive_P', 500, 1)
    assert x.shape == (500, CELL_F['HEPG2']) and codes.shape == (500, 256) and codes.max() <= 3 and 0 < y.mean() < 0.5
    from embrace_b200.infer import synthetic_availabilities
    av = synthetic_availabilities(10000, 0)
    assert abs((av.sum(1) == 2).mean() - 0.8) < 0.02 and (av.sum(1) >= 1).all()


def test_sweep_driver_hands_every_worker_jobs(tmp_path):
    """The fold-parallel driver's queue plumbing without a GPU (`--dry-run`: a job is a sleep).  Round 2 found six of eight
    workers leaving at once because `Queue.get_nowait()` raced the queue's feeder thread; jobs are now pulled with a blocking
    get and one sentinel per worker ends them."""
    from embrace_b200 import sweep
    line = sweep.main(['--gpus', '3', '--datasets', '4', '--folds', '3', '--trials', '2', '--dry-run', '0.05', '--out', str(tmp_path / 'sw')])
    assert line['jobs'] == 12 and line['trials'] == 24
    assert sorted(line['gpu_busy_s']) == ['0', '1', '2'], line['gpu_busy_s']
    assert line['balance'] > 0.6
    jobs = sweep.make_jobs(14, 3)
    assert len(jobs) == 42 and sweep.CELL_F[jobs[0]['cell_line']] == max(sweep.CELL_F.values())      # longest first
