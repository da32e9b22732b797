"""GPU tests of the callers of the hot path: hyper-parameter study, K-fold CV, the sweep driver (1 GPU) and sharded
inference with missing-modality masking (BASELINE configs 4 and 5 as parity / behaviour cases)."""
import json
import os

import numpy as np
import pytest

from oracle import embracenet_oracle as O
from tests.golden.cases import ARCH_S, make_inputs
from tests.test_gpu_parity import to_archspec

pytestmark = pytest.mark.gpu


def _data(n, F, seed):
    from embrace_b200.sweep import synthetic_dataset
    x, codes, y = synthetic_dataset('A549', 'active_E_vs_inactive_E', n, seed)
    return x[:, :F], codes, y


def test_param_search_runs_resumes_and_rebuilds_the_best_model(tmp_path, monkeypatch):
    import torch
    from embrace_b200.BIOINF_tesi.models import EmbraceNetMultimodal
    from embrace_b200.BIOINF_tesi.models.utils import Param_Search_Multimodal
    from embrace_b200.BIOINF_tesi.data_pipe import PackedDataset, build_loaders
    monkeypatch.chdir(tmp_path)
    x, codes, y = _data(700, 48, 3)
    tr = build_loaders(PackedDataset(x[:500], codes[:500], y[:500]), batch_size=100, training=True)
    te = build_loaders(PackedDataset(x[500:], codes[500:], y[500:]), batch_size=100, training=False)
    ps = Param_Search_Multimodal(EmbraceNetMultimodal, tr, te, num_epochs=2, study_name='A549_test_1', device='cuda', cell_line='A549',
                                 task='active_E_vs_inactive_E', sampler='TPE', n_trials=2, storage='studies.db', seed=5)
    ps.run_trial()
    assert set(['optimizer', 'lr', 'weight_decay', 'FFNN_n_layers', 'CNN_n_layers', 'EMBRACENET_embracement_size']) <= set(ps.best_params)
    assert os.path.exists('A549_test_10.pt') and os.path.exists('A549_test_11.pt') and os.path.exists('studies.jsonl')
    ck = torch.load(f'A549_test_1{ps.study.best_trial.number}.pt', weights_only=False)
    assert ck['model_params'] == ps.best_params
    sd = ps.best_model.state_dict()
    assert all(torch.equal(sd[k].cpu(), ck['model_state_dict'][k].cpu()) for k in sd)
    n_before = len(ps.study.trials)
    ps2 = Param_Search_Multimodal(EmbraceNetMultimodal, tr, te, num_epochs=2, study_name='A549_test_1', device='cuda', cell_line='A549',
                                  task='active_E_vs_inactive_E', n_trials=2, storage='studies.db')
    ps2.run_trial()                                        # load_if_exists: nothing left to run
    assert len(ps2.study.trials) == n_before and ps2.best_params == ps.best_params
    with pytest.raises(ValueError):
        Param_Search_Multimodal(EmbraceNetMultimodal, tr, te, 1, 's', 'cuda', 'HeLa', 'active_E_vs_inactive_E')


def test_kfold_cv_one_fold_and_the_sweep_driver(tmp_path, monkeypatch):
    from embrace_b200.BIOINF_tesi.models import EmbraceNetMultimodal
    from embrace_b200.BIOINF_tesi.models.utils.training_models_multimodal import Kfold_CV_Multimodal, ArrayPipeline
    monkeypatch.chdir(tmp_path)
    x, codes, y = _data(900, 48, 4)
    seqs = [''.join('acgt'[c] for c in row) for row in codes]            # strings, as the reference's data frames hold them
    cv = Kfold_CV_Multimodal()
    scores = cv(ArrayPipeline(x, seqs, y), 'A549', 'cuda', task='active_E_vs_inactive_E', model=EmbraceNetMultimodal, n_folds=3,
                num_epochs=2, batch_size=100, study_name='A549_EmbraceNetMultimodal', n_trials=2, storage='s.db', sampler_seed=1,
                folds=[2], test_model_path='best_A549')
    assert len(scores['final_test_AUPRC_scores']) == 1 and 0.0 <= scores['final_test_AUPRC_scores'][0] <= 1.0
    assert len(scores['iteration_n_2']['AUPRC_test']) >= 1
    assert os.path.exists('A549_EmbraceNetMultimodal_active_E_vs_inactive_E_2_test_.pt') and os.path.exists('models_/best_A549.pt')
    assert set(cv.best_params[2]) >= {'optimizer', 'lr', 'weight_decay'}
    from embrace_b200 import sweep
    line = sweep.main(['--gpus', '1', '--datasets', '1', '--folds', '2', '--trials', '2', '--rows', '600', '--epochs', '1', '--out', 'sw'])
    assert line['jobs'] == 2 and line['trials'] == 4 and line['value'] > 0
    assert len([l for l in open('sw/sweep_studies.jsonl') if json.loads(l)['state'] == 'COMPLETE']) == 4


@pytest.mark.parametrize('precision', ['fp32', 'bf16'])
def test_sharded_inference_equals_unsharded_and_masks_modalities(precision):
    import torch
    from embrace_b200 import Engine
    from embrace_b200.infer import score_regions, score_regions_sharded, synthetic_availabilities
    spec, n = ARCH_S, 300
    P = O.init_params(spec, 9)
    x, codes, _ = make_inputs(spec, n, 10)
    av = synthetic_availabilities(n, 1)
    av[:3] = [[1, 0], [0, 1], [1, 1]]
    xd, cd, avd = torch.from_numpy(x.astype(np.float32)).cuda(), torch.from_numpy(codes).cuda(), torch.from_numpy(av).cuda()

    def engine():
        e = Engine(to_archspec(spec), max_batch=64, precision=precision, seed=11)
        e.load_numpy(P)
        return e
    full = score_regions(engine(), xd, cd, avd, batch=64).cpu().numpy()
    assert full.shape == (n,) and np.isfinite(full).all() and (full >= 0).all() and (full <= 1).all()
    # single-modality rows are deterministic (every dimension comes from the available modality), whatever the draws and sharding
    one = av.sum(1) == 1
    parts = np.empty(n, dtype=np.float32)
    for r in range(3):
        lo, hi, s = score_regions_sharded(engine(), xd, cd, avd, rank=r, world=3, batch=64)
        parts[lo:hi] = s.cpu().numpy()
    tol = 1e-6 if precision == 'fp32' else 2e-2
    assert np.abs(parts[one] - full[one]).max() <= tol
    # reference semantics on those rows: the oracle with an all-one-modality selection
    u = np.full((n, spec['C']), 0.5)      # cum0 is 1 (FFNN only) or 0 (CNN only) on those rows: any u in (0, 1) selects the available one
    ref = O.predict_proba(spec, P, x, codes, u, availabilities=av)
    assert np.abs(full[one] - ref[one]).max() <= (2e-6 if precision == 'fp32' else 2e-2)


def test_concatnet_mirror_matches_the_oracle_and_runs_a_study(tmp_path, monkeypatch):
    """ConcatNetMultimodal: reference constructor order and state_dict keys, eval forward vs the oracle
    (itself pinned to the reference's golden vectors), and the class plugs into Param_Search_Multimodal unchanged."""
    import torch
    from embrace_b200.BIOINF_tesi.models import ConcatNetMultimodal, ConcatNetMultimodal_NoTrain
    from embrace_b200.BIOINF_tesi.models.utils import Param_Search_Multimodal
    from embrace_b200.BIOINF_tesi.data_pipe import PackedDataset, build_loaders
    from tests.golden.cases import CASES
    from tests.golden.ref_harness import FixedTrial, spec_to_trial_params
    monkeypatch.chdir(tmp_path)
    spec = CASES['concat_small']['spec']
    P = O.init_params(spec, 17)
    tp = spec_to_trial_params(spec)
    model = ConcatNetMultimodal(FixedTrial(tp), 'A549', 'active_E_vs_inactive_E', spec['F'], 'cuda', precision='fp32')
    assert list(model.state_dict()) == list(O.param_shapes(spec))
    model.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in P.items()})
    x, codes, _ = make_inputs(spec, 9, 18)
    model.eval()
    with torch.no_grad():
        got = model([torch.from_numpy(x), torch.from_numpy(codes)]).cpu().numpy()
    ref, _ = O.forward(spec, P, x, codes, training=False)
    assert np.abs(got - ref).max() <= 2e-5 * max(1.0, np.abs(ref).max())
    torch.save({'model_state_dict': model.state_dict(), 'model_params': tp}, 'A549_ConcatNetMultimodal_active_E_vs_inactive_E_1_test_.pt')
    twin = ConcatNetMultimodal_NoTrain('A549', 'active_E_vs_inactive_E', 1, spec['F'], 'cuda', precision='fp32')
    twin.load_state_dict(torch.load('A549_ConcatNetMultimodal_active_E_vs_inactive_E_1_test_.pt', weights_only=False)['model_state_dict'])
    twin.eval()
    with torch.no_grad():
        assert np.abs(twin([torch.from_numpy(x), torch.from_numpy(codes)]).cpu().numpy() - ref).max() <= 2e-5 * max(1.0, np.abs(ref).max())
    xs, cs, ys = _data(500, 48, 7)
    tr = build_loaders(PackedDataset(xs[:350], cs[:350], ys[:350]), batch_size=100, training=True)
    te = build_loaders(PackedDataset(xs[350:], cs[350:], ys[350:]), batch_size=100, training=False)
    ps = Param_Search_Multimodal(ConcatNetMultimodal, tr, te, num_epochs=1, study_name='A549_concat_1', device='cuda', cell_line='A549',
                                 task='active_E_vs_inactive_E', sampler='random', n_trials=2, storage='c.db', seed=3)
    ps.run_trial()
    assert 'CONCATNET_n_post_layers' in ps.best_params and type(ps.best_model).__name__ == 'ConcatNetMultimodal'
