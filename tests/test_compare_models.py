"""checkpoint interop + the batched Compare_Models_Result predict path.

Golden: tests/golden/compare_models.npz and tests/golden/ckpt/*.pt, written by the UNMODIFIED reference
(`Compare_Models_Result.get_model_predictions` over checkpoints saved from the reference's training classes;
tests/golden/make_golden_compare.py).  CPU tests pin the oracle to it; the GPU test runs the mirror class through the
engine (C ABI) and compares predictions and Wilcoxon p-values."""
import os
import shutil

import numpy as np
import pytest
import torch

from oracle import embracenet_oracle as O

GOLD = os.path.join(os.path.dirname(__file__), 'golden')
CELL, TASK, FOLD = 'A549', 'active_E_vs_inactive_E', 1
MODELS = ['EmbraceNetMultimodal', 'EmbraceNetMultimodal_augmentation', 'ConcatNetMultimodal', 'FFNN', 'CNN']
KIND = {'EmbraceNetMultimodal': 'embracenet', 'EmbraceNetMultimodal_augmentation': 'embracenet',
        'ConcatNetMultimodal': 'concatnet', 'FFNN': 'ffnn', 'CNN': 'cnn'}


def _gold():
    return np.load(os.path.join(GOLD, 'compare_models.npz'))


def _ckpt(model):
    return os.path.join(GOLD, 'ckpt', f'{CELL}_{model}_{TASK}_{FOLD}_test_.pt')


def _oracle_spec(model_params, kind, F):
    """reference trial-parameter names (as stored in the checkpoint) -> oracle spec dict."""
    pf, pc = ('FFNN_', 'CNN_') if kind in ('embracenet', 'concatnet') else ('', '')
    s = dict(kind=kind, F=F)
    if kind != 'cnn':
        n = model_params[f'{pf}n_layers']
        s['ffnn_units'] = [model_params[f'{pf}n_units_l{i}'] for i in range(n)]
        s['ffnn_dropout'] = [model_params[f'{pf}dropout_l{i}'] for i in range(n)]
    if kind != 'ffnn':
        n = model_params[f'{pc}n_layers']
        s['cnn_channels'] = [model_params[f'{pc}out_channels_l{i}'] for i in range(n)]
        s['cnn_kernels'] = [model_params[f'{pc}kernel_size_l{i}'] for i in range(n)]
        s['cnn_dropout'] = [model_params[f'{pc}dropout_l{i}'] for i in range(n)]
    if kind == 'embracenet':
        n = model_params['n_post_layers']
        s.update(C=model_params['EMBRACENET_embracement_size'], p_ffnn=model_params['selection_probabilities_FFNN'],
                 post_units=[model_params[f'EMBRACENET_n_units_l{i}'] for i in range(n)],
                 post_dropout=[model_params[f'EMBRACENET_dropout_l{i}'] for i in range(n)])
    if kind == 'concatnet':
        n = model_params['CONCATNET_n_post_layers']
        s.update(post_units=[model_params[f'CONCATNET_n_units_l{i}'] for i in range(n)],
                 post_dropout=[model_params[f'CONCATNET_dropout_l{i}'] for i in range(n)])
    return s


@pytest.mark.parametrize('model', MODELS)
def test_reference_checkpoints_load_and_oracle_matches(model):
    """The committed .pt files are what the reference writes (`model_state_dict` fp64 + `model_params`); the oracle's eval
    forward on their (fp32-rounded, visual.py:274-279) weights reproduces the reference's per-region predictions."""
    g = _gold()
    ck = torch.load(_ckpt(model), map_location='cpu', weights_only=False)
    assert set(ck) == {'model_state_dict', 'model_params'}
    kind = KIND[model]
    spec = _oracle_spec(ck['model_params'], kind, g['x1'].shape[1])
    assert set(ck['model_state_dict']) == set(O.param_shapes(spec))
    P = {}
    for k, v in ck['model_state_dict'].items():
        a = v.numpy()
        P[k] = a.astype(np.float32).astype(np.float64) if a.dtype == np.float64 else a
    draws = {'embrace_u': g[f'u_{model}']} if kind == 'embracenet' else {}
    logits, _ = O.forward(spec, P, g['x1'], g['bases'], draws, training=False)
    if kind == 'concatnet':
        out = logits[:, 1]                      # ConcatNetMultimodal_NoTrain.py:87 drops the softmax (typo)
    else:
        e = np.exp(logits - logits.max(axis=1, keepdims=True))
        out = (e / e.sum(axis=1, keepdims=True))[:, 1]
    np.testing.assert_allclose(out, g[f'pred_{model}'], atol=2e-7)      # the reference returns a fp32 torch.tensor([...])


def test_mirror_surface_without_gpu():
    from embrace_b200.BIOINF_tesi.visual import Compare_Models_Result
    from embrace_b200.BIOINF_tesi import models as M
    c = Compare_Models_Result()
    assert c.models_dict == {'EmbraceNetMultimodal': M.EmbraceNetMultimodal_NoTrain,
                             'EmbraceNetMultimodal_augmentation': M.EmbraceNetMultimodal_NoTrain,
                             'ConcatNetMultimodal': M.ConcatNetMultimodal_NoTrain, 'FFNN': M.FFNN_NoTrain, 'CNN': M.CNN_NoTrain}
    with pytest.raises(ValueError):
        c('cuda')                               # no data provider
    if not torch.cuda.is_available():           # no CPU fallback: building a twin without a B200 fails loudly
        g = _gold()
        c.set_data(g['x1'], g['bases'])
        cwd = os.getcwd()
        os.chdir(os.path.join(GOLD, 'ckpt'))
        try:
            with pytest.raises(Exception):
                c.get_model_predictions(CELL, TASK, 'FFNN', FOLD, 'cuda')
        finally:
            os.chdir(cwd)


@pytest.mark.gpu
def test_compare_models_result_matches_reference(tmp_path):
    from embrace_b200.BIOINF_tesi.visual import Compare_Models_Result
    g = _gold()
    for m in MODELS:
        shutil.copy(_ckpt(m), tmp_path)
    cwd = os.getcwd()
    os.chdir(tmp_path)
    try:
        c = Compare_Models_Result(precision='fp32', batch_size=10)           # 24 regions -> three engine batches
        rows = [torch.from_numpy(g['x1'][i:i + 1]) for i in range(len(g['x1']))]
        c.set_data(rows, O.onehot_from_bases(g['bases']))                      # Series-of-rows and one-hot forms
        preds = {}
        for m in MODELS:
            draws = {'embrace_u': g[f'u_{m}']} if KIND[m] == 'embracenet' else None
            preds[m] = c.get_model_predictions(CELL, TASK, m, FOLD, 'cuda', draws=draws).numpy()
            np.testing.assert_allclose(preds[m], g[f'pred_{m}'], atol=5e-6, err_msg=m)
        from scipy.stats import wilcoxon
        for b in MODELS[:2]:
            for m in MODELS:
                if m != b:
                    p = wilcoxon(preds[b].astype(np.float64), preds[m].astype(np.float64))[1]
                    np.testing.assert_allclose(p, float(g[f'pval_{b}__{m}']), rtol=1e-6, atol=1e-12)
        # the whole __call__ (Philox selection draws, bf16 production precision): structure, cache file, verdicts
        c2 = Compare_Models_Result()
        pv = c2('cuda', n_folds=1, cell_lines=CELL, tasks=TASK, data=lambda task, cell: (g['x1'], g['bases']))
        fold = pv[TASK][CELL]['1']
        assert set(fold) == {'EmbraceNetMultimodal', 'EmbraceNetMultimodal_augmentation'}
        assert set(fold['EmbraceNetMultimodal']) == {'FFNN', 'CNN', 'ConcatNetMultimodal', 'EmbraceNetMultimodal_augmentation'}
        assert fold['EmbraceNetMultimodal']['FFNN'] < 1e-5            # P(class 1) 0.59 vs 0.47 on every region
        assert 0.0 <= fold['EmbraceNetMultimodal']['EmbraceNetMultimodal_augmentation'] <= 1.0   # depends on the Philox selection draws
        assert os.path.exists(f'pval_results_dict_{TASK}.pickle')
        c3 = Compare_Models_Result()
        assert c3('cuda', pval_dict=pv) is pv
    finally:
        os.chdir(cwd)


def test_print_model_difference_counts_like_the_reference(capsys):
    """visual.py:298-325: significant (p < 0.05, or NaN: the reference tests `p >= p_val`) in at least two folds -> different."""
    from embrace_b200.BIOINF_tesi.visual import Compare_Models_Result
    c = Compare_Models_Result()
    pv = {'t': {'A549': {'1': {'E': {'FFNN': 0.01, 'CNN': 0.5, 'X': float('nan')}},
                         '2': {'E': {'FFNN': 0.04, 'CNN': 0.01, 'X': float('nan')}},
                         '3': {'E': {'FFNN': 0.9, 'CNN': 0.2, 'X': 0.7}}}}}
    assert c('cuda', pval_dict=pv) is pv
    assert c.counter_dict['t']['A549']['E'] == {'FFNN': 2, 'CNN': 1, 'X': 2}
    out = capsys.readouterr().out
    assert 'FFNN ===> different: True' in out and 'CNN ===> different: False' in out and 'BASE MODEL: E' in out
