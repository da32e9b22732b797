"""Host-side mirror of BIOINF_tesi.models: what can be checked without a GPU."""
import os

import numpy as np
import pytest
import torch

import embrace_b200
from embrace_b200.BIOINF_tesi.models import FFNN_pre, CNN_pre, FFNN_pre_NoTrain, CNN_pre_NoTrain
from embrace_b200.BIOINF_tesi.models.utils import (AUPRC, F1_precision_recall, EarlyStopping, get_loss_weights_from_labels,
                                                   output_size_from_model_params, get_single_model_params, size_out_convolution)
from embrace_b200.BIOINF_tesi.models.utils.training_models_multimodal import lift_optimizer, fit_multimodal
from embrace_b200.BIOINF_tesi.models.utils import optim as emb_optim
from oracle import embracenet_oracle as O
from tests.golden.cases import ARCH_S
from tests.golden.ref_harness import FixedTrial, spec_to_trial_params

GOLD = os.path.join(os.path.dirname(__file__), 'golden')


def test_metric_helpers_match_the_reference_fixture():
    rows = np.load(os.path.join(GOLD, 'metrics.npz'))['rows']
    for tp, fp, fn, tn, a, f_p, f_r, f_1, wp, wn in rows[:120]:
        tp, fp, fn, tn = int(tp), int(fp), int(fn), int(tn)
        y = torch.tensor([1] * tp + [0] * fp + [1] * fn + [0] * tn).reshape(-1, 1)
        out = torch.tensor([[0., 1.]] * (tp + fp) + [[1., 0.]] * (fn + tn))
        assert abs(AUPRC(out, y) - a) < 1e-12
        np.testing.assert_allclose(F1_precision_recall(out, y), [f_p, f_r, f_1], atol=1e-12)
        np.testing.assert_allclose(get_loss_weights_from_labels(y), [wp, wn], atol=1e-15)


def test_container_modules_have_the_reference_key_names():
    tp = spec_to_trial_params(ARCH_S)
    f = FFNN_pre(FixedTrial(tp), 48, device='cuda')
    c = CNN_pre(FixedTrial(tp), device='cuda')
    shapes = O.param_shapes(ARCH_S)
    got = {'FFNN.' + k: tuple(v.shape) for k, v in f.state_dict().items()}
    got.update({'CNN.' + k: tuple(v.shape) for k, v in c.state_dict().items()})
    want = {k: tuple(v) for k, v in shapes.items() if k.startswith(('FFNN.', 'CNN.'))}
    assert got == want
    assert f.output_size == 32 and c.output_size == 1600
    single = get_single_model_params(tp)
    assert output_size_from_model_params(single['CNN']) == 1600
    f2 = FFNN_pre_NoTrain(48, single['FFNN'], device='cuda')
    c2 = CNN_pre_NoTrain(single['CNN'], device='cuda')
    assert list(f2.state_dict()) == list(f.state_dict()) and list(c2.state_dict()) == list(c.state_dict())
    with pytest.raises(RuntimeError):
        f(torch.zeros(1, 48))          # containers never compute: no PyTorch fallback


def test_lift_optimizer():
    p = [torch.nn.Parameter(torch.zeros(3))]
    c = lift_optimizer(torch.optim.Adam(p, lr=4.1e-5, weight_decay=7.6e-4))
    assert c.kind == 0 and abs(c.lr - 4.1e-5) < 1e-12 and abs(c.weight_decay - 7.6e-4) < 1e-9 and abs(c.beta2 - 0.999) < 1e-7
    assert lift_optimizer(torch.optim.RMSprop(p, lr=1e-3)).kind == 3
    c = lift_optimizer(emb_optim.Nadam(p, lr=1e-3, weight_decay=1e-2))
    assert c.kind == 2 and abs(c.momentum_decay - 4e-3) < 1e-9
    assert lift_optimizer(torch.optim.AdamW(p, lr=1e-3)).kind == 1
    with pytest.raises(ValueError):
        lift_optimizer(torch.optim.SGD(p, lr=0.1))
    with pytest.raises(ValueError):
        lift_optimizer(torch.optim.RMSprop(p, lr=1e-3, momentum=0.9))


def test_fit_multimodal_argument_errors_match_the_reference():
    with pytest.raises(ValueError):
        fit_multimodal(None, {}, {}, 'cuda', 'HELA', 'active_E_vs_inactive_E', checkpoint_path='x')
    with pytest.raises(ValueError):
        fit_multimodal(None, {}, {}, 'cuda', 'A549', 'nonsense', checkpoint_path='x')
    with pytest.raises(TypeError):       # checkpoint_path=None crashes in os.path.exists, as upstream
        fit_multimodal(None, {}, {}, 'cuda', 'A549', 'active_E_vs_inactive_E', checkpoint_path=None)


def test_models_refuse_to_run_without_a_gpu():
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    from embrace_b200.BIOINF_tesi.models import EmbraceNetMultimodal
    with pytest.raises(embrace_b200.EmbError):
        EmbraceNetMultimodal(FixedTrial(spec_to_trial_params(ARCH_S)), cell_line='A549', task='active_E_vs_inactive_E', device='cuda',
                             in_features_FFNN=48)
    with pytest.raises(embrace_b200.EmbError):
        EmbraceNetMultimodal(FixedTrial(spec_to_trial_params(ARCH_S)), cell_line='A549', task='active_E_vs_inactive_E', device='cpu',
                             in_features_FFNN=48)


def test_misc_helpers():
    assert size_out_convolution(256, 5, 2, 1) == 256 and size_out_convolution(256, 10, 0, 2) == 124
    es = EarlyStopping(patience=2, trace_func=lambda *_: None)
    for s in (0.3, 0.3, 0.2, 0.25):
        es(s)
    assert es.early_stop
