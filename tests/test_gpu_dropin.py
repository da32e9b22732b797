"""The reference-facing surface (BIOINF_tesi.models mirrors, fit_multimodal, the _NoTrain predict pattern) on a B200."""
import os

import numpy as np
import pytest

from oracle import embracenet_oracle as O
from tests.golden.cases import CASES, ARCH_S, make_inputs
from tests.golden.ref_harness import FixedTrial, spec_to_trial_params

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), 'golden')


def build(spec, P, precision='fp32'):
    import torch
    from embrace_b200.BIOINF_tesi.models import EmbraceNetMultimodal
    m = EmbraceNetMultimodal(FixedTrial(spec_to_trial_params(spec)), cell_line='A549', task='active_E_vs_inactive_E', device='cuda',
                             in_features_FFNN=spec['F'], precision=precision)
    m.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in P.items()})
    return m.double().to('cuda')


def test_state_dict_surface_matches_reference():
    import torch
    spec = ARCH_S
    P = O.init_params(spec, 1)
    m = build(spec, P)
    sd = m.state_dict()
    shapes = O.param_shapes(spec)
    assert list(sd) == list(shapes)
    assert {k: tuple(v.shape) for k, v in sd.items()} == {k: tuple(v) for k, v in shapes.items()}
    for k, v in sd.items():
        if not k.endswith('num_batches_tracked'):
            np.testing.assert_allclose(v.cpu().numpy(), P[k].astype(np.float32), rtol=0, atol=0)
    assert m.FFNN_pre_output_size == 32 and m.CNN_pre_output_size == 1600
    assert tuple(m.selection_probabilities.shape) == (2,) and 'selection_probabilities' not in sd
    # weight_reset through .apply re-initialises Linear/Conv1d in place (arena views) and leaves BatchNorm alone
    from embrace_b200.BIOINF_tesi.models.utils import weight_reset
    before = m.state_dict()
    m.apply(weight_reset)
    after = m.state_dict()
    assert not torch.equal(before['post.0.weight'], after['post.0.weight'])
    assert torch.equal(before['CNN.CNN_model.1.weight'], after['CNN.CNN_model.1.weight'])
    assert torch.equal(before['CNN.CNN_model.1.running_var'], after['CNN.CNN_model.1.running_var'])


def test_reference_call_pattern_with_torch_loss_and_optimizer():
    """output = model([x1, x2], is_training=True); loss = criterion(output.float(), y); loss.backward(); optimizer.step()"""
    import torch
    case = CASES['small2']
    spec, B = case['spec'], case['B']
    P = O.init_params(spec, case['seed'])
    x, bases, y = make_inputs(spec, B, case['seed'] + 1)
    g = np.load(os.path.join(GOLD, 'case_small2.npz'))
    m = build(spec, P)
    opt = torch.optim.Adam(m.parameters(), lr=1e-2, weight_decay=1e-2)
    x1, x2 = torch.from_numpy(x), torch.from_numpy(O.onehot_from_bases(bases))
    target = torch.from_numpy(y.reshape(-1, 1)).cuda()
    from embrace_b200.BIOINF_tesi.models.utils import get_loss_weights_from_labels, AUPRC
    m.train()
    ref_P = {k: v.copy() for k, v in P.items()}
    st = O.opt_init(ref_P, 'adam')
    for step in range(2):
        draws = O.make_draws(spec, B, case['seed'] + 100 + step, force_modal=case['force_modal'][step])
        ref = O.train_step(spec, ref_P, x, bases, y, draws, st, lr=1e-2, wd=1e-2)
        w_pos, w_neg = get_loss_weights_from_labels(target)
        criterion = torch.nn.CrossEntropyLoss(weight=torch.tensor([w_neg, w_pos]))
        opt.zero_grad()
        output = m([x1.double(), x2.double()], is_training=True, draws=draws)
        loss = criterion.float().to('cuda')(output.float(), target.squeeze())
        loss.backward()
        opt.step()
        np.testing.assert_allclose(output.detach().cpu().numpy(), ref['logits'], atol=3e-5 * np.abs(ref['logits']).max() * (1 + 20 * step))
        assert abs(loss.item() - float(ref['loss'])) < 2e-5
        assert abs(AUPRC(output, target) - ref['auprc']) < 1e-12
        if step == 0:
            np.testing.assert_allclose(output.detach().cpu().numpy(), g['s0_logits'], atol=3e-5)
            for k, gr in ref['grads'].items():
                got = dict(m.named_parameters())[k].grad.cpu().numpy()
                assert np.abs(got - gr).max() <= 3e-4 * max(np.abs(gr).max(), 1e-3), k
    sd = m.state_dict()
    assert int(sd['CNN.CNN_model.1.num_batches_tracked']) == 2
    for k in ('CNN.CNN_model.1.running_mean', 'CNN.CNN_model.6.running_var'):
        np.testing.assert_allclose(sd[k].cpu().numpy(), ref_P[k], rtol=2e-5, atol=1e-7)


def test_fit_multimodal_matches_reference_fit():
    """Same lists as the reference's fit_multimodal on the golden loaders (fp32 precision, replayed draws)."""
    import tempfile
    import torch
    from embrace_b200.BIOINF_tesi.models.utils import fit_multimodal
    g = np.load(os.path.join(GOLD, 'fit_small2.npz'))
    spec = CASES['small2']['spec']
    P = O.init_params(spec, 777)
    nb_train, nb_test, Btr, Bte, epochs = 3, 2, 8, 6, 3
    batches = [make_inputs(spec, Btr if b < nb_train else Bte, 800 + b) for b in range(nb_train + nb_test)]
    train, test = {'FFNN': [], 'CNN': []}, {'FFNN': [], 'CNN': []}
    for b, (xf, bs, y) in enumerate(batches):
        dst = train if b < nb_train else test
        dst['FFNN'].append((torch.from_numpy(xf), torch.from_numpy(y.reshape(-1, 1))))
        dst['CNN'].append((torch.from_numpy(O.onehot_from_bases(bs)), torch.from_numpy(y.reshape(-1, 1))))
    m = build(spec, P)
    opt = torch.optim.Adam(m.parameters(), lr=5e-3, weight_decay=1e-3)
    with tempfile.TemporaryDirectory() as td:
        ck = os.path.join(td, 'ck.pt')
        a_tr, a_te, f1 = fit_multimodal(m, train, test, 'cuda', 'A549', 'active_E_vs_inactive_E', optimizer=opt, num_epochs=epochs,
                                        patience=10, verbose=False, checkpoint_path=ck,
                                        draws_train=lambda ep, b: O.make_draws(spec, Btr, 10000 + ep * 100 + b),
                                        draws_test=lambda ep, b: O.make_draws(spec, Bte, 20000 + ep * 100 + b))
        np.testing.assert_allclose(a_tr, g['auprc_train'], atol=1e-12)
        np.testing.assert_allclose(a_te, g['auprc_test'], atol=1e-12)
        np.testing.assert_allclose(np.array(f1), g['f1pr_test'], atol=1e-12)
        np.testing.assert_allclose(m.state_dict()['post.3.weight'].cpu().numpy(), g['final_logit_w'], rtol=2e-3, atol=2e-5)
        saved = torch.load(ck, weights_only=False)
        assert set(saved) == {'model_state_dict', 'AUPRC_train_scores', 'AUPRC_test_scores', 'F1_precision_recall_test_scores'}
        # second call short-circuits on the checkpoint, like the reference (:95-100)
        again = fit_multimodal(m, train, test, 'cuda', 'A549', 'active_E_vs_inactive_E', optimizer=opt, num_epochs=epochs,
                               checkpoint_path=ck)
        np.testing.assert_allclose(again[1], a_te)


def test_notrain_predict_pattern():
    """visual.py:263-295: rebuild from model_params, load_state_dict, .double().to(device), eval(), batch-1 loop taking [1]."""
    import tempfile
    import torch
    from embrace_b200.BIOINF_tesi.models import EmbraceNetMultimodal_NoTrain
    gold = np.load(os.path.join(GOLD, 'notrain_small2.npz'))['probs']
    spec, B = CASES['small2']['spec'], CASES['small2']['B']
    P = O.init_params(spec, 4242)
    x, bases, _ = make_inputs(spec, B, 4243)
    u = np.random.RandomState(4244).random_sample((B, spec['C']))
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as td:
        os.chdir(td)
        try:
            sd = {k: torch.from_numpy(np.asarray(v).copy()) for k, v in P.items()}
            torch.save({'model_state_dict': sd, 'model_params': spec_to_trial_params(spec)},
                       'A549_EmbraceNetMultimodal_active_E_vs_inactive_E_1_test_.pt')
            model_ = EmbraceNetMultimodal_NoTrain('A549', 'active_E_vs_inactive_E', 1, spec['F'], device='cuda', precision='fp32')
            state = torch.load('A549_EmbraceNetMultimodal_active_E_vs_inactive_E_1_test_.pt', weights_only=False)
            model_.load_state_dict(state['model_state_dict'])
        finally:
            os.chdir(cwd)
    model_.double().to('cuda')
    model_.eval()
    x1, x2 = torch.from_numpy(x), torch.from_numpy(O.onehot_from_bases(bases))
    with torch.no_grad():
        out = torch.tensor([model_([x1[i:i + 1], x2[i:i + 1]], draws={'embrace_u': u[i:i + 1]})[1] for i in range(B)])
    np.testing.assert_allclose(out.numpy(), gold, atol=3e-6)
    batched = model_.predict_proba(x1, x2, draws={'embrace_u': u})
    np.testing.assert_allclose(batched.cpu().numpy(), gold, atol=3e-6)


def test_single_modality_models_and_fit():
    import tempfile
    import torch
    from embrace_b200.BIOINF_tesi.models import FFNN, CNN
    from embrace_b200.BIOINF_tesi.models.utils import fit
    for name, cls in (('ffnn_only', FFNN), ('cnn_only', CNN)):
        case = CASES[name]
        spec, B = case['spec'], case['B']
        P = O.init_params(spec, case['seed'])
        x, bases, y = make_inputs(spec, B, case['seed'] + 1)
        trial = FixedTrial(spec_to_trial_params(spec))
        m = cls(trial, spec['F'], 'cuda', precision='fp32') if name == 'ffnn_only' else cls(trial, 'cuda', precision='fp32')
        assert list(m.state_dict()) == list(O.param_shapes(spec))
        m.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in P.items()})
        g = np.load(os.path.join(GOLD, f'case_{name}.npz'))
        draws = O.make_draws(spec, B, case['seed'] + 100)
        m.train()
        inp = torch.from_numpy(x) if name == 'ffnn_only' else torch.from_numpy(O.onehot_from_bases(bases))
        out = m(inp.double(), draws=draws)
        np.testing.assert_allclose(out.detach().cpu().numpy(), g['s0_logits'], atol=3e-5)
        loader = [(inp, torch.from_numpy(y.reshape(-1, 1)))] * 3
        with tempfile.TemporaryDirectory() as td:
            a_tr, a_te, f1 = fit(m, loader, loader[:1], 'cuda', optimizer=torch.optim.Adam(m.parameters(), lr=1e-3), num_epochs=2,
                                 verbose=False, checkpoint_path=os.path.join(td, 'c.pt'))
        assert len(a_tr) == 2 and len(a_te) == 2 and f1[0].shape == (3,)


def test_training_curve_auprc_parity_on_planted_signal():
    """AUPRC parity after training (north_star: within 0.002): the engine and the PyTorch-CPU port of the reference train the
    same model on the same planted-signal data with the same replayed draws -- 200 Adam steps at lr 3e-3 and 40 at 3e-4 -- and
    both AUPRC definitions are compared on 8192 held-out rows.

    Why a converged protocol: training is a chaotic map of its rounding.  The REFERENCE re-run with its initial weights
    perturbed by 1e-6 relative lands, after 40 steps in the steep part of the learning curve, up to 0.033 (ranking AUPRC) away
    from its own unperturbed run on the weak-signal task round 1 used; once converged on this task it lands 0.001 (ranking) /
    0.0013 (hard) away (r2, fp64 port, /tmp experiment recorded in DESIGN.md 2).  The engine runs in deterministic mode so that
    the outcome of this test is reproducible."""
    import torch
    from sklearn.metrics import average_precision_score
    from oracle import torch_port as TP
    from embrace_b200.BIOINF_tesi.models.utils.training_models_multimodal import lift_optimizer
    from embrace_b200 import _native as NAT
    spec = dict(kind='embracenet', F=16, ffnn_units=[32, 16], ffnn_dropout=[0.2, 0.0], cnn_channels=[16, 32], cnn_kernels=[5, 5],
                cnn_dropout=[0.2, 0.0], C=64, post_units=[32], post_dropout=[0.2], p_ffnn=0.5)
    rs = np.random.RandomState(3)
    N, B, steps, decay_at = 2048, 128, 240, 200

    def data(n):
        x = rs.random_sample((n, spec['F'])).astype(np.float32).astype(np.float64)
        bases = rs.randint(0, 4, size=(n, 256)).astype(np.uint8)
        motif = np.array([0, 2, 2, 1, 3, 0], dtype=np.uint8)
        y = (rs.random_sample(n) < 1 / (1 + np.exp(-(3.0 * 6 * (x[:, :4].mean(1) - 0.5) - 1.0)))).astype(np.int64)
        for i in np.nonzero(y)[0]:
            if rs.random_sample() < 0.7:
                pos = rs.randint(0, 250)
                bases[i, pos:pos + 6] = motif
        return x, bases, y
    xtr, btr, ytr = data(N)
    xte, bte, yte = data(8192)
    P = O.init_params(spec, 17)
    torch.set_num_threads(os.cpu_count() or 1)
    st = TP.TrainState(spec, {k: v.copy() for k, v in P.items()}, 'adam', lr=3e-3, wd=1e-4)
    NAT.set_option('deterministic', 1)
    try:
        m = build(spec, P, precision='fp32')
        opt = torch.optim.Adam(m.parameters(), lr=3e-3, weight_decay=1e-4)
        m.train()
        for s in range(steps):
            lr = 3e-3 if s < decay_at else 3e-4
            for g in list(st.opt.param_groups) + list(opt.param_groups):
                g['lr'] = lr
            cfg = lift_optimizer(opt)
            lo = (s * B) % N
            xb, bb, yb = xtr[lo:lo + B], btr[lo:lo + B], ytr[lo:lo + B]
            draws = O.make_draws(spec, B, 5000 + s)
            st.step(torch.from_numpy(xb), torch.from_numpy(O.onehot_from_bases(bb)), yb, draws)
            m.train_batch(torch.from_numpy(xb), torch.from_numpy(bb), torch.from_numpy(yb), cfg, draws=draws)
    finally:
        NAT.set_option('deterministic', 0)
    u = np.random.RandomState(9).random_sample((len(yte), spec['C']))
    Pt = {k: v.detach().numpy() for k, v in st.T.items()}
    ref_logits, _ = O.forward(spec, Pt, xte, bte, {'embrace_u': u}, training=False)
    m.eval()
    got_logits = m([torch.from_numpy(xte), torch.from_numpy(bte)], draws={'embrace_u': u}).cpu().numpy()
    s_ref, s_got = ref_logits[:, 1] - ref_logits[:, 0], got_logits[:, 1] - got_logits[:, 0]
    rank_ref, rank_got = average_precision_score(yte, s_ref), average_precision_score(yte, s_got)
    hard_ref, hard_got = O.auprc_hard(ref_logits, yte), O.auprc_hard(got_logits, yte)
    print('ranking AUPRC ref/got', rank_ref, rank_got, 'hard AUPRC ref/got', hard_ref, hard_got, 'base rate', yte.mean())
    assert rank_ref > yte.mean() + 0.3, 'the planted signal must be learnable'
    assert abs(rank_ref - rank_got) <= 0.002
    assert abs(hard_ref - hard_got) <= 0.004
