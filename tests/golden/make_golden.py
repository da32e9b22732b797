"""Generate tests/golden/*.npz from the UNMODIFIED reference (run in the build
container only:  python tests/golden/make_golden.py ).

Every fixture holds outputs of the reference's own classes / functions
(BIOINF_tesi.models.* imported from /root/reference, PyTorch CPU fp64) on seeded
synthetic inputs, with the random draws replayed from explicit tensors
(ref_harness.replay_draws).  Inputs, weights and draws are NOT stored: they are
regenerated from the seeds by oracle.make_draws / oracle.init_params / cases.make_inputs
(numpy legacy RandomState, stable across versions).
"""
import json
import os
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, '..', '..'))
sys.path.insert(0, HERE)

from oracle import embracenet_oracle as O      # noqa: E402
import ref_harness as RH                       # noqa: E402
from cases import CASES, BENCH_CASES, make_inputs, compress  # noqa: E402


def run_case(M, name, case):
    spec, B = case['spec'], case['B']
    P0 = O.init_params(spec, case['seed'])
    x_ffnn, bases, y = make_inputs(spec, B, case['seed'] + 1)
    model = RH.build_reference_model(M, spec, P0)
    kind = spec.get('kind', 'embracenet')
    out = {'spec_json': np.array(json.dumps(spec)), 'B': np.array(B), 'seed': np.array(case['seed'])}
    opt_kind = case.get('opt', 'adam')
    lr, wd = case.get('lr', 1e-2), case.get('wd', 1e-2)
    if opt_kind == 'adam':
        opt = torch.optim.Adam(model.parameters(), lr=lr, weight_decay=wd)
    elif opt_kind == 'rmsprop':
        opt = torch.optim.RMSprop(model.parameters(), lr=lr, weight_decay=wd)
    elif opt_kind == 'nadam':   # timm.optim.Nadam stand-in: same published rule
        opt = torch.optim.NAdam(model.parameters(), lr=lr, weight_decay=wd, momentum_decay=4e-3)
    x1 = torch.from_numpy(x_ffnn).double()
    x2 = torch.from_numpy(O.onehot_from_bases(bases)).double()
    target = torch.from_numpy(y.reshape(-1, 1))
    # the reference's own helpers for the loss weights and the metric
    from BIOINF_tesi.models.utils.utils import get_loss_weights_from_labels, AUPRC, F1_precision_recall
    model.train()
    for step in range(case.get('steps', 2)):
        draws = O.make_draws(spec, B, case['seed'] + 100 + step, force_modal=case.get('force_modal', [None, None])[step])
        queue = RH.draws_to_queue(spec, draws, training=True)
        w_pos, w_neg = get_loss_weights_from_labels(target)
        criterion = torch.nn.CrossEntropyLoss(weight=torch.tensor([w_neg, w_pos]))
        opt.zero_grad()
        with RH.replay_draws(queue) as log:
            if kind == 'embracenet':
                hook_idx = {}
                orig = torch.multinomial

                def spy(*a, **k):
                    r = orig(*a, **k)
                    hook_idx['idx'] = r.clone()
                    return r
                torch.multinomial = spy
                output = model([x1, x2], is_training=True)
                torch.multinomial = orig
            elif kind == 'concatnet':
                output = model([x1, x2])
            elif kind == 'ffnn':
                output = model(x1)
            else:
                output = model(x2)
        assert not queue, f'{name}: unused draws {[(t, a.shape) for t, a in queue]}'
        loss = criterion.float()(output.float(), target.squeeze(1))
        loss.backward()
        pre = f's{step}_'
        out[pre + 'logits'] = output.detach().numpy().copy()
        out[pre + 'loss'] = np.array(loss.item(), dtype=np.float64)
        out[pre + 'auprc'] = np.array(AUPRC(output, target), dtype=np.float64)
        out[pre + 'f1pr'] = np.asarray(F1_precision_recall(output, target), dtype=np.float64)
        if kind == 'embracenet':
            out[pre + 'idx'] = np.packbits(hook_idx['idx'].numpy().astype(np.uint8), axis=1)
        for k, p in model.named_parameters():
            compress(out, pre + 'grad.' + k, p.grad.detach().numpy())
        opt.step()
        for k, v in model.state_dict().items():
            compress(out, pre + 'param.' + k, v.detach().numpy())
    # eval-mode forward with the trained weights/buffers (multinomial still sampled: quirk 2)
    model.eval()
    draws = O.make_draws(spec, B, case['seed'] + 900)
    with torch.no_grad(), RH.replay_draws(RH.draws_to_queue(spec, draws, training=False)):
        if kind == 'embracenet':
            ev = model([x1, x2])
            av = np.ones((B, 2), dtype=np.float32)
            av[0::3, 0] = 0
            av[1::3, 1] = 0
        elif kind == 'concatnet':
            ev = model([x1, x2])
        elif kind == 'ffnn':
            ev = model(x1)
        else:
            ev = model(x2)
    out['eval_logits'] = ev.numpy().copy()
    if kind == 'embracenet':
        with torch.no_grad(), RH.replay_draws(RH.draws_to_queue(spec, draws, training=False)):
            ev2 = model([x1, x2], availabilities=torch.from_numpy(av))
        out['eval_logits_avail'] = ev2.numpy().copy()
        out['eval_avail'] = av
    np.savez_compressed(os.path.join(HERE, f'case_{name}.npz'), **out)
    print(f'case {name}: loss0={out["s0_loss"]:.6f} logits0[0]={out["s0_logits"][0]}')


def notrain_fixture(M):
    """EmbraceNetMultimodal_NoTrain rebuilt from a checkpoint's model_params (visual.py:263-295 pattern)."""
    case = CASES['small2']
    spec, B = case['spec'], case['B']
    P0 = O.init_params(spec, 4242)
    x_ffnn, bases, _ = make_inputs(spec, B, 4243)
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as td:
        os.chdir(td)
        try:
            sd = {k: torch.from_numpy(np.asarray(v).copy()) for k, v in P0.items()}
            torch.save({'model_state_dict': sd, 'model_params': RH.spec_to_trial_params(spec)},
                       'A549_EmbraceNetMultimodal_active_E_vs_inactive_E_1_test_.pt')
            model = M.EmbraceNetMultimodal_NoTrain('A549', 'active_E_vs_inactive_E', 1, spec['F'], device='cpu')
            ck = torch.load('A549_EmbraceNetMultimodal_active_E_vs_inactive_E_1_test_.pt')
            model.load_state_dict(ck['model_state_dict'])
            model.double().eval()
        finally:
            os.chdir(cwd)
    rs = np.random.RandomState(4244)
    u = rs.random_sample((B, spec['C']))
    probs = []
    with torch.no_grad():
        for i in range(B):           # the reference's batch-1 loop
            with RH.replay_draws([('multinomial', u[i:i + 1])]):
                o = model([torch.from_numpy(x_ffnn[i:i + 1]).double(),
                           torch.from_numpy(O.onehot_from_bases(bases[i:i + 1])).double()])
            assert o.shape == (2,)
            probs.append(float(o[1]))
    np.savez_compressed(os.path.join(HERE, 'notrain_small2.npz'), probs=np.array(probs), B=np.array(B))
    print('notrain: probs', np.array(probs))


def multinomial_fixture():
    """The real torch.multinomial CPU path vs the (u > cum0) rule with fp64 uniforms from the same seed."""
    res = {}
    for j, (B, C, seed) in enumerate([(16, 512, 7), (5, 768, 11), (64, 64, 13)]):
        rs = np.random.RandomState(seed)
        p = rs.random_sample((B, 2)).astype(np.float32)
        p[0] = [0.0, 1.0] if B > 0 else p[0]
        p[1] = [1.0, 0.0]
        p = p / p.sum(-1, keepdims=True, dtype=np.float32)
        torch.manual_seed(seed)
        idx = torch.multinomial(torch.from_numpy(p), num_samples=C, replacement=True).numpy()
        torch.manual_seed(seed)
        u = torch.rand(B * C, dtype=torch.float64).numpy().reshape(B, C)
        res[f'p{j}'] = p
        res[f'u{j}'] = u
        res[f'idx{j}'] = np.packbits(idx.astype(np.uint8), axis=1)
        res[f'C{j}'] = np.array(C)
    np.savez_compressed(os.path.join(HERE, 'multinomial_rule.npz'), **res)


def metrics_fixture(M):
    from BIOINF_tesi.models.utils.utils import AUPRC, F1_precision_recall, get_loss_weights_from_labels
    rs = np.random.RandomState(5)
    rows = []
    for n in range(400):
        B = int(rs.randint(1, 40))
        y = (rs.random_sample(B) < rs.choice([0.0, 0.1, 0.5, 1.0])).astype(np.int64)
        logits = rs.standard_normal((B, 2))
        if n % 7 == 0:
            logits[:, 1] = logits[:, 0]            # ties -> argmax picks class 0
        if n % 11 == 0:
            logits[:, 1] -= 10                     # nothing predicted positive
        tp, fp, fn, tn = O.confusion_counts(logits, y)
        a = AUPRC(torch.from_numpy(logits), torch.from_numpy(y.reshape(-1, 1)))
        f = F1_precision_recall(torch.from_numpy(logits), torch.from_numpy(y.reshape(-1, 1)))
        wp, wn = get_loss_weights_from_labels(torch.from_numpy(y.reshape(-1, 1)))
        rows.append([tp, fp, fn, tn, a, f[0], f[1], f[2], wp, wn])
    np.savez_compressed(os.path.join(HERE, 'metrics.npz'), rows=np.array(rows, dtype=np.float64))


def fit_fixture(M):
    """fit_multimodal (training_models_multimodal.py:40-226) on list loaders, 3 epochs, replayed draws."""
    from BIOINF_tesi.models.utils.training_models_multimodal import fit_multimodal
    case = CASES['small2']
    spec = case['spec']
    P0 = O.init_params(spec, 777)
    nb_train, nb_test, Btr, Bte = 3, 2, 8, 6
    train = {'FFNN': [], 'CNN': []}
    test = {'FFNN': [], 'CNN': []}
    queue = []
    epochs = 3
    batches = []
    for b in range(nb_train + nb_test):
        B = Btr if b < nb_train else Bte
        batches.append(make_inputs(spec, B, 800 + b))
    for b, (xf, bs, y) in enumerate(batches):
        dst = train if b < nb_train else test
        dst['FFNN'].append((torch.from_numpy(xf), torch.from_numpy(y.reshape(-1, 1))))
        dst['CNN'].append((torch.from_numpy(O.onehot_from_bases(bs)), torch.from_numpy(y.reshape(-1, 1))))
    for ep in range(epochs):
        for b in range(nb_train):
            queue += RH.draws_to_queue(spec, O.make_draws(spec, Btr, 10000 + ep * 100 + b), training=True)
        for b in range(nb_test):
            queue += RH.draws_to_queue(spec, O.make_draws(spec, Bte, 20000 + ep * 100 + b), training=False)
    model = RH.build_reference_model(M, spec, P0)
    opt = torch.optim.Adam(model.parameters(), lr=5e-3, weight_decay=1e-3)
    with tempfile.TemporaryDirectory() as td, RH.replay_draws(queue):
        a_tr, a_te, f1 = fit_multimodal(model, train, test, 'cpu', 'A549', 'active_E_vs_inactive_E', optimizer=opt,
                                        num_epochs=epochs, patience=10, verbose=False,
                                        checkpoint_path=os.path.join(td, 'ck.pt'))
    assert not queue
    np.savez_compressed(os.path.join(HERE, 'fit_small2.npz'), auprc_train=np.array(a_tr), auprc_test=np.array(a_te),
                        f1pr_test=np.array(f1), final_logit_w=model.state_dict()['post.3.weight'].numpy())
    print('fit: train', a_tr, 'test', a_te)


if __name__ == '__main__':
    assert RH.reference_available(), 'needs /root/reference'
    M = RH.import_reference()
    torch.set_num_threads(4)
    if len(sys.argv) > 1 and sys.argv[1] == 'bench':       # only the benchmark-architecture cases (added later)
        for name, case in BENCH_CASES.items():
            run_case(M, name, case)
        sys.exit(0)
    for name, case in {**CASES, **BENCH_CASES}.items():
        run_case(M, name, case)
    notrain_fixture(M)
    multinomial_fixture()
    metrics_fixture(M)
    fit_fixture(M)
    tot = sum(os.path.getsize(os.path.join(HERE, f)) for f in os.listdir(HERE) if f.endswith('.npz'))
    print('total fixture bytes', tot)
