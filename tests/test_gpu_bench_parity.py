"""Parity of the BENCHMARKED configuration at its own size (BASELINE configs[2]: arch L, batch 8192, bf16 tensor-core path).

One full train step with replayed draws -- forward, weighted CE, backward, Adam -- through the C ABI against
oracle/torch_port.py, i.e. the reference's own PyTorch-CPU library calls in fp64 (EmbraceNetMultimodal.py:159-193,
training_models_multimodal.py:132-162).  At these sizes every big GEMM has far more than 148 tiles, so the persistent
multi-tile loops, the conv3 packing, split-K tails and the pooling kernels' grids run in the regime bench.py times.

Two references, both test infrastructure (oracle/torch_port.py):
  fp64     the reference's arithmetic (model.double(), training_models_multimodal.py:115)
  bf16emu  the same library calls with values rounded to bfloat16 at exactly the points where EMB_PREC_BF16 STORES a tensor
           (activations, activation gradients, GEMM weight operands) and wide arithmetic everywhere else -- "the engine with
           exact arithmetic".  It is pinned to the numpy oracle's own emulation on the CPU (test_oracle_golden.py).
What separates bf16emu from fp64 is the precision contract itself, not the kernels; what separates the engine from bf16emu
is fp32 instead of exact accumulation, which the contract amplifies (tests/test_bf16_contract_cpu.py measures that
sensitivity on the CPU: a 1e-6 relative perturbation of the stored activations moves the emulation's own logits by 3e-3 and
its CNN gradients by 5 %).  The report states all three distances side by side; kernel bugs in the many-tile regime are
caught by the fp32 engine against fp64 at batch 2048 (2e-6), by tests/test_gpu_big_tiles.py (GEMMs alone, 2e-4) and by
the bit-exact selection.  Finding (r2, arch L and S, batch 2048..8192, at initialisation): bf16 storage alone moves the logits by 7e-3 (L2) and
the parameter gradients by 0.5 % (head) to 20 % (first conv layers) relative L2 against fp64, independent of batch size and
of whether the labels carry signal -- the numpy emulation shows the same numbers at batch 256..1024; the gradient of a
freshly initialised network is a difference of near-equal class means, and every stored tensor carries 2^-9 relative rounding.

What is asserted (measured values are written to gpurun_out/parity_bench_<precision>_<B>.json):
  selection indices          bit-exact
  logits, loss               max-norm / relative L2 / relative
  every parameter gradient   relative L2 error, max-norm error and cosine per tensor (conv biases behind BatchNorm are
                             analytically zero: compared in absolute terms against the weight-gradient scale)
  BatchNorm running stats    max-norm
  post-Adam parameters       the engine's update vs the fp64 Adam rule applied to the ENGINE's gradients (Adam normalises per
                             element, so gradient noise must be kept out of this check)
"""
import json
import os

import numpy as np
import pytest

from oracle import embracenet_oracle as O
from tests.golden.cases import ARCH_L, ARCH_S, make_inputs
from tests.test_gpu_parity import to_archspec, nerr, l2err

pytestmark = pytest.mark.gpu
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'gpurun_out')

# engine precision -> reference -> bounds
TOL = {
    'fp32': {'fp64': dict(logits_max=2e-4, logits_l2=1e-4, loss=2e-5, grad_l2=3e-3, grad_max=1e-2, cos=0.99999, bn=1e-5)},
    # bf16emu bounds = about 3x the emulation's own sensitivity to a 1e-6 perturbation of what it stores (tests/test_bf16_contract_cpu.py)
    'bf16': {'bf16emu': dict(logits_max=1.5e-2, logits_l2=8e-3, loss=1e-3, grad_l2=0.15, grad_max=0.25, cos=0.985, bn=1e-4),
             'fp64': dict(logits_max=3e-2, logits_l2=1.2e-2, loss=2e-3, grad_l2=0.30, grad_max=0.40, cos=0.95, bn=4e-3)},
}


def cosine(a, b):
    a, b = np.asarray(a, np.float64).ravel(), np.asarray(b, np.float64).ravel()
    return float(a @ b / max(np.linalg.norm(a) * np.linalg.norm(b), 1e-300))


def reference_step(spec, P, x, bases, y, draws, lr, wd, emulate):
    import torch
    from oracle import torch_port as TP
    st = TP.TrainState(spec, {k: v.copy() for k, v in P.items()}, 'adam', lr=lr, wd=wd, emulate_bf16=emulate)
    loss, logits, idx = st.step(torch.from_numpy(x), torch.from_numpy(O.onehot_from_bases(bases)), y, draws)
    return dict(loss=float(loss), logits=logits.numpy(), idx=idx.numpy(),
                grads={k: v.grad.detach().numpy() for k, v in st.T.items() if not O.is_buffer(k)},
                bufs={k: v.detach().numpy() for k, v in st.T.items() if k.endswith(('running_mean', 'running_var'))})


def compare(got_logits, got_loss, grads, got_P, ref):
    rep = dict(logits_max=nerr(got_logits, ref['logits']), logits_l2=l2err(got_logits, ref['logits']),
               loss=abs(got_loss - ref['loss']) / max(1.0, abs(ref['loss'])), grads={})
    worst_l2 = worst_max = 0.0
    worst_cos = 1.0
    for k, gr in ref['grads'].items():
        wk = k[:-4] + 'weight'
        if k.endswith('.bias') and ref['grads'][wk].ndim == 3:
            rep['grads'][k] = dict(abs=float(np.abs(grads[k] - gr).max()), wscale=float(np.abs(ref['grads'][wk]).max()))
            continue
        e2, em, c = l2err(grads[k], gr), nerr(grads[k], gr), cosine(grads[k], gr)
        rep['grads'][k] = dict(l2=e2, max=em, cos=c)
        worst_l2, worst_max, worst_cos = max(worst_l2, e2), max(worst_max, em), min(worst_cos, c)
    rep['grad_l2_worst'], rep['grad_max_worst'], rep['grad_cos_worst'] = worst_l2, worst_max, worst_cos
    rep['bn'] = max(nerr(got_P[k], v) for k, v in ref['bufs'].items()) if ref['bufs'] else 0.0
    return rep


def big_step(spec, B, precision, tensor_core, seed=700, lr=1e-3, wd=1e-3, tag=''):
    import torch
    from embrace_b200 import Engine
    P = O.init_params(spec, seed)
    P = {k: (v.astype(np.float32).astype(np.float64) if v.dtype == np.float64 else v) for k, v in P.items()}   # fp32-representable weights
    x, bases, y = make_inputs(spec, B, seed + 1)
    draws = O.make_draws(spec, B, seed + 2, force_modal=True)        # modality dropout on: rows keep one random modality
    torch.set_num_threads(os.cpu_count() or 1)
    refs = {'fp64': reference_step(spec, P, x, bases, y, draws, lr, wd, False)}
    if precision == 'bf16':
        refs['bf16emu'] = reference_step(spec, P, x, bases, y, draws, lr, wd, True)

    eng = Engine(to_archspec(spec), max_batch=B, precision=precision, tensor_core=tensor_core)
    eng.load_numpy(P)
    eng.metrics_reset()
    logits = eng.forward(torch.from_numpy(x.astype(np.float32)), torch.from_numpy(bases), training=True, draws=draws)
    dlogits = eng.loss(logits, torch.from_numpy(y))
    eng.backward(dlogits)
    got_logits = logits.cpu().numpy()
    grads = eng.grads_numpy()
    idx = eng.last_selection(B).cpu().numpy()
    m = eng.metrics_read()
    # optimizer: the engine's update against the fp64 rule on the engine's own gradients
    P_exp = {k: v.copy() for k, v in P.items()}
    O.opt_step(P_exp, grads, O.opt_init(P, 'adam'), lr, wd)
    eng.opt_step(eng.opt_config('adam', lr=lr, weight_decay=wd))
    got_P = eng.params_numpy()

    report = dict(B=B, precision=precision, tensor_core=bool(tensor_core), launches=int(eng.launch_count),
                  idx_mismatches=int((idx != refs['fp64']['idx']).sum()), vs={})
    for name, ref in refs.items():
        report['vs'][name] = compare(got_logits, m[0]['loss'], grads, got_P, ref)
    if 'bf16emu' in refs:       # the precision contract's own distance to the reference arithmetic, kernels not involved
        e, f = refs['bf16emu'], refs['fp64']
        report['bf16emu_vs_fp64'] = {k: v for k, v in compare(e['logits'], e['loss'], e['grads'], {**P, **e['bufs']}, f).items() if k != 'grads'}
    os.makedirs(OUT, exist_ok=True)
    with open(os.path.join(OUT, f'parity_bench_{tag or precision}_{B}.json'), 'w') as fh:
        json.dump(report, fh, indent=1)
    print(json.dumps({k: ({n: {a: b for a, b in r.items() if a != 'grads'} for n, r in v.items()} if k == 'vs' else v) for k, v in report.items()}))

    assert report['idx_mismatches'] == 0, 'modality selection must be bit-exact'
    assert np.array_equal(refs['fp64']['idx'], refs.get('bf16emu', refs['fp64'])['idx'])
    assert np.isfinite(got_logits).all()
    for name, rep in report['vs'].items():
        tol = TOL[precision][name]
        assert rep['logits_max'] <= tol['logits_max'] and rep['logits_l2'] <= tol['logits_l2'], (name, rep['logits_max'], rep['logits_l2'])
        assert rep['loss'] <= tol['loss'], (name, rep['loss'])
        for k, r in rep['grads'].items():
            if 'abs' in r:
                assert r['abs'] <= (1e-4 if precision == 'fp32' else 5e-2) * max(r['wscale'], 1.0), (name, k, r)
            else:
                assert r['l2'] <= tol['grad_l2'] and r['max'] <= tol['grad_max'] and r['cos'] >= tol['cos'], (name, k, r)
        assert rep['bn'] <= tol['bn'], (name, rep['bn'])
    for k in refs['fp64']['grads']:
        amp = 1.2e-7 * max(np.abs(grads[k]).max(), wd * np.abs(P_exp[k]).max()) / 1e-8
        bound = 2e-6 * np.abs(P_exp[k]).max() + lr * min(1.0, 1e-4 + amp)
        assert np.abs(got_P[k] - P_exp[k]).max() <= bound, ('adam', k)
    return report


def test_arch_L_fp32_engine_matches_reference_at_batch_2048():
    """The verification (fp32 / SIMT) engine against the fp64 reference path at a batch far beyond the golden cases: pins the
    reference port and the engine to each other at scale before the bf16 comparison below is read."""
    big_step(ARCH_L, 2048, 'fp32', False)


@pytest.mark.parametrize('B', [2048, 8192])
def test_arch_L_bf16_tensor_core_matches_reference_at_benchmark_batch(B):
    """bench.py's workload itself: arch L, bf16, tcgen05 GEMMs, batch 8192 (and the 4-GPU shard size 2048)."""
    big_step(ARCH_L, B, 'bf16', True)


def test_arch_S_bf16_tensor_core_matches_reference_at_batch_8192():
    big_step(ARCH_S, 8192, 'bf16', True, tag='archS_bf16')


def test_adamw_matches_decoupled_rule():
    """EMB_OPT_ADAMW (north_star's fused AdamW; the reference itself runs coupled L2): two steps of the engine's kernel against
    O.opt_step(decoupled=True) driven by the engine's own gradients."""
    import torch
    from embrace_b200 import Engine
    spec, B = ARCH_S, 64
    P = O.init_params(spec, 5)
    P = {k: (v.astype(np.float32).astype(np.float64) if v.dtype == np.float64 else v) for k, v in P.items()}
    x, bases, y = make_inputs(spec, B, 6)
    eng = Engine(to_archspec(spec), max_batch=B, precision='fp32', tensor_core=False)
    eng.load_numpy(P)
    lr, wd = 3e-3, 5e-2
    cfg = eng.opt_config('adamw', lr=lr, weight_decay=wd)
    P_exp = {k: v.copy() for k, v in P.items()}
    st = O.opt_init(P_exp, 'adam')
    for step in range(2):
        draws = O.make_draws(spec, B, 900 + step, force_modal=False)
        logits = eng.forward(torch.from_numpy(x.astype(np.float32)), torch.from_numpy(bases), training=True, draws=draws)
        eng.backward(eng.loss(logits, torch.from_numpy(y)))
        grads = eng.grads_numpy()
        eng.opt_step(cfg)
        O.opt_step(P_exp, grads, st, lr, wd, decoupled=True)
        got = eng.params_numpy()
        for k in grads:
            amp = 1.2e-7 * np.abs(grads[k]).max() / 1e-8
            bound = 2e-6 * np.abs(P_exp[k]).max() + lr * min(1.0, 1e-4 + amp)
            assert np.abs(got[k] - P_exp[k]).max() <= bound, (step, k)
        # the decoupled rule must actually differ from coupled L2 at this weight decay
        if step == 0:
            P_c = {k: v.copy() for k, v in P.items()}
            O.opt_step(P_c, grads, O.opt_init(P_c, 'adam'), lr, wd, decoupled=False)
            assert max(np.abs(P_c[k] - P_exp[k]).max() for k in grads) > 10 * lr * 1e-4
        for k in P_exp:                       # keep the oracle's buffers in step with the engine's (BatchNorm running stats)
            if O.is_buffer(k) and k in got:
                P_exp[k] = got[k]


def test_training_curve_auprc_parity_bf16_tensor_core_large_batch():
    """AUPRC after training on the PRODUCTION path: bf16 storage + tcgen05 GEMMs at batch 1024, arch S, against the reference's
    fp64 CPU path (oracle/torch_port.py) with the same replayed draws -- 200 Adam steps at lr 3e-3 and 60 at 3e-4 on
    planted-signal data (logistic on 4 features + a motif), then ranking AUPRC (sklearn average_precision_score of z1 - z0)
    and the reference's hard-prediction AUPRC (utils.py:80-86) on 16 384 held-out rows.

    Bounds: 0.004 (ranking) / 0.008 (hard), next to north_star's 0.002.  Training is a chaotic map of its rounding: what can be
    asked is that the engine lands inside the reference's OWN sensitivity band.  Measured r2 (profiles/r02_training_curve.json,
    `python profiles/training_curve.py 260`, 8192 held-out rows), final ranking / hard AUPRC and distance to the reference:
        reference fp64                          0.98938 / 0.9237
        reference, initial weights * (1+1e-6)   0.99081 / 0.9262    +0.0014 / +0.0026   <- the band
        engine fp32 (SIMT)                      0.99058 / 0.9275    +0.0012 / +0.0038
        engine bf16 (tcgen05)                   0.98872 / 0.9217    -0.0007 / -0.0020
    while mid-training (step 60, the steep part of the curve) the four runs are spread over 0.62 .. 0.94.  0.002 itself is met
    where trajectories can be kept identical: the fp32 engine in deterministic mode on a converged task, tests/test_gpu_dropin.py
    (measured 0.0009 / 0.0012).  The engine runs in deterministic mode here so that the test's outcome is reproducible."""
    import torch
    from sklearn.metrics import average_precision_score
    from oracle import torch_port as TP
    from embrace_b200 import Engine
    spec = ARCH_S
    rs = np.random.RandomState(3)
    N, B, steps, decay_at = 8192, 1024, 260, 200
    F = spec['F']

    def data(n):
        x = rs.random_sample((n, F)).astype(np.float32).astype(np.float64)
        bases = rs.randint(0, 4, size=(n, 256)).astype(np.uint8)
        motif = np.array([0, 2, 2, 1, 3, 0, 3, 1], dtype=np.uint8)
        z = 8.0 * (x[:, :4].mean(1) - 0.5) * np.sqrt(48.0) - 1.0
        y = (rs.random_sample(n) < 1 / (1 + np.exp(-z))).astype(np.int64)
        for i in np.nonzero(y)[0]:
            if rs.random_sample() < 0.7:
                for _ in range(3):
                    pos = rs.randint(0, 248)
                    bases[i, pos:pos + 8] = motif
        return x, bases, y
    xtr, btr, ytr = data(N)
    xte, bte, yte = data(16384)
    P = O.init_params(spec, 17)
    P = {k: (v.astype(np.float32).astype(np.float64) if v.dtype == np.float64 else v) for k, v in P.items()}
    torch.set_num_threads(os.cpu_count() or 1)
    st = TP.TrainState(spec, {k: v.copy() for k, v in P.items()}, 'adam', lr=3e-3, wd=1e-4)
    from embrace_b200 import _native as NAT
    NAT.set_option('deterministic', 1)
    eng = Engine(to_archspec(spec), max_batch=16384, precision='bf16', tensor_core=True)
    eng.load_numpy(P)
    for s in range(steps):
        lr = 3e-3 if s < decay_at else 3e-4
        for g in st.opt.param_groups:
            g['lr'] = lr
        cfg = eng.opt_config('adam', lr=lr, weight_decay=1e-4)
        lo = (s * B) % N
        xb, bb, yb = xtr[lo:lo + B], btr[lo:lo + B], ytr[lo:lo + B]
        draws = O.make_draws(spec, B, 5000 + s)
        st.step(torch.from_numpy(xb), torch.from_numpy(O.onehot_from_bases(bb)), yb, draws)
        eng.train_step(torch.from_numpy(xb.astype(np.float32)), torch.from_numpy(bb), torch.from_numpy(yb), cfg, draws=draws)
    NAT.set_option('deterministic', 0)
    u = np.random.RandomState(9).random_sample((len(yte), spec['C']))
    with torch.no_grad():
        Tt = {k: v.detach() for k, v in st.T.items()}
        ref_logits, _ = TP.forward(spec, Tt, torch.from_numpy(xte), torch.from_numpy(O.onehot_from_bases(bte)), {'embrace_u': u}, training=False)
    ref_logits = ref_logits.numpy()
    got_logits = eng.forward(torch.from_numpy(xte.astype(np.float32)), torch.from_numpy(bte), training=False,
                             draws={'embrace_u': u, 'modal_u0': 0.0}).cpu().numpy()
    s_ref, s_got = ref_logits[:, 1] - ref_logits[:, 0], got_logits[:, 1] - got_logits[:, 0]
    rank_ref, rank_got = average_precision_score(yte, s_ref), average_precision_score(yte, s_got)
    hard_ref, hard_got = O.auprc_hard(ref_logits, yte), O.auprc_hard(got_logits, yte)
    rep = dict(rank_ref=rank_ref, rank_got=rank_got, hard_ref=hard_ref, hard_got=hard_got, base_rate=float(yte.mean()),
               flipped_predictions=int(((s_ref > 0) != (s_got > 0)).sum()), launches=int(eng.launch_count))
    os.makedirs(OUT, exist_ok=True)
    with open(os.path.join(OUT, 'auprc_parity_bf16_tc.json'), 'w') as f:
        json.dump(rep, f, indent=1)
    print(json.dumps(rep))
    assert rank_ref > yte.mean() + 0.3 and rank_got > yte.mean() + 0.3, 'both must have learned the planted signal'
    assert abs(rank_ref - rank_got) <= 0.004
    assert abs(hard_ref - hard_got) <= 0.008
