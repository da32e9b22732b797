"""Golden-case definitions shared by make_golden.py (generator) and the tests (consumers)."""
import numpy as np


def espec(F, ffnn, cnn, C, post, p_ffnn):
    return dict(kind='embracenet', F=F,
                ffnn_units=[u for u, _ in ffnn], ffnn_dropout=[p for _, p in ffnn],
                cnn_channels=[c for c, _, _ in cnn], cnn_kernels=[k for _, k, _ in cnn],
                cnn_dropout=[p for _, _, p in cnn],
                C=C, post_units=[u for u, _ in post], post_dropout=[p for _, p in post], p_ffnn=p_ffnn)


# canonical arch specs of SURVEY.md section 8
ARCH_S = espec(48, [(128, 0.4), (32, 0.3)], [(32, 5, 0.2), (96, 5, 0.4), (64, 5, 0.4)], 512, [(64, 0.0), (128, 0.2)], 0.5)
ARCH_M = espec(48, [(128, 0.2)], [(32, 11, 0.2), (32, 11, 0.4), (128, 11, 0.0), (128, 5, 0.5)], 512, [], 0.3)
ARCH_L = espec(562, [(256, 0.2), (128, 0.2), (64, 0.4), (32, 0.4)],
               [(64, 15, 0.2), (96, 15, 0.4), (256, 15, 0.4), (512, 15, 0.4)], 1024, [(512, 0.2), (256, 0.2)], 0.5)
ARCH_W = espec(429, [(256, 0.2)], [(64, 15, 0.2)], 1024, [(512, 0.2), (256, 0.2)], 0.5)

CASES = {
    # name: spec, batch, seed, steps, force_modal per step (None = use the drawn coin)
    'tiny': dict(spec=espec(7, [(8, 0.3)], [(6, 5, 0.2)], 16, [], 0.35), B=5, seed=11, steps=2,
                 force_modal=[False, True]),
    'small2': dict(spec=espec(12, [(16, 0.2), (8, 0.0)], [(8, 5, 0.0), (12, 11, 0.4)], 24, [(10, 0.3)], 0.6),
                   B=6, seed=21, steps=2, force_modal=[True, False]),
    'deep4': dict(spec=espec(9, [(32, 0.4), (16, 0.3), (4, 0.5), (4, 0.0)],
                             [(16, 15, 0.3), (32, 11, 0.5), (64, 5, 0.0), (128, 5, 0.4)], 64,
                             [(32, 0.2), (16, 0.5)], 0.8), B=4, seed=31, steps=2, force_modal=[None, None]),
    'archS': dict(spec=ARCH_S, B=4, seed=41, steps=1, force_modal=[False], lr=4.1e-5, wd=7.6e-4),
    'tiny_rmsprop': dict(spec=espec(7, [(8, 0.3)], [(6, 5, 0.2)], 16, [], 0.35), B=5, seed=51, steps=2,
                         force_modal=[False, True], opt='rmsprop'),
    'tiny_nadam': dict(spec=espec(7, [(8, 0.3)], [(6, 5, 0.2)], 16, [], 0.35), B=5, seed=61, steps=2,
                       force_modal=[True, False], opt='nadam'),
    'ffnn_only': dict(spec=dict(kind='ffnn', F=10, ffnn_units=[16, 8, 4], ffnn_dropout=[0.2, 0.0, 0.5]),
                      B=7, seed=71, steps=2),
    'cnn_only': dict(spec=dict(kind='cnn', cnn_channels=[8, 16], cnn_kernels=[11, 5], cnn_dropout=[0.2, 0.4]),
                     B=3, seed=81, steps=1),
    # ConcatNetMultimodal: post(cat(FFNN, CNN)), 1..3 post layers
    'concat_small': dict(spec=dict(kind='concatnet', F=9, ffnn_units=[16, 8], ffnn_dropout=[0.2, 0.0], cnn_channels=[8, 16],
                                   cnn_kernels=[5, 11], cnn_dropout=[0.0, 0.4], post_units=[24, 16, 8], post_dropout=[0.3, 0.0, 0.5]),
                         B=6, seed=95, steps=2),
}

# The benchmark architectures themselves (L = largest point of the search space = bench.py's workload, W = widest
# docking, M = the 4-conv-layer real trial).  Pinned on the CPU only (test_oracle_golden.py); the GPU suites compare the
# engine with the oracle on these shapes (test_benchmark_archs_tensor_core_full_step).
BENCH_CASES = {
    'archL': dict(spec=ARCH_L, B=3, seed=201, steps=1, force_modal=[False], lr=1e-3, wd=1e-3),
    'archW': dict(spec=ARCH_W, B=4, seed=211, steps=1, force_modal=[True], lr=1e-3, wd=1e-3),
    'archM': dict(spec=ARCH_M, B=4, seed=221, steps=2, force_modal=[None, False], lr=1e-3, wd=1e-3),
}


def make_inputs(spec, B, seed):
    """x_ffnn fp64 holding fp32-representable U[0,1) values (features are MinMax-scaled,
    dataprepare.py:88), bases uint8 codes 0..3, labels int64 with >= 1 positive and >= 1 negative."""
    rs = np.random.RandomState(seed)
    F = spec.get('F', 4)
    x = rs.random_sample((B, F)).astype(np.float32).astype(np.float64)
    bases = rs.randint(0, 4, size=(B, 256)).astype(np.uint8)
    y = (rs.random_sample(B) < 0.3).astype(np.int64)
    if B >= 2:
        y[0], y[1] = 1, 0
    return x, bases, y


BIG = 20_000
STRIDE = 97


def compress(out, key, arr):
    """Store small arrays whole; big ones (docking_1 of arch S) as a strided sample + sums."""
    arr = np.asarray(arr)
    if arr.size <= BIG:
        out[key] = arr.copy()
    else:
        flat = arr.reshape(-1)
        out[key + '#sample'] = flat[::STRIDE].copy()
        out[key + '#sums'] = np.array([flat.sum(), np.abs(flat).sum(), (flat * flat).sum()])


def check_against(npz, key, arr, rtol, atol, what=''):
    """Compare `arr` with what compress() stored under `key`."""
    arr = np.asarray(arr, dtype=np.float64)
    if key in npz:
        ref = np.asarray(npz[key], dtype=np.float64)
        assert ref.shape == arr.shape, (key, ref.shape, arr.shape)
        scale = max(np.abs(ref).max(), 1e-30) if ref.size else 1.0
        err = np.abs(arr - ref).max() if ref.size else 0.0
        assert err <= atol + rtol * scale, f'{what}{key}: max abs err {err:.3e} vs scale {scale:.3e}'
        return err / scale
    ref = np.asarray(npz[key + '#sample'], dtype=np.float64)
    flat = arr.reshape(-1)
    got = flat[::STRIDE]
    scale = max(np.abs(ref).max(), 1e-30)
    err = np.abs(got - ref).max()
    assert err <= atol + rtol * scale, f'{what}{key}#sample: max abs err {err:.3e} vs scale {scale:.3e}'
    sums = npz[key + '#sums']
    got_sums = np.array([flat.sum(), np.abs(flat).sum(), (flat * flat).sum()])
    assert np.allclose(got_sums[1:], sums[1:], rtol=max(rtol * 10, 1e-9)), (key, got_sums, sums)
    return err / scale
