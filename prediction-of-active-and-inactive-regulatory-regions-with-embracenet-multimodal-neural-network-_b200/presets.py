"""Canonical architecture specs of the reference's search space (real best trials S and M of the A549
notebook, the largest point L, the widest docking W) and the synthetic batch generator the benchmarks use.  Product
code: nothing here touches oracle/ or tests/."""
import numpy as np

from .archspec import ArchSpec


def _e(F, ffnn, cnn, C, post, p_ffnn):
    return ArchSpec(kind='embracenet', in_features=F, ffnn_units=[u for u, _ in ffnn], ffnn_dropout=[p for _, p in ffnn],
                    cnn_channels=[c for c, _, _ in cnn], cnn_kernels=[k for _, k, _ in cnn], cnn_dropout=[p for _, _, p in cnn],
                    embracement_size=C, post_units=[u for u, _ in post], post_dropout=[p for _, p in post], p_ffnn=p_ffnn)


def arch(name):
    """'S' | 'M' | 'L' | 'W' -> a fresh ArchSpec."""
    if name == 'S':
        return _e(48, [(128, 0.4), (32, 0.3)], [(32, 5, 0.2), (96, 5, 0.4), (64, 5, 0.4)], 512, [(64, 0.0), (128, 0.2)], 0.5)
    if name == 'M':
        return _e(48, [(128, 0.2)], [(32, 11, 0.2), (32, 11, 0.4), (128, 11, 0.0), (128, 5, 0.5)], 512, [], 0.3)
    if name == 'L':
        return _e(562, [(256, 0.2), (128, 0.2), (64, 0.4), (32, 0.4)],
                  [(64, 15, 0.2), (96, 15, 0.4), (256, 15, 0.4), (512, 15, 0.4)], 1024, [(512, 0.2), (256, 0.2)], 0.5)
    if name == 'W':
        return _e(429, [(256, 0.2)], [(64, 15, 0.2)], 1024, [(512, 0.2), (256, 0.2)], 0.5)
    raise KeyError(name)


def fwd_flops_per_sample(spec: ArchSpec):
    """Dense-equivalent forward FLOPs (2*MAC) per sample of the GEMM-shaped ops; the one-hot layer counted as a gather."""
    f, fin = 0, spec.in_features
    for u in spec.ffnn_units:
        f += 2 * fin * u
        fin = u
    cin, L = 4, 256
    for i, ((co, k), (Lc, Lp)) in enumerate(zip(zip(spec.cnn_channels, spec.cnn_kernels), spec.cnn_lengths())):
        f += (k * co * Lc) if i == 0 else 2 * cin * k * co * Lc
        cin, L = co, Lp
    C = spec.embracement_size
    f += 2 * (spec.ffnn_units[-1] + cin * L) * C
    fin = C
    for u in spec.post_units:
        f += 2 * fin * u
        fin = u
    return f + 2 * fin * 2


def synthetic_batches(spec: ArchSpec, B, n, seed, base_rate_logit=-1.9):
    """n batches of (x [B,F] float32 in [0,1) -- the reference MinMax-scales its features --, bases [B,256] uint8 codes
    a,c,g,t = 0..3, labels [B] int32) with a planted signal: logistic on 4 features (base rate ~1/7)."""
    rs = np.random.RandomState(seed)
    out = []
    for _ in range(n):
        x = rs.random_sample((B, spec.in_features)).astype(np.float32)
        bases = rs.randint(0, 4, size=(B, 256)).astype(np.uint8)
        z = 3.0 * (x[:, :4].sum(1) - 2.0) + base_rate_logit
        y = (rs.random_sample(B) < 1 / (1 + np.exp(-z))).astype(np.int32)
        out.append((x, bases, y))
    return out
