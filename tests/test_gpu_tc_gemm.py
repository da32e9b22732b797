"""The tcgen05/TMEM/TMA GEMM kernel (and the SIMT kernel it replaces) against numpy on bf16-rounded inputs,
for every GEMM-shaped op of the step: Linear fwd/dgrad/wgrad, Conv1d implicit GEMM fwd/dgrad/wgrad."""
import ctypes as C
import os

import numpy as np
import pytest

from oracle import embracenet_oracle as O

pytestmark = pytest.mark.gpu


def run(kind, backend, a, b, out_shape, M=0, N=0, K=0, B=0, L=0, Cin=0, Cout=0, taps=0):
    import torch
    from embrace_b200 import _native as N_
    lib = N_.lib()
    ta = torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).cuda()
    tb = torch.from_numpy(np.ascontiguousarray(b, dtype=np.float32)).cuda()
    out = torch.full(out_shape, float('nan'), dtype=torch.float32, device='cuda')
    N_.check(lib.emb_k_gemm(kind, backend, C.c_void_p(ta.data_ptr()), C.c_void_p(tb.data_ptr()), C.c_void_p(out.data_ptr()),
                            M, N, K, B, L, Cin, Cout, taps, C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    return out.cpu().numpy().astype(np.float64)


def q(x):
    return O.bf16_round(x)


def check(got, ref, what):
    err = np.abs(got - ref).max() / max(np.abs(ref).max(), 1e-30)
    assert np.isfinite(got).all(), what
    assert err < 2e-3, (what, err)
    return err


LINEAR_SHAPES = [(128, 64, 64), (256, 128, 192), (200, 96, 72), (77, 32, 568), (1024, 256, 1024), (130, 1000, 64), (512, 16, 32)]


@pytest.mark.parametrize('backend', [0, 1])
@pytest.mark.parametrize('M,N,K', LINEAR_SHAPES)
def test_linear_fwd(backend, M, N, K):
    rs = np.random.RandomState(M + N + K)
    a, b = rs.standard_normal((M, K)), rs.standard_normal((N, K))
    check(run(0, backend, a, b, (M, N), M=M, N=N, K=K), q(a) @ q(b).T, ('fwd', backend, M, N, K))


@pytest.mark.parametrize('backend', [0, 1])
@pytest.mark.parametrize('M,N,K', [(128, 64, 64), (256, 128, 192), (200, 96, 72), (77, 568, 32), (300, 32, 128), (1024, 1024, 256), (64, 4096, 512)])
def test_linear_dgrad(backend, M, N, K):
    rs = np.random.RandomState(M + N + K + 1)
    a, b = rs.standard_normal((M, K)), rs.standard_normal((K, N))
    check(run(1, backend, a, b, (M, N), M=M, N=N, K=K), q(a) @ q(b), ('dgrad', backend, M, N, K))


@pytest.mark.parametrize('backend', [0, 1])
@pytest.mark.parametrize('M,N,K', [(128, 64, 128), (256, 128, 1000), (96, 72, 333), (32, 568, 77), (512, 1024, 4096), (1024, 64, 256), (16, 32, 64)])
def test_linear_wgrad(backend, M, N, K):
    rs = np.random.RandomState(M + N + K + 2)
    a, b = rs.standard_normal((K, M)), rs.standard_normal((K, N))
    check(run(2, backend, a, b, (M, N), M=M, N=N, K=K), q(a).T @ q(b), ('wgrad', backend, M, N, K))


def conv_ref(x_blc, W):
    """x [B,L,Cin] channels-last, W [Cout,Cin,k] -> [B,L,Cout]"""
    y = O.conv1d_fwd(np.transpose(x_blc, (0, 2, 1)), W, np.zeros(W.shape[0]))
    return np.transpose(y, (0, 2, 1))


CONV_SHAPES = [(3, 124, 64, 96, 15), (7, 58, 96, 256, 15), (11, 25, 256, 512, 5), (5, 124, 32, 32, 11), (4, 58, 16, 64, 5), (9, 25, 64, 128, 11),
               (6, 25, 256, 512, 15), (10, 8, 128, 256, 15), (5, 124, 16, 32, 15)]


@pytest.mark.parametrize('backend', [0, 1])
@pytest.mark.parametrize('B,L,Cin,Cout,k', CONV_SHAPES)
def test_conv_fwd(backend, B, L, Cin, Cout, k):
    rs = np.random.RandomState(B + L + Cin + Cout + k)
    x, W = rs.standard_normal((B, L, Cin)), rs.standard_normal((Cout, Cin, k))
    got = run(3, backend, x, W, (B * L, Cout), B=B, L=L, Cin=Cin, Cout=Cout, taps=k)
    check(got.reshape(B, L, Cout), conv_ref(q(x), q(W)), ('conv fwd', backend, B, L, Cin, Cout, k))


@pytest.mark.parametrize('backend', [0, 1])
@pytest.mark.parametrize('B,L,Cin,Cout,k', CONV_SHAPES)
def test_conv_dgrad(backend, B, L, Cin, Cout, k):
    rs = np.random.RandomState(B + L + Cin + Cout + k + 1)
    g, W = rs.standard_normal((B, L, Cout)), rs.standard_normal((Cout, Cin, k))
    x0 = np.zeros((B, Cin, L))
    dx, _, _ = O.conv1d_bwd(x0, q(W), np.transpose(q(g), (0, 2, 1)))
    got = run(4, backend, g, W, (B * L, Cin), B=B, L=L, Cin=Cin, Cout=Cout, taps=k)
    check(got.reshape(B, L, Cin), np.transpose(dx, (0, 2, 1)), ('conv dgrad', backend, B, L, Cin, Cout, k))


@pytest.mark.parametrize('backend', [0, 1])
@pytest.mark.parametrize('B,L,Cin,Cout,k', CONV_SHAPES + [(300, 25, 64, 128, 5)])
def test_conv_wgrad(backend, B, L, Cin, Cout, k):
    rs = np.random.RandomState(B + L + Cin + Cout + k + 2)
    g, x = rs.standard_normal((B, L, Cout)), rs.standard_normal((B, L, Cin))
    _, dW, _ = O.conv1d_bwd(np.transpose(q(x), (0, 2, 1)), np.zeros((Cout, Cin, k)), np.transpose(q(g), (0, 2, 1)))
    got = run(5, backend, g, x, (Cout, Cin, k), B=B, L=L, Cin=Cin, Cout=Cout, taps=k)
    check(got, dW, ('conv wgrad', backend, B, L, Cin, Cout, k))


@pytest.mark.parametrize('B,C1,k', [(1, 64, 15), (3, 64, 15), (37, 64, 11), (300, 64, 15), (5, 32, 5), (9, 16, 11), (150, 8, 15), (8, 24, 1)])
def test_onehot_conv_wgrad_tc(B, C1, k):
    """First-layer weight gradient as a tensor-core contraction against an in-smem one-hot Toeplitz operand
    (csrc/onehot_wgrad_tc.cuh) vs the oracle's histogram form."""
    import torch
    from embrace_b200 import _native as N_
    lib = N_.lib()
    rs = np.random.RandomState(B * 131 + C1 + k)
    bases = rs.randint(0, 4, size=(B, 256)).astype(np.uint8)
    bases[0, :8] = [0, 1, 2, 3, 3, 2, 1, 0]
    g = q(rs.standard_normal((B, C1, 256)))                                    # oracle layout [B, C, L]
    ref, _ = O.onehot_conv_bwd(bases, g, k)
    dy = torch.from_numpy(np.ascontiguousarray(g.transpose(0, 2, 1))).to(torch.bfloat16).cuda()       # channels-last
    tb = torch.from_numpy(bases).cuda()
    dw = torch.full((C1, 4, k), float('nan'), dtype=torch.float32, device='cuda')
    for _ in range(2):                                                         # twice: the call zeroes its output itself
        N_.check(lib.emb_k_onehot_conv_wgrad_tc(C.c_void_p(tb.data_ptr()), C.c_void_p(dy.data_ptr()), B, C1, k, C.c_void_p(dw.data_ptr()),
                                                C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    torch.cuda.synchronize()
    got = dw.cpu().numpy().astype(np.float64)
    assert np.isfinite(got).all()
    err = np.abs(got - ref).max() / max(np.abs(ref).max(), 1e-30)
    assert err < 1e-5, err                                                     # exact products, fp32 accumulation


@pytest.mark.parametrize('B,C1,k', [(1, 64, 15), (3, 64, 15), (37, 64, 11), (700, 64, 15), (5, 32, 5), (9, 16, 11), (150, 8, 15), (8, 24, 1)])
def test_onehot_conv_fwd_tc(B, C1, k):
    """K1 forward through the in-smem one-hot Toeplitz operand with exactly split (hi/mid/lo bf16) weights vs the oracle's
    gather-sum of the fp32 weights; BatchNorm statistics of the bf16-rounded outputs."""
    import torch
    from embrace_b200 import _native as N_
    lib = N_.lib()
    rs = np.random.RandomState(B * 17 + C1 + k)
    bases = rs.randint(0, 4, size=(B, 256)).astype(np.uint8)
    W = rs.standard_normal((C1, 4, k)).astype(np.float32)
    bias = rs.standard_normal(C1).astype(np.float32)
    ref = O.onehot_conv_fwd(bases, W.astype(np.float64), bias.astype(np.float64)).transpose(0, 2, 1)      # [B, L, C]
    tb, tw, tbias = torch.from_numpy(bases).cuda(), torch.from_numpy(W).cuda(), torch.from_numpy(bias).cuda()
    y = torch.full((B, 256, C1), float('nan'), dtype=torch.bfloat16, device='cuda')
    stats = torch.zeros(2, C1, dtype=torch.float64, device='cuda')
    N_.check(lib.emb_k_onehot_conv_fwd_tc(C.c_void_p(tb.data_ptr()), C.c_void_p(tw.data_ptr()), C.c_void_p(tbias.data_ptr()), B, C1, k,
                                          C.c_void_p(y.data_ptr()), C.c_void_p(stats.data_ptr()), C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    torch.cuda.synchronize()
    got = y.float().cpu().numpy().astype(np.float64)
    assert np.isfinite(got).all()
    # the fp32 sum is exact to ~1e-5 of the weight scale; the output is then rounded to bf16 once
    assert np.abs(got - ref).max() <= 2.0 ** -8 * np.abs(ref).max() + 1e-6
    near = np.abs(got - q(ref)) > 0                     # differs from the correctly rounded value only at rounding ties / 1-ulp flips
    assert near.mean() < 5e-4, near.mean()
    st = stats.cpu().numpy()
    np.testing.assert_allclose(st[0], got.sum(axis=(0, 1)), rtol=1e-5, atol=1e-3)
    np.testing.assert_allclose(st[1], (got * got).sum(axis=(0, 1)), rtol=1e-5, atol=1e-3)


@pytest.mark.parametrize('B,L,Cin,Cout,k', [(64, 124, 64, 96, 15), (16, 58, 64, 128, 11), (9, 25, 64, 64, 5), (1024, 124, 64, 96, 15), (300, 124, 64, 32, 5)])
def test_conv_wgrad_fused_taps(B, L, Cin, Cout, k):
    """Multi-tap wgrad with up to four taps per tcgen05.mma (EMB_WGRAD_FUSE_TAPS: the 64-wide N blocks of one instruction are
    the same staged tile one row apart, LBO = 128 bytes) must equal the one-MMA-per-tap form."""
    rs = np.random.RandomState(B + L + Cin + Cout + k + 7)
    g, x = rs.standard_normal((B, L, Cout)), rs.standard_normal((B, L, Cin))
    _, dW, _ = O.conv1d_bwd(np.transpose(q(x), (0, 2, 1)), np.zeros((Cout, Cin, k)), np.transpose(q(g), (0, 2, 1)))
    from embrace_b200 import _native as N
    res = {}
    for fuse in (1, 0):
        N.set_option('wgrad_fuse_taps', fuse)
        try:
            res[fuse] = run(5, 1, g, x, (Cout, Cin, k), B=B, L=L, Cin=Cin, Cout=Cout, taps=k)
        finally:
            N.set_option('wgrad_fuse_taps', 1)
        check(res[fuse], dW, ('conv wgrad, fused taps' if fuse else 'conv wgrad, one MMA per tap', B, L, Cin, Cout, k))
    assert np.abs(res[1] - res[0]).max() <= 1e-4 * np.abs(dW).max()
