"""The tensor-core GEMM kernels in the regime the benchmark runs them in: MANY tiles per launch (> 148 and > 296, i.e. more
than one and more than two tiles per persistent CTA), so the multi-tile loop with its TMEM double-buffer phase flips, the
conv3 `whole samples per 128-row tile` packing at B >= 2048, ragged last tiles and split-K tails are all exercised.
(tests/test_gpu_tc_gemm.py covers small shapes on both back ends.)

Reference: torch CPU fp32 library calls (nn.Linear / nn.Conv1d forward and their autograd formulas -- the ops the reference
model is made of, FFNN_pre.py:33, CNN_pre.py:39) on the SAME bf16-rounded inputs; tolerance 2e-4 of the output maximum:
bf16 x bf16 products are exact in fp32, so only the accumulation (order, and the tensor core's internal alignment of the
addends) separates the two -- which also shows that whatever distance the bf16 train step keeps from the exact-arithmetic
bf16-storage emulation (tests/test_gpu_bench_parity.py) does not come from the GEMMs."""
import ctypes as C

import numpy as np
import pytest

from oracle import embracenet_oracle as O

pytestmark = pytest.mark.gpu


def q32(x):
    return O.bf16_round(x).astype(np.float32)


def run_tc(kind, a, b, out_shape, **dims):
    import torch
    from embrace_b200 import _native as N_
    lib = N_.lib()
    ta, tb = torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).cuda(), torch.from_numpy(np.ascontiguousarray(b, dtype=np.float32)).cuda()
    out = torch.full(out_shape, float('nan'), dtype=torch.float32, device='cuda')
    d = dict(M=0, N=0, K=0, B=0, L=0, Cin=0, Cout=0, taps=0)
    d.update(dims)
    N_.check(lib.emb_k_gemm(kind, 1, C.c_void_p(ta.data_ptr()), C.c_void_p(tb.data_ptr()), C.c_void_p(out.data_ptr()),
                            d['M'], d['N'], d['K'], d['B'], d['L'], d['Cin'], d['Cout'], d['taps'], C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    return out.cpu().numpy()


def check(got, ref, what, tol=2e-4):
    assert np.isfinite(got).all(), what
    err = np.abs(got.astype(np.float64) - ref.astype(np.float64)).max() / max(np.abs(ref).max(), 1e-30)
    ERRS.append((str(what), float(err)))
    assert err < tol, (what, err)


ERRS = []


def teardown_module(module):
    import json
    import os
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'gpurun_out')
    os.makedirs(out, exist_ok=True)
    with open(os.path.join(out, 'big_tiles_errors.json'), 'w') as f:
        json.dump(ERRS, f, indent=1)


def tiles(M, N, n_cap=256):
    nt = min(n_cap, -(-N // 16) * 16)
    return -(-M // 128) * -(-N // nt)


@pytest.mark.parametrize('M,N,K', [(8192, 1024, 512), (20000, 512, 256), (8192, 1024, 4096), (2048, 1024, 4096), (8192, 32, 1024)])
def test_linear_fwd_many_tiles(M, N, K):
    import torch
    rs = np.random.RandomState(M + N + K)
    a, b = q32(rs.standard_normal((M, K))), q32(rs.standard_normal((N, K)))
    ref = (torch.from_numpy(a) @ torch.from_numpy(b).T).numpy()
    check(run_tc(0, a, b, (M, N), M=M, N=N, K=K), ref, ('fwd', M, N, K, tiles(M, N)))


@pytest.mark.parametrize('M,N,K', [(8192, 4096, 1024), (8192, 1024, 512), (19000, 256, 512), (2048, 4096, 1024)])
def test_linear_dgrad_many_tiles(M, N, K):
    import torch
    rs = np.random.RandomState(M + N + K + 1)
    a, b = q32(rs.standard_normal((M, K))), q32(rs.standard_normal((K, N)))
    ref = (torch.from_numpy(a) @ torch.from_numpy(b)).numpy()
    assert tiles(M, N) > 148
    check(run_tc(1, a, b, (M, N), M=M, N=N, K=K), ref, ('dgrad', M, N, K, tiles(M, N)))


@pytest.mark.parametrize('M,N,K', [(1024, 4096, 8192), (512, 1024, 8200), (1024, 4096, 2048), (256, 512, 8192), (32, 64, 8192)])
def test_linear_wgrad_many_tiles_and_split_k(M, N, K):
    import torch
    rs = np.random.RandomState(M + N + K + 2)
    a, b = q32(rs.standard_normal((K, M))), q32(rs.standard_normal((K, N)))
    ref = (torch.from_numpy(a).T @ torch.from_numpy(b)).numpy()
    check(run_tc(2, a, b, (M, N), M=M, N=N, K=K), ref, ('wgrad', M, N, K, tiles(M, N, 128)))


# (B, L, Cin, Cout, k): the benchmark architecture's conv1 / conv2 / conv3 at large batch, plus arch S / M layers
CONV_BIG = [(512, 124, 64, 96, 15), (1024, 58, 96, 256, 15), (2048, 25, 256, 512, 15), (2050, 25, 256, 512, 15),
            (1100, 58, 32, 128, 11), (2048, 58, 96, 64, 5), (4096, 8, 128, 128, 5)]


def _conv_torch(x_blc, W, g_blc=None):
    """x [B,L,Cin], W [Cout,Cin,k] (fp32, already bf16-rounded).  Returns y [B,L,Cout], or (dx [B,L,Cin], dW) for g."""
    import torch
    import torch.nn.functional as F
    torch.set_num_threads(max(torch.get_num_threads(), 8))
    x = torch.from_numpy(np.ascontiguousarray(x_blc.transpose(0, 2, 1))).requires_grad_(g_blc is not None)
    w = torch.from_numpy(W).requires_grad_(g_blc is not None)
    y = F.conv1d(x, w, None, stride=1, padding=(W.shape[2] - 1) // 2)
    if g_blc is None:
        return y.detach().numpy().transpose(0, 2, 1)
    y.backward(torch.from_numpy(np.ascontiguousarray(g_blc.transpose(0, 2, 1))))
    return x.grad.numpy().transpose(0, 2, 1), w.grad.numpy()


@pytest.mark.parametrize('B,L,Cin,Cout,k', CONV_BIG)
def test_conv_fwd_many_tiles(B, L, Cin, Cout, k):
    rs = np.random.RandomState(B + L + Cin + Cout + k)
    x, W = q32(rs.standard_normal((B, L, Cin))), q32(rs.standard_normal((Cout, Cin, k)))
    got = run_tc(3, x, W, (B * L, Cout), B=B, L=L, Cin=Cin, Cout=Cout, taps=k)
    check(got.reshape(B, L, Cout), _conv_torch(x, W), ('conv fwd', B, L, Cin, Cout, k))


@pytest.mark.parametrize('B,L,Cin,Cout,k', CONV_BIG)
def test_conv_dgrad_and_wgrad_many_tiles(B, L, Cin, Cout, k):
    rs = np.random.RandomState(B + L + Cin + Cout + k + 1)
    x, W = q32(rs.standard_normal((B, L, Cin))), q32(rs.standard_normal((Cout, Cin, k)))
    g = q32(rs.standard_normal((B, L, Cout)))
    dx, dW = _conv_torch(x, W, g)
    got = run_tc(4, g, W, (B * L, Cin), B=B, L=L, Cin=Cin, Cout=Cout, taps=k)
    check(got.reshape(B, L, Cin), dx, ('conv dgrad', B, L, Cin, Cout, k))
    got = run_tc(5, g, x, (Cout, Cin, k), B=B, L=L, Cin=Cin, Cout=Cout, taps=k)
    check(got, dW, ('conv wgrad', B, L, Cin, Cout, k))
