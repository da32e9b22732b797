"""Per-shape timing of the tensor-core GEMM kernels (emb_k_gemm_time): python scratch/gemm_bench.py [set]"""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from embrace_b200 import _native as N_
lib = N_.lib()


def t(kind, M=0, N=0, K=0, B=0, L=0, Cin=0, Cout=0, taps=0, reps=6, label=''):
    if kind == 0: na, nb, no, fl = M * K, N * K, M * N, 2.0 * M * N * K
    elif kind == 1: na, nb, no, fl = M * K, K * N, M * N, 2.0 * M * N * K
    elif kind == 2: na, nb, no, fl = K * M, K * N, M * N, 2.0 * M * N * K
    elif kind == 3: na, nb, no, fl = B * L * Cin, Cout * Cin * taps, B * L * Cout, 2.0 * B * L * Cin * Cout * taps
    elif kind == 4: na, nb, no, fl = B * L * Cout, Cout * Cin * taps, B * L * Cin, 2.0 * B * L * Cin * Cout * taps
    else: na, nb, no, fl = B * L * Cout, B * L * Cin, Cout * Cin * taps, 2.0 * B * L * Cin * Cout * taps
    a = torch.randn(na, device='cuda'); b = torch.randn(nb, device='cuda'); out = torch.zeros(no, device='cuda')
    ms = C.c_float(0)
    N_.check(lib.emb_k_gemm_time(kind, 1, C.c_void_p(a.data_ptr()), C.c_void_p(b.data_ptr()), C.c_void_p(out.data_ptr()),
                                 M, N, K, B, L, Cin, Cout, taps, reps, C.byref(ms), None))
    print(f'{label:28s} kind {kind}  {ms.value * 1e3:9.1f} us  {fl / 1e9:9.2f} GFLOP  {fl / (ms.value * 1e-3) / 1e12:8.1f} TFLOP/s', flush=True)


which = sys.argv[1] if len(sys.argv) > 1 else 'all'
Bn = 8192
if which in ('all', 'conv'):
    for name, (L, Cin, Cout) in {'conv1': (124, 64, 96), 'conv2': (58, 96, 256), 'conv3': (25, 256, 512)}.items():
        for kind, kn in ((3, 'fwd'), (4, 'dgrad'), (5, 'wgrad')):
            t(kind, B=Bn, L=L, Cin=Cin, Cout=Cout, taps=15, label=f'{name} {kn}')
if which in ('all', 'lin'):
    t(0, M=Bn, N=1024, K=4096, label='dock1 fwd')
    t(1, M=Bn, N=4096, K=1024, label='dock1 dgrad')
    t(2, M=1024, N=4096, K=Bn, label='dock1 wgrad')
    t(0, M=Bn, N=256, K=568, label='ffnn0 fwd')
    t(2, M=256, N=568, K=Bn, label='ffnn0 wgrad')
    t(2, M=128, N=256, K=Bn, label='ffnn1 wgrad')
    t(2, M=64, N=128, K=Bn, label='ffnn2 wgrad')
    t(2, M=32, N=64, K=Bn, label='ffnn3 wgrad')
    t(2, M=512, N=1024, K=Bn, label='post0 wgrad')
    t(2, M=256, N=512, K=Bn, label='post1 wgrad')
    t(1, M=Bn, N=64, K=32, label='ffnn3 dgrad')
    t(0, M=Bn, N=512, K=1024, label='post0 fwd')
