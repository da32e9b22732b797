import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tests.test_gpu_tc_gemm import run, q, O
for (B, L, Cin, Cout, k) in [(7, 58, 96, 256, 15), (3, 124, 64, 96, 15), (5, 124, 32, 32, 11)]:
    for rep in range(3):
        rs = np.random.RandomState(B + L + Cin + Cout + k + 1)
        g, W = rs.standard_normal((B, L, Cout)), rs.standard_normal((Cout, Cin, k))
        dx, _, _ = O.conv1d_bwd(np.zeros((B, Cin, L)), q(W), np.transpose(q(g), (0, 2, 1)))
        ref = np.transpose(dx, (0, 2, 1))
        got = run(4, 1, g, W, (B * L, Cin), B=B, L=L, Cin=Cin, Cout=Cout, taps=k).reshape(B, L, Cin)
        err = np.abs(got - ref) / np.abs(ref).max()
        bad = np.argwhere(err > 2e-3)
        print((B, L, Cin, Cout, k), rep, 'max err', err.max(), 'n bad', len(bad), 'nan', np.isnan(got).sum())
        if len(bad):
            print(' bad samples', sorted(set(bad[:, 0])), 'positions', sorted(set(bad[:, 1]))[:20], 'channels', sorted(set(bad[:, 2]))[:20], len(set(bad[:, 2])))
