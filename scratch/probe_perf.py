import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from embrace_b200 import _native as N
lib = N.lib()
rs = np.random.RandomState(0)
for mode in (0, 1):
    a = rs.standard_normal((144, 64) if mode == 0 else (128, 128)); b = rs.standard_normal((64, 64) if mode == 0 else (144, 64))
    ta, tb = torch.tensor(a, dtype=torch.float32).cuda(), torch.tensor(b, dtype=torch.float32).cuda()
    for ncols in (32, 64, 96, 128):
        res = []
        for shift in (0, 1, 3, 4, 7, 8):
            out = torch.zeros(128 * 64 + 8, device='cuda')
            N.check(lib.emb_k_umma_shift_probe(mode, shift, (ncols // 32) * 2, C.c_void_p(ta.data_ptr()), C.c_void_p(tb.data_ptr()), C.c_void_p(out.data_ptr()), None))
            res.append('shift%d: %.0f' % (shift, float(out[128 * 64])))
        print('mode', mode, 'N', ncols, 'cycles/MMA(128xNx16):', '  '.join(res))
