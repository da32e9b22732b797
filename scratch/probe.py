import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from embrace_b200 import _native as N
from oracle.embracenet_oracle import bf16_round as q
lib = N.lib()
rs = np.random.RandomState(0)
for mode in (0, 1):
    a = rs.standard_normal((144, 64) if mode == 0 else (128, 128)); b = rs.standard_normal((64, 64) if mode == 0 else (144, 64))
    ta, tb = torch.tensor(a, dtype=torch.float32).cuda(), torch.tensor(b, dtype=torch.float32).cuda()
    for use_bo in (0, 1):
        res = []
        for shift in range(0, 16):
            out = torch.zeros(128, 64, device='cuda')
            N.check(lib.emb_k_umma_shift_probe(mode, shift, use_bo, C.c_void_p(ta.data_ptr()), C.c_void_p(tb.data_ptr()), C.c_void_p(out.data_ptr()), None))
            if mode == 0:
                ref = q(a)[shift:shift + 128] @ q(b).T
            else:
                ref = q(a).T @ q(b)[shift:shift + 128]
            err = np.abs(out.cpu().numpy() - ref).max() / np.abs(ref).max()
            res.append('%d:%s' % (shift, 'ok' if err < 2e-3 else '%.1e' % err))
        print('mode', mode, 'base_offset' if use_bo else 'no base_offset', ' '.join(res))
