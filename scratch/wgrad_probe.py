import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.argv = [sys.argv[0], 'none']
from scratch.gemm_bench import t
for K in (256, 1024, 2048, 8192, 32768):
    t(2, M=128, N=256, K=K, label=f'wgrad 128x256 K={K}')
for K in (1024, 8192):
    t(2, M=1024, N=1024, K=K, label=f'wgrad 1024x1024 K={K}')
    t(1, M=K, N=1024, K=1024, label=f'dgrad M={K} 1024x1024')
