import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.argv = [sys.argv[0], 'none']
from scratch.gemm_bench import t
for name, (L, Cin, Cout) in {'conv1': (124, 64, 96), 'conv2': (58, 96, 256), 'conv3': (25, 256, 512)}.items():
    t(5, B=8192, L=L, Cin=Cin, Cout=Cout, taps=15, label=f'{name} wgrad')
