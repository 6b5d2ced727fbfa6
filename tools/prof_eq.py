"""Small equalizer run for ncu (2x2, 31 taps, CMA->RDE, 2^15 symbols)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tools.bench_rxdsp import make_signal, Bag
from opticommpy_b200.equalization import mimoAdaptEqualizer
n = 1 << 15
x, _ = make_signal(n)
p = Bag(nTaps=31, SpS=2, M=16, constType="qam", alg=["cma", "rde"], mu=[5e-3, 2e-4], L=[n // 5, n - n // 5], prgsBar=False)
for _ in range(2):
    y = mimoAdaptEqualizer(x, p)
print(y.shape)
