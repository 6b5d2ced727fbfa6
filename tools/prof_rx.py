"""cfg3 receiver chain once through rxChain (2^21 symbols x 2 pol by default) for an ncu launch list; never a bench number."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden"))
import numpy as np

from cfg3_signal import make_signal
from opticommpy_b200.modulation import grayMapping
from opticommpy_b200.rxchain import rxChain


class Bag:
    def __init__(self, **kw):
        self.__dict__.update(kw)


nsym = 1 << (int(sys.argv[1]) if len(sys.argv) > 1 else 21)
c = grayMapping(16, "qam").astype(np.complex128)
c /= np.sqrt(np.mean(np.abs(c) ** 2))
x, _ = make_signal(nsym, c, seed=0)
x = x.astype(np.complex64)
t = {}
out = rxChain(x, Bag(L=800, D=16, Fc=193.1e12, Fs=64e9, Rs=32e9),
              Bag(nTaps=31, SpS=2, M=16, constType="qam", alg=["cma", "rde"], mu=[1e-3, 2e-4], L=[int(0.2 * nsym), int(0.8 * nsym)], prgsBar=False),
              Bag(alg="bps", M=16, constType="qam", N=25, B=64, runFOE=False), timing=t)
print(t)
