"""The device-side transmitter once (11-channel DP-16QAM, 2^16 symbols x 16 SpS) for an ncu capture; never a bench number."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from opticommpy_b200.tx import wdm_tx_rows_device


class Bag:
    def __init__(self, **kw):
        self.__dict__.update(kw)


for _ in range(2):
    rows, symb, p = wdm_tx_rows_device(Bag(M=16, Rs=32e9, SpS=16, nBits=4 * (1 << 16), pulseType="rrc", nFilterTaps=1024, pulseRollOff=0.01,
                                           powerPerChannel=-2.0, nChannels=11, wdmGridSpacing=37.5e9, nPolModes=2, seed=123, prgsBar=False))
print(rows.shape, float(abs(rows).mean()))
