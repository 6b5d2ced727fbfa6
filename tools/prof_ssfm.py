"""Short cfg2-shaped propagation for profiling under ncu (never a bench number):
N = 2^20 dual-pol, hz = 0.08 km, `--steps` SSFM steps (default 12), fixed step, no amplifier."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import bench
from opticommpy_b200.channels import manakov_rows_device

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=12)
ap.add_argument("--n", type=int, default=1 << 20)
a = ap.parse_args()
x = bench.synth_waveform(1, a.n)
rows = torch.from_numpy(np.ascontiguousarray(x.T.astype(np.complex64))).cuda()
p = bench.channel_param(1, Ltotal=0.08 * a.steps, Lspan=0.08 * a.steps, amp=None)
for rep in range(2):
    r = rows.clone()
    st = manakov_rows_device(r, p)
torch.cuda.synchronize()
print(st)
