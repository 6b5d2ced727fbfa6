"""Generate tests/golden/ref_cfg3.npz: the UNMODIFIED reference's receiver chain on cfg3 geometry
(BASELINE configs[2]): edc(800 km, 448 taps) -> 2x2 mimoAdaptEqualizer(CMA -> RDE, nTaps = 31, SpS = 2) ->
cpr/bps (B = 64, window 25) on 2^17 symbols x 2 polarisations.  Build container only:

    NUMBA_CACHE_DIR=/tmp/nbcache PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_cfg3.py

The input is rebuilt from its seed by tests/golden/cfg3_signal.py (numpy only), so only outputs are stored:
all hard decisions (uint8), the BPS phases, the final equalizer taps and every 8th output sample.
"""
import os
import sys
from unittest.mock import MagicMock

os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/nbcache")
sys.dont_write_bytecode = True
for _m in ["matplotlib", "matplotlib.pyplot", "matplotlib.mlab", "matplotlib.cm", "matplotlib.colors",
           "matplotlib.animation", "mpl_scatter_density", "simple_pid", "prettytable"]:
    sys.modules[_m] = MagicMock()
sys.path.insert(0, os.environ.get("OPTICOMMPY_REF", "/root/reference"))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

import numpy as np  # noqa: E402

from cfg3_signal import make_signal  # noqa: E402
from optic.comm.modulation import grayMapping  # noqa: E402
from optic.dsp.carrierRecovery import cpr  # noqa: E402
from optic.dsp.equalization import edc, mimoAdaptEqualizer  # noqa: E402
from optic.utils import parameters  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_cfg3.npz")
NSYM = 1 << 17
MU = [1e-3, 2e-4]  # CMA at 5e-3 does not converge on this unit-power 31-tap geometry (checked with the oracle)
c = grayMapping(16, "qam")
c = c / np.sqrt(np.mean(np.abs(c) ** 2))
x, sym = make_signal(NSYM, c, seed=0)


def P(**kw):
    p = parameters()
    for k, v in kw.items():
        setattr(p, k, v)
    return p


y1 = edc(x, P(L=800, D=16, Fc=193.1e12, Fs=64e9, Rs=32e9))
pq = P(nTaps=31, SpS=2, M=16, constType="qam", alg=["cma", "rde"], mu=MU,
       L=[int(0.2 * NSYM), int(0.8 * NSYM)], prgsBar=False, returnResults=True, prec=np.complex64)
y2, H, errSq, Hiter = mimoAdaptEqualizer(y1, pq)
y3, ph = cpr(y2, P(alg="bps", M=16, constType="qam", N=25, B=64, runFOE=False, returnPhases=True))


def dec(z):
    return np.argmin(np.abs(z[..., None] - c), axis=-1).astype(np.uint8)


G = {
    "nsym": np.int64(NSYM), "seed": np.int64(0), "mu": np.array(MU),
    "edc_sub": y1[::8].astype(np.complex64),
    "eq_y_sub": y2[::8], "eq_H": H, "eq_dec": dec(y2), "eq_err_sub": errSq[:, ::8],
    "cpr_ph": ph.astype(np.float64), "cpr_out_sub": y3[::8], "cpr_dec": dec(y3),
}
np.savez_compressed(OUT, **G)
print({k: (np.asarray(v).shape, np.asarray(v).dtype) for k, v in G.items()})
t = y3[NSYM // 2:NSYM - 1000]
d = np.min(np.abs(t[..., None] - c), axis=-1)
print("rms distance of the recovered tail to the nearest constellation point:", np.sqrt(np.mean(d ** 2)),
      "(half the minimum spacing is", 1 / np.sqrt(10), ")")
