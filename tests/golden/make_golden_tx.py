"""Generate tests/golden/ref_tx.npz (WDM transmitter) by running the UNMODIFIED reference
(/root/reference) with seeds.  Build container only:

    NUMBA_CACHE_DIR=/tmp/nbcache PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_tx.py

Same import recipe as make_golden.py (plotting modules stubbed, numba cache in scratch).  The reference's firFilter for
the transmitter is optic.dsp.core.firFilter here (no CuPy in the container: tx.py:29-37 falls back to it).
"""
import os
import sys
from unittest.mock import MagicMock

os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/nbcache")
sys.dont_write_bytecode = True
for _m in ["matplotlib", "matplotlib.pyplot", "matplotlib.mlab", "matplotlib.cm", "matplotlib.colors",
           "matplotlib.animation", "mpl_scatter_density", "simple_pid", "prettytable", "tqdm", "tqdm.notebook"]:
    sys.modules[_m] = MagicMock()
sys.modules["tqdm.notebook"].tqdm = lambda it, **kw: it
sys.path.insert(0, os.environ.get("OPTICOMMPY_REF", "/root/reference"))

import numpy as np  # noqa: E402

from optic.comm.sources import symbolSource  # noqa: E402
from optic.dsp.core import phaseNoise, pulseShape  # noqa: E402
from optic.models.devices import basicLaserModel, iqm  # noqa: E402
from optic.models.tx import simpleWDMTx  # noqa: E402
from optic.utils import parameters  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_tx.npz")
G = {}

# symbolSource: constellation families, uniform and shaped pmf
for name, M, ct, dist, sf in (("qam16", 16, "qam", "uniform", 0.0), ("qam64mb", 64, "qam", "maxwell-boltzmann", 0.05),
                              ("psk8", 8, "psk", "uniform", 0.0), ("pam4", 4, "pam", "uniform", 0.0)):
    p = parameters()
    p.nSymbols, p.M, p.constType, p.dist, p.shapingFactor, p.seed = 3000, M, ct, dist, sf, 99
    G[f"src_{name}"] = np.asarray(symbolSource(p))

# pulseShape
for name, pt, sps, nt, ro in (("rrc16", "rrc", 16, 1024, 0.01), ("rrc8", "rrc", 8, 257, 0.1), ("rect", "rect", 8, 0, 0.0)):
    q = parameters()
    q.pulseType, q.SpS, q.nFilterTaps, q.rollOff = pt, sps, nt, ro
    G[f"pulse_{name}"] = np.asarray(pulseShape(q))

# phaseNoise and iqm
G["pn"] = phaseNoise(100e3, 4096, 1 / 512e9, seed=5)
rng = np.random.default_rng(3)
u = 0.5 * (rng.uniform(-1, 1, 2000) + 1j * rng.uniform(-1, 1, 2000))
G["iqm_u"] = u
G["iqm_out"] = iqm(np.exp(1j * G["pn"][:2000]), u)

# basicLaserModel: LO with linewidth, RIN and a frequency shift; and the ideal CW case
pl = parameters()
pl.P, pl.lw, pl.RIN_var, pl.Fs, pl.Ns, pl.seed, pl.freqShift = 10, 100e3, 1e-20, 512e9, 5000, 789, 37.5e9 - 128e6
G["laser_pn"] = basicLaserModel(pl)
pl2 = parameters()
pl2.P, pl2.lw, pl2.RIN_var, pl2.Fs, pl2.Ns, pl2.seed = 7, 0.0, 0, 64e9, 1000, 1
G["laser_cw"] = basicLaserModel(pl2)

# simpleWDMTx: dual-pol 5-channel DP-16QAM (cfg5 shape, shorter), single-pol 3-channel QPSK with a laser linewidth and
# per-channel powers, even channel count
def run(tag, **kw):
    p = parameters()
    p.prgsBar = False
    for k, v in kw.items():
        setattr(p, k, v)
    sig, symb, p = simpleWDMTx(p)
    G[f"tx_{tag}_sig"] = sig
    G[f"tx_{tag}_symb"] = symb
    G[f"tx_{tag}_grid"] = np.asarray(p.wdmFreqGrid)


run("dp5", M=16, Rs=32e9, SpS=8, nBits=4 * 2048, pulseType="rrc", nFilterTaps=1024, pulseRollOff=0.01, powerPerChannel=-2.0,
    nChannels=5, wdmGridSpacing=37.5e9, nPolModes=2, seed=321)
run("sp3", M=4, Rs=10e9, SpS=16, nBits=2 * 1024, pulseType="rrc", nFilterTaps=256, pulseRollOff=0.1,
    powerPerChannel=[-1.0, 0.0, 1.5], nChannels=3, wdmGridSpacing=25e9, nPolModes=1, seed=17, laserLinewidth=100e3)
run("dp4", M=16, Rs=32e9, SpS=4, nBits=4 * 1024, pulseType="rrc", nFilterTaps=128, pulseRollOff=0.2, powerPerChannel=0.0,
    nChannels=4, wdmGridSpacing=40e9, nPolModes=2, seed=5, mzmScale=0.25)
np.savez_compressed(OUT, **G)
print("wrote", OUT, {k: v.shape for k, v in G.items()})
