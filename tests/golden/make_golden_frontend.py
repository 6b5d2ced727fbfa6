"""Generate tests/golden/ref_frontend.npz (Rx front-end glue) by running the UNMODIFIED
reference (/root/reference) on seeded inputs.  Build container only:

    NUMBA_CACHE_DIR=/tmp/nbcache PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_frontend.py

Same import recipe as make_golden.py (plotting modules stubbed, numba cache in scratch).
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

import numpy as np  # noqa: E402

from optic.dsp.core import decimate, firFilter, pulseShape, symbolSync  # noqa: E402
from optic.utils import parameters  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_frontend.npz")
G = {}
rng = np.random.default_rng(77)


def rrc(sps, ntaps, rolloff):
    q = parameters()
    q.pulseType, q.SpS, q.nFilterTaps, q.rollOff = "rrc", sps, ntaps, rolloff
    return pulseShape(q)


# firFilter: root-raised-cosine matched filter (even and odd tap counts), complex and real taps, 1-D and 2-D inputs
x = (rng.normal(size=(6000, 2)) + 1j * rng.normal(size=(6000, 2))).astype(np.complex128)
G["fir_in"] = x
h_rrc = rrc(8, 257, 0.1)
h_rrc = h_rrc / np.max(np.abs(h_rrc))
G["fir_h_rrc"] = h_rrc
G["fir_rrc"] = firFilter(h_rrc, x)
h_even = rng.normal(size=64) + 1j * rng.normal(size=64)
G["fir_h_even"] = h_even
G["fir_even_1d"] = firFilter(h_even, x[:, 0])
G["fir_real_in"] = firFilter(h_rrc, x.real.copy())
G["fir_c64"] = firFilter(h_rrc.astype(np.float32), x.astype(np.complex64))

# decimate: 16 -> 2 samples per symbol on a pulse-shaped 2-mode signal with different timing offsets per mode
sps = 16
nsym = 2048
sym = (rng.integers(0, 2, size=(nsym, 2)) * 2 - 1) + 1j * (rng.integers(0, 2, size=(nsym, 2)) * 2 - 1)
up = np.zeros((nsym * sps, 2), dtype=complex)
up[::sps] = sym
pulse = rrc(sps, 1025, 0.2)
sig = firFilter(pulse / np.max(np.abs(pulse)), up)
sig = firFilter(pulse / np.sum(pulse), sig)
sig[:, 1] = np.roll(sig[:, 1], 5)
sig += 0.02 * (rng.normal(size=sig.shape) + 1j * rng.normal(size=sig.shape))
G["dec_in"] = sig
p = parameters()
p.SpSin, p.SpSout = sps, 2
G["dec_16_2"] = decimate(sig, p)
p.SpSout = 1
G["dec_16_1_1d"] = decimate(sig[:, 1], p)
p.SpSin, p.SpSout = 4, 2
G["dec_4_2"] = decimate(sig[: 4 * 1000], p)

# symbolSync: swapped, delayed (and, for 'real' mode, rotated / conjugated) transmit sequences against a noisy
# 2-SpS received signal
nss = 1200
c16 = np.array([a + 1j * b for a in (-3, -1, 1, 3) for b in (-3, -1, 1, 3)]) / np.sqrt(10)
txs = c16[rng.integers(0, 16, size=(nss, 2))]
rxs = np.repeat(txs, 2, axis=0)
rxs += 0.05 * (rng.normal(size=rxs.shape) + 1j * rng.normal(size=rxs.shape))
tx_amp = np.stack([np.roll(txs[:, 1], 7), np.roll(txs[:, 0], -11)], axis=1)           # swapped + delayed
G["sync_rx"], G["sync_tx_amp"] = rxs, tx_amp
G["sync_amp"] = symbolSync(rxs.copy(), tx_amp.copy(), 2, "amp")
tx_real = np.stack([1j * np.roll(txs[:, 1], 5), np.conj(np.roll(txs[:, 0], -3))], axis=1)  # + rotation / conjugation
G["sync_tx_real"] = tx_real
G["sync_real"] = symbolSync(rxs.copy(), tx_real.copy(), 2, "real")
G["sync_amp_1d"] = symbolSync(rxs[:, 0].copy(), np.roll(txs[:, 0], 9), 2, "amp")

# pdmCoherentReceiver (ideal photodiodes) and delaySignal: a 2-pol field with a frequency-shifted, phase-noisy LO
from optic.dsp.core import delaySignal  # noqa: E402
from optic.models.devices import basicLaserModel, pdmCoherentReceiver  # noqa: E402

nfe = 1 << 13
Fs_fe = 64e9
Es = (rng.normal(size=(nfe, 2)) + 1j * rng.normal(size=(nfe, 2))) * np.sqrt(1e-3 / 4)
Es = np.fft.ifft(np.fft.fft(Es, axis=0) * (np.abs(np.fft.fftfreq(nfe)) < 0.3)[:, None], axis=0)
pl = parameters()
pl.P, pl.lw, pl.RIN_var, pl.Ns, pl.Fs, pl.seed, pl.freqShift = 10, 100e3, 0, nfe, Fs_fe, 789, 1.5e9
Elo = basicLaserModel(pl)
G["fe_Es"], G["fe_Elo"] = Es, Elo
ppd = parameters()
ppd.B, ppd.Fs, ppd.ideal, ppd.seed = 32e9, Fs_fe, True, 1011
pfe = parameters()
pfe.Fs, pfe.polRotation, pfe.pdl, pfe.polDelay = Fs_fe, np.pi / 3, 0, 0
G["fe_rot"] = pdmCoherentReceiver(Es, Elo, pfe, ppd)
pfe = parameters()
pfe.Fs, pfe.polRotation, pfe.pdl, pfe.polDelay = Fs_fe, 0.4, 1.5, 0
pfe.phaseImbX, pfe.phaseImbY, pfe.ampImbX, pfe.ampImbY = 3 * np.pi / 180, -2 * np.pi / 180, 0.5, -0.3
ppd.R = 0.8
G["fe_imb"] = pdmCoherentReceiver(Es, Elo, pfe, ppd)
ppd.R = 1
pfe = parameters()
pfe.Fs, pfe.polRotation, pfe.pdl, pfe.polDelay = Fs_fe, np.pi / 3, 0, 3 / 32e9   # the notebook's front end
G["fe_delay"] = pdmCoherentReceiver(Es, Elo, pfe, ppd)
pfe = parameters()
pfe.Fs, pfe.polRotation, pfe.timeSkewX, pfe.timeSkewY = Fs_fe, 0.2, 4e-12, -6e-12
G["fe_skew"] = pdmCoherentReceiver(Es, Elo, pfe, ppd)
G["fe_1pol"] = pdmCoherentReceiver(Es[:, 0].copy(), Elo, pfe, ppd)
G["delay_c"] = delaySignal(Es[:, 0].copy(), 7.3e-12, Fs_fe)
G["delay_r"] = delaySignal(Es[:, 1].real.copy(), -2.6e-11, Fs_fe)

np.savez_compressed(OUT, **G)
print({k: (v.shape, v.dtype) for k, v in G.items()})
