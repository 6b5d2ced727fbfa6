"""Generate tests/golden/ref_metrics.npz (decisions / error counting) by running
the UNMODIFIED reference (/root/reference) on seeded inputs.  Build container only:

    NUMBA_CACHE_DIR=/tmp/nbcache PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_metrics.py

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

from optic.comm.metrics import fastBERcalc  # noqa: E402
from optic.comm.modulation import demodulateGray, grayMapping, minEuclid  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_metrics.npz")
G = {}
rng = np.random.default_rng(2024)


def noisy(M, constType, n, modes, sigma, rot=0.0, gain=1.0):
    c = grayMapping(M, constType)
    c = c / np.sqrt(np.mean(np.abs(c) ** 2))
    tx = c[rng.integers(0, M, size=(n, modes))].astype(np.complex128)
    w = sigma * (rng.normal(size=tx.shape) + 1j * rng.normal(size=tx.shape))
    return tx, gain * np.exp(1j * rot) * (tx + w)


# 1. minEuclid / demodulateGray on a noisy 16-QAM and an 8-PSK column, plus exact ties (origin, axes)
tx, rx = noisy(16, "qam", 3000, 1, 0.18)
c16 = grayMapping(16, "qam")
G["me_qam16_in"] = rx[:, 0] * np.sqrt(10)
G["me_qam16_idx"] = minEuclid(G["me_qam16_in"], c16)
G["dg_qam16_bits"] = demodulateGray(G["me_qam16_in"], 16, "qam")
ties = np.array([0, 2, 2j, -2, -2j, 1 + 1j, 2 + 2j, 0.5 + 2j, -2 - 1j, 4 + 4j, 1e-30], dtype=np.complex128)
G["me_ties_in"] = ties
G["me_ties_idx"] = minEuclid(ties, c16)
tx8, rx8 = noisy(8, "psk", 1000, 1, 0.12)
G["me_psk8_in"] = rx8[:, 0]
G["me_psk8_idx"] = minEuclid(rx8[:, 0], grayMapping(8, "psk"))
G["dg_psk8_bits"] = demodulateGray(rx8[:, 0], 8, "psk")
pam = (grayMapping(4, "pam")[rng.integers(0, 4, 500)] + 0.4 * rng.normal(size=500)).astype(np.complex128)
G["me_pam4_in"] = pam
G["dg_pam4_bits"] = demodulateGray(pam, 4, "pam")

# 2. fastBERcalc: (a) 16-QAM, 2 modes, common rotation + gain; (b) 64-QAM 1-D; (c) 8-PSK; (d) complex64 input;
#    (e) shaped pmf; (f) wide (modes, symbols) orientation; (g) error-free
tx, rx = noisy(16, "qam", 6000, 2, 0.16, rot=0.21, gain=0.63)
G["ber_a_tx"], G["ber_a_rx"] = tx, rx
G["ber_a"] = np.array(fastBERcalc(rx, tx, 16, "qam"))
tx, rx = noisy(64, "qam", 5000, 1, 0.07, rot=-0.1, gain=1.7)
G["ber_b_tx"], G["ber_b_rx"] = tx[:, 0], rx[:, 0]
G["ber_b"] = np.array(fastBERcalc(rx[:, 0], tx[:, 0], 64, "qam"))
tx, rx = noisy(8, "psk", 4000, 2, 0.2, rot=0.05)
G["ber_c_tx"], G["ber_c_rx"] = tx, rx
G["ber_c"] = np.array(fastBERcalc(rx, tx, 8, "psk"))
tx, rx = noisy(16, "qam", 4000, 2, 0.2, rot=0.3)
G["ber_d_tx"], G["ber_d_rx"] = tx.astype(np.complex64), rx.astype(np.complex64)
G["ber_d"] = np.array(fastBERcalc(G["ber_d_rx"], G["ber_d_tx"], 16, "qam"))
px = np.exp(-0.05 * np.abs(grayMapping(16, "qam")) ** 2)
px = px / px.sum()
G["ber_e_px"] = px
G["ber_e"] = np.array(fastBERcalc(G["ber_a_rx"], G["ber_a_tx"], 16, "qam", px))
G["ber_f"] = np.array(fastBERcalc(G["ber_a_rx"].T.copy(), G["ber_a_tx"].T.copy(), 16, "qam"))
with np.errstate(divide="ignore"):
    G["ber_g"] = np.array(fastBERcalc(G["ber_a_tx"], G["ber_a_tx"], 16, "qam"))
tx, rx = noisy(16, "apsk", 3000, 1, 0.15)  # no rotation correction for apsk (metrics.py:176)
G["ber_h_tx"], G["ber_h_rx"] = tx, rx
G["ber_h"] = np.array(fastBERcalc(rx, tx, 16, "apsk"))

np.savez_compressed(OUT, **G)
print(f"wrote {OUT}: {len(G)} arrays, {os.path.getsize(OUT) / 1e6:.2f} MB")
for k in sorted(G):
    if k.startswith("ber_") and G[k].shape == (3,) + G[k].shape[1:] and G[k].ndim == 2 and G[k].shape[0] == 3:
        print(k, G[k].tolist())
