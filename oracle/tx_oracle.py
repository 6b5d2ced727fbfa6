"""CPU oracle of the WDM transmitter (TEST INFRASTRUCTURE — never imported by the product path).

A numpy restatement, in float64 / complex128, of what the reference computes in
``optic.models.tx.simpleWDMTx`` (optic/models/tx.py:42-228) and its helpers.  Pinned to the unmodified reference by
``tests/golden/ref_tx.npz`` (made by ``tests/golden/make_golden_tx.py``, which imports ``/root/reference``).

    symbol_source     optic/comm/sources.py:112-211   (legacy numpy generator: seed, then choice with a pmf)
    pulse_shape       optic/dsp/core.py:128-173 (rrc), 176-208 (rc), 211-269 (pulseShape)
    phase_noise       optic/dsp/core.py:791-826
    iqm               optic/models/devices.py:147-216; calcMZM / calcPM optic/dsp/core.py:1075-1130
    simple_wdm_tx     optic/models/tx.py:101-228
"""
from __future__ import annotations

import numpy as np


def raster_constellation(M, constType):
    """Constellation in the reference's raster (pre-Gray) order: qamConst / pamConst / pskConst (modulation.py:121-197)."""
    if constType == "qam":
        side = int(np.sqrt(M))
        lev = np.arange(-(side - 1), side, 2).astype(float)
        rows = []
        for r in range(side):
            re = lev[::-1] if r % 2 else lev
            rows.append(re + 1j * lev[side - 1 - r])
        return np.concatenate(rows)
    if constType == "pam":
        return np.arange(-(M - 1), M, 2).astype(float)
    if constType == "psk":
        return np.exp(1j * np.arange(M) * 2 * np.pi / M)
    raise ValueError(constType)


def symbol_source(nSymbols, M, constType, seed, shapingFactor=0.0, dist="uniform"):
    """sources.py:167-211: np.random.seed(seed); unit-power constellation; np.random.choice(const, nSymbols, p=px)."""
    rs = np.random.RandomState(seed)
    c = raster_constellation(M, constType).astype(complex)
    if dist == "uniform":
        px = np.ones(M) / M
    else:
        px = np.exp(-shapingFactor * np.abs(c) ** 2)
        px = px / np.sum(px)
    c = c / np.sqrt(np.sum(px * np.abs(c) ** 2))
    # numpy's legacy choice with p: uniform draws searched in the normalised cdf, side='right'
    cdf = np.cumsum(px)
    cdf /= cdf[-1]
    idx = np.searchsorted(cdf, rs.random_sample(nSymbols), side="right")
    return c[idx]


def rrc_taps(t, alpha, Ts=1.0):
    """core.py:128-173."""
    out = np.zeros(len(t))
    for i, ti in enumerate(t):
        if ti == 0:
            out[i] = (1 / Ts) * (1 + alpha * (4 / np.pi - 1))
        elif abs(ti) == Ts / (4 * alpha):
            out[i] = (alpha / (Ts * np.sqrt(2))) * ((1 + 2 / np.pi) * np.sin(np.pi / (4 * alpha)) + (1 - 2 / np.pi) * np.cos(np.pi / (4 * alpha)))
        else:
            t1, t2 = np.pi * ti / Ts, 4 * alpha * ti / Ts
            out[i] = (1 / Ts) * (np.sin(t1 * (1 - alpha)) + 4 * alpha * ti / Ts * np.cos(t1 * (1 + alpha))) / (np.pi * ti * (1 - t2 ** 2))
    return out


def pulse_shape(pulseType, SpS, nFilterTaps, rollOff):
    """core.py:211-269 ('rrc' and 'rect'; the taps are normalised to unit sum)."""
    if pulseType == "rect":
        p = np.concatenate((np.zeros(int(SpS / 2)), np.ones(SpS), np.zeros(int(SpS / 2))))
    elif pulseType == "rrc":
        t = np.linspace(-nFilterTaps // 2, nFilterTaps // 2, nFilterTaps) * (1 / SpS)
        p = rrc_taps(t, rollOff, 1)
    else:
        raise ValueError(pulseType)
    return p / np.sum(p)


def phase_noise(lw, N, Ts, seed):
    """core.py:817-826: random walk with steps N(0, 2 pi lw Ts); the first sample is 0."""
    rs = np.random.RandomState(seed)
    steps = rs.normal(0, np.sqrt(2 * np.pi * lw * Ts), N - 1) if N > 1 else np.zeros(0)
    return np.concatenate(([0.0], np.cumsum(steps)))


def fir_same(h, x):
    """firFilter (core.py:87-125): fftconvolve(x, h, 'same')."""
    full = np.convolve(x, h) if len(x) * len(h) < 1 << 22 else _fftconv(x, h)
    d = (len(h) - 1) // 2
    return full[d:d + len(x)]


def _fftconv(x, h):
    n = len(x) + len(h) - 1
    nfft = 1 << int(np.ceil(np.log2(n)))
    return np.fft.ifft(np.fft.fft(x, nfft) * np.fft.fft(h, nfft))[:n]


def iqm(Ei, u, Vpi=2.0, VbI=-2.0, VbQ=-2.0, Vphi=1.0, ERI=60.0, ERQ=60.0):
    """devices.py:199-216 with calcMZM (core.py:1101-1108) and calcPM (:1130)."""
    def mzm(E, v, Vb, ER):
        er = 10 ** (ER / 10)
        g = 2 * np.sqrt(er) / (er + 1)
        pm = lambda s, w: s * np.exp(1j * (w / Vpi) * np.pi)
        return np.sqrt(1 + g) * pm(E / 2, (v + Vb) / 2) + np.sqrt(1 - g) * pm(E / 2, -(v + Vb) / 2)
    EoI = mzm(Ei / np.sqrt(2), u.real, VbI, ERI)
    EoQ = mzm(Ei / np.sqrt(2), u.imag, VbQ, ERQ)
    return EoI + EoQ * np.exp(1j * (Vphi / Vpi) * np.pi)


def simple_wdm_tx(M=16, constType="qam", Rs=32e9, SpS=16, seed=None, nBits=60000, pulseType="rrc", nFilterTaps=1024,
                  pulseRollOff=0.01, mzmScale=0.5, powerPerChannel=-3, nChannels=5, wdmGridSpacing=50e9, nPolModes=1,
                  laserLinewidth=0.0):
    """tx.py:101-228 for a given seed.  Returns (sigTxWDM (N, nPol), symbTxWDM (nSym, nPol, nCh), freqGrid)."""
    Fs = 1 / ((1 / Rs) / SpS)
    nSym = nBits // int(np.log2(M))
    pulse = pulse_shape(pulseType, SpS, nFilterTaps, pulseRollOff)
    grid = np.arange(-np.floor(nChannels / 2), np.floor(nChannels / 2) + 1, 1) * wdmGridSpacing
    if nChannels % 2 == 0:
        grid = grid + wdmGridSpacing / 2
    Pch = 10 ** (np.asarray(powerPerChannel, dtype=float) / 10) * 1e-3 * np.ones(nChannels)
    N = nSym * SpS
    sig = np.zeros((N, nPolModes), dtype=complex)
    symb = np.zeros((nSym, nPolModes, nChannels), dtype=complex)
    t = np.arange(N) * (1 / Fs)
    s = seed
    for ch in range(nChannels):
        for m in range(nPolModes):
            sy = symbol_source(nSym, M, constType, s)
            s += 1
            symb[:, m, ch] = sy
            up = np.zeros(N, dtype=complex)
            up[::SpS] = sy
            x = fir_same(pulse, up)
            x = x / np.max(np.abs(x))
            if m == 0:
                lo = np.exp(1j * phase_noise(laserLinewidth, N, 1 / Fs, seed))
            e = iqm(lo, mzmScale * x)
            e = np.sqrt(Pch[ch] / nPolModes) * e / np.sqrt(np.mean(np.abs(e) ** 2))
            sig[:, m] += e * np.exp(1j * 2 * np.pi * grid[ch] * t)
    return sig, symb, grid
