"""Seeded cfg3-shaped receiver input (BASELINE configs[2] geometry): random DP-16QAM symbols at
2 SpS, low-pass pulse, 2x2 polarisation rotation, chromatic dispersion of `L_km`, laser phase noise, AWGN.
numpy only, so that the golden generator (which imports the reference) and the GPU tests (which cannot) build the
identical array from the seed instead of shipping it."""
import numpy as np

C_KMS = 299792.458


def make_signal(nsym, const, seed=0, L_km=800.0, Fs=64e9, D=16.0, Fc=193.1e12, theta=0.5, lw=100e3, sigma=0.07):
    """Returns (x (2 nsym, 2) complex128 unit power, sym (nsym, 2) complex128).  `const` = unit-power constellation."""
    rng = np.random.default_rng(seed)
    sym = const[rng.integers(0, len(const), size=(nsym, 2))]
    up = np.zeros((2 * nsym, 2), dtype=complex)
    up[0::2] = sym
    X = np.fft.fft(up, axis=0)
    f = np.fft.fftfreq(2 * nsym)
    X *= (np.abs(f) < 0.27)[:, None]
    lam = C_KMS / Fc
    beta2 = -(D * lam ** 2) / (2 * np.pi * C_KMS)
    w = 2 * np.pi * Fs * f
    X *= np.exp(1j * (beta2 / 2) * w ** 2 * L_km)[:, None]   # all-pass CD, sign of linearFiberChannel (channels.py:99)
    x = np.fft.ifft(X, axis=0) * 2
    rot = np.array([[np.cos(theta), -np.sin(theta)], [np.sin(theta), np.cos(theta)]])
    x = x @ rot.T
    pn = np.cumsum(rng.normal(scale=np.sqrt(2 * np.pi * lw / Fs), size=2 * nsym))
    x = x * np.exp(1j * pn)[:, None]
    x += sigma * (rng.normal(size=x.shape) + 1j * rng.normal(size=x.shape))
    return x / np.sqrt(np.mean(np.abs(x) ** 2)), sym
