"""ORACLE (test infrastructure, NOT product code) — CPU restatement of the reference's receiver
DSP chain: ``edc`` (numpy), ``mimoAdaptEqualizer`` (stage driver in Python around the C loop in
``rxdsp_oracle.c``), ``bps`` (C) and the ``cpr`` wrapper (numpy).  float64/complex128 arithmetic.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference``
legs may import this module.  Pinned against the reference through ``tests/golden/``.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np
import scipy.constants as const
from numpy.fft import fft, fftfreq, fftshift, ifft

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle_rxdsp.so")
_lib = None

ALG_IDS = {"cma": 0, "rde": 1, "nlms": 2, "dd-lms": 3, "da-rde": 4, "static": 5}


def build(force=False):
    if force or not os.path.exists(_LIB_PATH):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.oracle_core_adapt_eq.restype = C.c_int
        _lib.oracle_core_adapt_eq.argtypes = [C.c_void_p] * 7 + [C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int,
                                                                 C.c_double, C.c_void_p, C.c_int, C.c_int]
        _lib.oracle_bps.restype = None
        _lib.oracle_bps.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                    C.c_void_p, C.c_void_p]
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


# ---- EDC ------------------------------------------------------------------------------------------
def edc_taps(L, D, Fc, Fs, Rs=32e9, NfilterCoeffs=None):
    """Number of taps and time-domain taps of the CD-compensation FIR
    (optic/dsp/equalization.py:90-105; optic/dsp/core.py:1016)."""
    c_kms = const.c / 1e3
    lam = c_kms / Fc
    beta2 = -(D * lam**2) / (2 * np.pi * c_kms)
    if NfilterCoeffs is None:
        NfilterCoeffs = int(2 * np.ceil(6.67 * np.abs(beta2) * L * Rs**2 * (Fs / Rs)))
    w = 2 * np.pi * Fs * fftfreq(NfilterCoeffs)
    Hf = np.exp(-1j * (beta2 / 2) * (w**2) * L)
    return fftshift(ifft(Hf))


def edc(sigIn, L, D, Fc, Fs, Rs=32e9, NfilterCoeffs=None, Nfft=None):
    """Overlap-save CD compensation restated from optic/dsp/core.py:1002-1046: block hop
    ``d = NFFT-K+1``, front pad ``K-1``, keep ``y_blk[K-1:]``, return ``y[D:D+len]`` with
    ``D=(K-1)//2``; complex128 arithmetic; result cast to the input dtype (equalization.py:110)."""
    sig = np.asarray(sigIn)
    one_d = sig.ndim == 1
    x2 = sig.reshape(sig.shape[0], -1)
    h = edc_taps(L, D, Fc, Fs, Rs, NfilterCoeffs)
    K = len(h)
    if Nfft is None:
        Nfft = 2 ** int(np.ceil(np.log2(K)))
    d = Nfft - K + 1
    Dly = (K - 1) // 2
    Hf = fft(np.pad(h, (0, Nfft - K)))
    out = np.zeros(x2.shape, dtype=np.complex128)
    for m in range(x2.shape[1]):
        n = x2.shape[0]
        nblk = int(np.ceil((n + K - 1) / d))
        xp = np.pad(x2[:, m].astype(np.complex128), (K - 1, nblk * d + (K - 1) - n + Dly))
        y = np.zeros(nblk * d, dtype=np.complex128)
        for b in range(nblk):
            y[b * d:(b + 1) * d] = ifft(fft(xp[b * d:b * d + Nfft]) * Hf)[K - 1:]
        out[:, m] = y[Dly:Dly + n]
    out = out.astype(sig.dtype) if np.iscomplexobj(sig) else out.real.astype(sig.dtype)
    return out.reshape(-1) if one_d else out


def fir_direct(x, h):
    """conv(x, h)[D : D+len(x)] by direct summation — the definition overlap-save must equal."""
    K = len(h)
    return np.convolve(x, h)[(K - 1) // 2:(K - 1) // 2 + len(x)]


# ---- adaptive equalizer ------------------------------------------------------------------------------
def normalized_constellation(c, shapingFactor=0.0):
    c = np.asarray(c).astype(np.complex128)
    px = np.exp(-shapingFactor * np.abs(c) ** 2)
    px = px / np.sum(px)
    return c / np.sqrt(np.sum(np.abs(c) ** 2 * px))


def core_adapt_eq(x, ref, H, Hwl, L, SpS, alg, mu, constSymb, runWL=False, storeCoeff=False):
    """One training stage (optic/dsp/equalization.py:354-516) in C, complex128."""
    x = np.ascontiguousarray(x, dtype=np.complex128)
    nModes = x.shape[1]
    nTaps = H.shape[1]
    ref_c = np.ascontiguousarray(ref, dtype=np.complex128) if ref is not None else None
    H = np.ascontiguousarray(H, dtype=np.complex128).copy()
    Hwl = np.ascontiguousarray(Hwl, dtype=np.complex128).copy()
    y = np.zeros((L, nModes), dtype=np.complex128)
    err = np.zeros((nModes, L), dtype=np.float64)
    Hiter = np.zeros((nModes**2, nTaps, L), dtype=np.complex128) if storeCoeff else None
    c = np.ascontiguousarray(constSymb, dtype=np.complex128)
    if (L - 1) * SpS + nTaps > x.shape[0]:
        raise IndexError("window runs past the end of the input")
    rc = lib().oracle_core_adapt_eq(_p(x), _p(ref_c), _p(H), _p(Hwl), _p(y), _p(err), _p(Hiter), L, nModes, nTaps,
                                    SpS, ALG_IDS[alg], float(mu), _p(c), len(c), int(runWL))
    if rc:
        raise ValueError("Equalization algorithm not specified (or incorrectly specified).")
    return y, H, Hwl, err, (Hiter if storeCoeff else H[..., None].copy())


def rls_stage(x, ref, H, L, SpS, lam, constSymb, dd=False, storeCoeff=False):
    """One 'rls' (dd=False) or 'dd-rls' (dd=True) stage, float64 restatement of coreAdaptEq + rlsUp / ddrlsUp
    (optic/dsp/equalization.py:461-473, 576-644, 712-785).  Sd starts from the identity per input mode
    (:447-451; the reference leaves it undefined for 'dd-rls' — the identity is used for both here)."""
    nT = H.shape[1]
    nM = x.shape[1]
    H = H.astype(np.complex128).copy()
    Sd = [np.eye(nT, dtype=np.complex128) for _ in range(nM)]
    y = np.zeros((L, nM), dtype=np.complex128)
    err2 = np.zeros((nM, L))
    Hiter = np.zeros((nM * nM, nT, L), dtype=np.complex128) if storeCoeff else None
    for ind in range(L):
        win = x[ind * SpS: ind * SpS + nT, :]
        out = np.array([sum(H[m + N * nM] @ win[:, N] for N in range(nM)) for m in range(nM)])  # :464-468
        y[ind] = out
        if dd:
            target = np.array([constSymb[np.argmin(np.abs(out[m] - constSymb))] for m in range(nM)])  # :751-753
        else:
            target = ref[ind]
        err = target - out  # :614 / :754
        for N in range(nM):
            u = np.conj(win[:, N])                       # :627
            A = Sd[N] @ u                                # :632
            B = np.conj(u) @ Sd[N]                       # :633
            C = np.conj(u) @ A                           # :634
            Sd[N] = (Sd[N] - np.outer(A, B) / (lam + C)) / lam   # :635-637
            Yv = Sd[N] @ u                               # :639
            for m in range(nM):
                H[m + N * nM] += err[m] * Yv             # :641
        err2[:, ind] = np.abs(err) ** 2
        if storeCoeff:
            Hiter[:, :, ind] = H
    return y, H, err2, Hiter


def mimo_adapt_equalizer(sigIn, symbRef, constSymb, nTaps=15, SpS=2, alg=("nlms",), mu=(1e-3,), L=None,
                         numIter=1, runWL=False, storeCoeff=False, H=None, shapingFactor=0.0, lambdaRLS=0.99):
    """Stage driver restated from optic/dsp/equalization.py:205-319: orientation, casts
    (float32 step sizes), zero padding of floor(nTaps/2) rows, centre-spike taps, stages with the
    taps carried over and stage 0 repeated ``numIter`` times.  ``constSymb`` is the raw
    grayMapping constellation (complex64); it is power-normalised here (:234-241)."""
    sig = np.asarray(sigIn)
    if sig.ndim == 1:
        sig = sig.reshape(-1, 1)
    if sig.shape[1] > sig.shape[0]:
        sig = sig.T
    ref = sig.copy() if symbRef is None or not len(symbRef) else np.asarray(symbRef)
    if ref.ndim == 1:
        ref = ref.reshape(-1, 1)
    if ref.shape[1] > ref.shape[0]:
        ref = ref.T
    nModes = sig.shape[1]
    sig = sig.astype(np.complex64).astype(np.complex128)  # the reference rounds inputs to prec (:222-223)
    ref = ref.astype(np.complex64).astype(np.complex128)
    mu = np.atleast_1d(np.array(mu).astype(np.float32)).astype(np.float64)
    Lpad = nTaps // 2
    pad = np.zeros((Lpad, nModes), dtype=np.complex128)
    x = np.concatenate((pad, sig, pad))
    c = np.asarray(constSymb).astype(np.complex64)
    px = np.exp(-shapingFactor * np.abs(c) ** 2)
    px = px / np.sum(px)
    c = c / np.sqrt(np.sum(np.abs(c) ** 2 * px))  # complex64 arithmetic, like the reference
    c = c.astype(np.complex128)
    total = int(np.fix((len(x) - nTaps) / SpS + 1))
    if not L:
        L = [total]
    if H is None:
        H = np.zeros((nModes**2, nTaps), dtype=np.complex128)
        for i in range(nModes):
            H[i + i * nModes, nTaps // 2] = 1.0
    Hwl = np.zeros((nModes**2, nTaps), dtype=np.complex128)
    y = np.zeros((total, nModes), dtype=np.complex128)
    err = np.zeros((nModes, total), dtype=np.float64)
    Hiter = None
    n0 = 0
    for stage, name in enumerate(alg):
        n1 = n0 + int(L[stage])
        for _ in range(numIter if stage == 0 else 1):
            if name in ("rls", "dd-rls"):
                ys, H, es, Hiter = rls_stage(x[n0 * SpS:(n1 + 2 * Lpad) * SpS], ref[n0:n1], H, n1 - n0, SpS,
                                             float(np.float32(lambdaRLS)), c, name == "dd-rls", storeCoeff)
                y[n0:n1] = ys
                err[:, n0:n1] = es
                continue
            ys, H, Hwl, es, Hiter = core_adapt_eq(x[n0 * SpS:(n1 + 2 * Lpad) * SpS], ref[n0:n1], H, Hwl, n1 - n0, SpS,
                                                   name, mu[stage], c, runWL, storeCoeff)
            y[n0:n1] = ys
            err[:, n0:n1] = es
        n0 = n1
    return y, H, Hwl, err, Hiter


# ---- carrier recovery -------------------------------------------------------------------------------
def bps(sigIn, N, constSymb, B):
    """Blind phase search (optic/dsp/carrierRecovery.py:172-223) in C; returns (phases, indices)."""
    x = np.ascontiguousarray(np.asarray(sigIn), dtype=np.complex128)
    if x.ndim == 1:
        x = x.reshape(-1, 1)
    c = np.ascontiguousarray(constSymb, dtype=np.complex128)
    idx = np.zeros(x.shape, dtype=np.int32)
    ph = np.zeros(x.shape, dtype=np.float64)
    lib().oracle_bps(_p(x), x.shape[0], x.shape[1], _p(c), len(c), int(B), int(N), _p(idx), _p(ph))
    return ph, idx


def pnorm(x):
    return x / np.sqrt(np.mean(x * np.conj(x)).real)


def fourth_power_foe(sig, Fs, M=4):
    """optic/dsp/carrierRecovery.py:333-371."""
    n = sig.shape[0]
    f = fftshift(Fs * fftfreq(n))
    t = np.arange(n) / Fs
    out = sig.copy()
    fo = np.zeros(sig.shape[1])
    for m in range(sig.shape[1]):
        fo[m] = f[np.argmax(10 * np.log10(np.abs(fftshift(fft(sig[:, m] ** M)))))] / M
        out[:, m] = sig[:, m] * np.exp(-1j * 2 * np.pi * fo[m] * t)
    return out, fo


def cpr_bps(sigIn, constSymb, N=35, B=64, Ts=1 / 32e9, runFOE=True, foeM=4, shapingFactor=0.0):
    """``cpr(..., alg='bps')`` (optic/dsp/carrierRecovery.py:110-169): optional FOE + pnorm,
    bps with N//2, unwrap(4φ)/4, pnorm(x e^{jφ}).  ``constSymb``: raw grayMapping constellation."""
    sig = np.asarray(sigIn)
    one_d = sig.ndim == 1
    if one_d:
        sig = sig.reshape(-1, 1)
    c = np.asarray(constSymb).astype(np.complex64)
    px = np.exp(-shapingFactor * np.abs(c) ** 2)
    px = px / np.sum(px)
    c = c / np.sqrt(np.sum(np.abs(c) ** 2 * px))
    if runFOE:
        sig, _ = fourth_power_foe(sig, 1 / Ts, foeM)
        sig = pnorm(sig)
    ph, _ = bps(sig, N // 2, c, B)
    ph = np.unwrap(4 * ph, axis=0) / 4
    out = pnorm(sig * np.exp(1j * ph))
    if one_d:
        return out.flatten(), ph.flatten()
    return out, ph
