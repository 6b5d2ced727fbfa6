"""ORACLE (test infrastructure, NOT product code) — numpy restatement of the reference's fiber
propagators, float64/complex128 throughout, single-threaded like the reference.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference``
legs may import this module; the product path (``opticommpy_b200``) never does.

Pinned against the reference itself: ``tests/golden/make_golden.py`` imports OptiCommPy v0.11.0
from ``/root/reference`` in the build container, runs ``ssfm`` / ``manakovSSF`` / ``manakovDBP`` /
``edfa`` on seeded inputs and stores the outputs under ``tests/golden/``;
``tests/test_oracle_golden.py`` checks this file against those vectors (rel. error <= 1e-12).

Each function cites the reference lines it restates (paths relative to the OptiCommPy tree).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np
import scipy.constants as const
from numpy.fft import fft, fftfreq, ifft

H_PLANCK = const.h
C_KMS = const.c / 1e3


@dataclass
class FiberConfig:
    """Physical/simulation parameters with the reference defaults
    (optic/models/channels.py:158-170, 305-322; optic/dsp/equalization.py:1026-1041)."""
    Fs: float
    Ltotal: float = 400
    Lspan: float = 80
    hz: float = 0.5
    alpha: float = 0.2
    D: float = 16
    gamma: float = 1.3
    Fc: float = 193.1e12
    amp: object = "edfa"
    NF: float = 4.5
    maxIter: int = 10
    tol: float = 1e-5
    nlprMethod: bool = True
    maxNlinPhaseRot: float = 2e-2
    seed: object = None
    saveSpanN: list = field(default_factory=list)

    @classmethod
    def from_param(cls, param, manakov=True):
        kw = {"Fs": param.Fs}
        for name in ("Ltotal", "Lspan", "hz", "alpha", "D", "gamma", "Fc", "amp", "NF", "maxIter", "tol",
                     "nlprMethod", "maxNlinPhaseRot", "seed"):
            if hasattr(param, name):
                kw[name] = getattr(param, name)
        cfg = cls(**kw)
        if manakov:
            cfg.saveSpanN = getattr(param, "saveSpanN", [cfg.Ltotal // cfg.Lspan])
        return cfg

    # channels.py:187-190 / 344-347
    @property
    def alpha_lin(self):
        return self.alpha / (10 * np.log10(np.exp(1)))

    @property
    def beta2(self):
        lam = C_KMS / self.Fc
        return -(self.D * lam**2) / (2 * np.pi * C_KMS)


def legacy_noise(shape, var, seed):
    """optic/dsp/core.py:758-763: seed the legacy global MT19937 stream, draw all real parts,
    then all imaginary parts, each N(0, var/2)."""
    rs = np.random.RandomState(seed) if seed is not None else np.random.RandomState()
    s = np.sqrt(var / 2)
    re = rs.normal(0, s, shape)
    im = rs.normal(0, s, shape)
    return re + 1j * im


def edfa_numbers(G_dB, NF_dB, Fc, Fs):
    """Linear gain and ASE noise power (optic/models/devices.py:712-722)."""
    NF_lin = 10 ** (NF_dB / 10)
    G_lin = 10 ** (G_dB / 10)
    nsp = (G_lin * NF_lin - 1) / (2 * (G_lin - 1))
    return G_lin, (G_lin - 1) * nsp * H_PLANCK * Fc * Fs


def edfa(E, G_dB, NF_dB, Fc, Fs, seed=None, noise=None):
    """E*sqrt(G) + CN(0, p_noise) (devices.py:724-726).  ``noise`` overrides the RNG draw."""
    G_lin, p_noise = edfa_numbers(G_dB, NF_dB, Fc, Fs)
    if noise is None:
        noise = legacy_noise(E.shape, p_noise, seed)
    return E * np.sqrt(G_lin) + noise


def step_sizes_fixed(Lspan, hz):
    """The list of step sizes the reference's float-accumulated loop executes in fixed-step mode
    (channels.py:387, 398-403, 441) — including the degenerate last step when hz does not divide
    Lspan in binary floating point."""
    out, z = [], 0
    while z < Lspan:
        h = Lspan - z if (Lspan - z < hz) else hz
        out.append(h)
        z += h
    return out


def nlse_ssfm(Ei, cfg: FiberConfig):
    """Scalar NLSE symmetric SSFM (channels.py:201-238): per step
    ``X*L -> ifft -> e*exp(jγ|e|²hz) -> fft -> X*L`` with ``L = exp(-(α/2)(hz/2) + j(β2/2)ω²(hz/2))``."""
    E = np.asarray(Ei).reshape(len(Ei))
    n = len(E)
    w = 2 * np.pi * cfg.Fs * fftfreq(n)
    a, b2, g, hz = cfg.alpha_lin, cfg.beta2, cfg.gamma, cfg.hz
    n_spans = int(np.floor(cfg.Ltotal / cfg.Lspan))
    n_steps = int(np.floor(cfg.Lspan / hz))
    half = np.exp(-(a / 2) * (hz / 2) + 1j * (b2 / 2) * (w**2) * (hz / 2))
    for _ in range(n_spans):
        X = fft(E)
        for _ in range(n_steps):
            e = ifft(X * half)
            e = e * np.exp(1j * g * (e * np.conj(e)) * hz)
            X = fft(e) * half
        E = ifft(X)
        if cfg.amp == "edfa":
            E = edfa(E, cfg.alpha * cfg.Lspan, cfg.NF, cfg.Fc, cfg.Fs, cfg.seed)
        elif cfg.amp == "ideal":
            E = E * np.exp(a / 2 * n_steps * hz)
    return E


def manakov(Ei, cfg: FiberConfig, direction=+1, stats=None):
    """Manakov symmetric SSFM (direction=+1: channels.py:359-463) or digital back-propagation
    (direction=-1: equalization.py:1060-1171).

    Per step: first half linear step, then a fixed-point loop on the nonlinear phase
    φ = (8/9)γ(P_start + |Ex_c|² + |Ey_c|²)/2 (channels.py:493) — rotate the half-dispersed field,
    second half linear step, stop when sqrt(Σ|E_new-E_c|²)/sqrt(Σ|E_c|²) < tol (channels.py:517-519).
    Returns an (N, 2K) array, or (N, 2K*len(saveSpanN)) when span snapshots are requested.
    """
    Ei = np.asarray(Ei)
    n = Ei.shape[0]
    X = Ei[:, 0::2].T.astype(np.complex128)
    Y = Ei[:, 1::2].T.astype(np.complex128)
    w = 2 * np.pi * cfg.Fs * fftfreq(n)
    a, b2, g = cfg.alpha_lin, cfg.beta2, cfg.gamma
    sgn = 1.0 if direction > 0 else -1.0
    arg = (sgn * (-(a / 2) + 1j * (b2 / 2) * (w**2))).reshape(1, -1)  # channels.py:368 / equalization.py:1077
    n_spans = int(np.floor(cfg.Ltotal / cfg.Lspan))
    Lspan = cfg.Lspan
    save = cfg.saveSpanN
    snaps = np.zeros((n, Ei.shape[1] * len(save)), dtype=np.complex128) if save else None
    rec = 0
    n_steps = n_iters = 0

    def phase(xc, yc, P):
        return ((8 / 9) * g * (P + xc * np.conj(xc) + yc * np.conj(yc)) / 2).real

    for span in range(1, n_spans + 1):
        if direction < 0 and cfg.amp in ("edfa", "ideal"):  # equalization.py:1090-1092
            X = X * np.exp(-a / 2 * Lspan)
            Y = Y * np.exp(-a / 2 * Lspan)
        Xc, Yc = X.copy(), Y.copy()
        z = 0
        while z < Lspan:
            P = X * np.conj(X) + Y * np.conj(Y)
            phi = phase(Xc, Yc, P)
            if cfg.nlprMethod:  # channels.py:392-397
                cand = cfg.maxNlinPhaseRot / np.max(phi)
                h = cand if (Lspan - z >= cand) else (Lspan - z)
            elif Lspan - z < cfg.hz:
                h = Lspan - z
            else:
                h = cfg.hz
            lin = np.exp(arg * (h / 2))
            Xh = ifft(fft(X) * lin)
            Yh = ifft(fft(Y) * lin)
            for it in range(cfg.maxIter):
                rot = np.exp(sgn * 1j * phi * h)  # channels.py:414 / equalization.py:1129
                Xn = ifft(fft(Xh * rot) * lin)
                Yn = ifft(fft(Yh * rot) * lin)
                num = np.linalg.norm(Xn - Xc) ** 2 + np.linalg.norm(Yn - Yc) ** 2
                den = np.linalg.norm(Xc) ** 2 + np.linalg.norm(Yc) ** 2
                lim = np.sqrt(num) / np.sqrt(den)
                Xc, Yc = Xn, Yn
                n_iters += 1
                if lim < cfg.tol:
                    break
                phi = phase(Xc, Yc, P)
            X, Y = Xc.copy(), Yc.copy()
            z += h
            n_steps += 1
        if direction > 0:  # channels.py:443-451
            if cfg.amp == "edfa":
                X = edfa(X, cfg.alpha * Lspan, cfg.NF, cfg.Fc, cfg.Fs, cfg.seed)
                Y = edfa(Y, cfg.alpha * Lspan, cfg.NF, cfg.Fc, cfg.Fs, cfg.seed)
            elif cfg.amp == "ideal":
                X = X * np.exp(a / 2 * Lspan)
                Y = Y * np.exp(a / 2 * Lspan)
        if save and span in save:  # channels.py:453-456 (K=1 only, like the reference)
            snaps[:, 2 * rec:2 * rec + 1] = X.T
            snaps[:, 2 * rec + 1:2 * rec + 2] = Y.T
            rec += 1
    if stats is not None:
        stats["steps"], stats["iterations"] = n_steps, n_iters
        stats["z_last_step"] = float(h)
    if save:
        return snaps
    out = np.empty(Ei.shape, dtype=np.complex128)
    out[:, 0::2] = X.T
    out[:, 1::2] = Y.T
    return out


def manakov_nl_pass(Ehd_x, Ehd_y, Ec_x, Ec_y, Pch, gamma, hz, direction=+1):
    """One rotation of the fixed-point loop (channels.py:414-417 with φ from :493)."""
    phi = (8 / 9) * gamma * (Pch + np.abs(Ec_x) ** 2 + np.abs(Ec_y) ** 2) / 2
    rot = np.exp(direction * 1j * phi * hz)
    return Ehd_x * rot, Ehd_y * rot


def linear_fiber(Ei, L, alpha, D, Fc, Fs):
    """Linear-limit check used by the reference's TestSSFM (optic/models/channels.py:30-109):
    one all-pass/attenuation multiply in the frequency domain."""
    lam = C_KMS / Fc
    a = alpha / (10 * np.log10(np.exp(1)))
    b2 = -(D * lam**2) / (2 * np.pi * C_KMS)
    E = np.asarray(Ei)
    w = 2 * np.pi * Fs * fftfreq(E.shape[0])
    shape = (-1,) + (1,) * (E.ndim - 1)
    return ifft(fft(E, axis=0) * np.exp(-a / 2 * L + 1j * (b2 / 2) * (w**2) * L).reshape(shape), axis=0)
