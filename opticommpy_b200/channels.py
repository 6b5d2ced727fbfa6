"""Drop-in mirrors of ``optic.models.channels.ssfm`` and ``optic.models.channels.manakovSSF``.

Same call signatures, ``parameters`` attributes, defaults, default write-back into ``param``,
return values and error behaviour as the reference (optic/models/channels.py:112-249, 252-468);
the propagation itself runs on the GPU through the C-ABI (``ocb_nlse_run_host`` /
``ocb_manakov_run_host``) in complex64 with a float64-evaluated linear-operator phase.
numpy in, numpy out — like the reference's own GPU mirrors (optic/models/modelsGPU.py:404, 502).

Usage (the reference's own swap convention, examples/test_WDM_transmission.ipynb cell 3)::

    from opticommpy_b200.channels import manakovSSF, ssfm
"""
from __future__ import annotations

import ctypes as C
import logging as logg

import numpy as np
import scipy.constants as const

from . import _cabi, _engine

_AMP = {"edfa": _cabi.AMP_EDFA, "ideal": _cabi.AMP_IDEAL, None: _cabi.AMP_NONE}


def _fiber_constants(alpha, D, Fc):
    """α [1/km] and β2 [s²/km] from dB/km and ps/nm/km (channels.py:187-190, 344-347)."""
    c_kms = const.c / 1e3
    lam = c_kms / Fc
    alpha_lin = alpha / (10 * np.log10(np.exp(1)))
    beta2 = -(D * lam**2) / (2 * np.pi * c_kms)
    return alpha_lin, beta2


def _edfa_numbers(G_dB, NF_dB, Fc, Fs):
    """Gain and ASE power of the simple EDFA model (devices.py:709-722)."""
    assert G_dB > 0, "EDFA gain should be a positive scalar"
    assert NF_dB >= 3, "The minimal EDFA noise figure is 3 dB"
    NF_lin = 10 ** (NF_dB / 10)
    G_lin = 10 ** (G_dB / 10)
    nsp = (G_lin * NF_lin - 1) / (2 * (G_lin - 1))
    N_ase = (G_lin - 1) * nsp * const.h * Fc
    return G_lin, N_ase * Fs


def _require_Fs(param):
    try:
        return param.Fs
    except AttributeError:
        logg.error("Simulation sampling frequency (Fs) not provided.")
        # the reference continues and dies with an unbound-name error (channels.py:152-155, 198)
        raise NameError("name 'Fs' is not defined") from None


def _noise_setup(amp, seed, shape, noise_var, noiseRNG):
    """(noise_mode, host noise buffer or None, philox seed)."""
    if amp != "edfa":
        return _cabi.NOISE_PHILOX, None, 0
    if seed is not None and noiseRNG != "philox":
        # Reference CPU behaviour: one MT19937 realisation, reused for x, y and every span.
        w = _engine.legacy_complex_noise(shape, noise_var, seed).astype(np.complex64)
        return _cabi.NOISE_INJECTED, np.ascontiguousarray(w), 0
    key = int(seed) if seed is not None else int(np.random.SeedSequence().generate_state(2, np.uint64)[0])
    return _cabi.NOISE_PHILOX, None, key & 0xFFFFFFFFFFFFFFFF


def _manakov_params(param, direction, Nspans, n_save, *, alpha_lin, beta2, Fs, noise_var, gain_lin,
                    noise_mode, key):
    amp = param.amp
    amp_mode = _AMP[amp] if amp in _AMP else _cabi.AMP_NONE  # unknown strings: no amplification, like the reference
    return _cabi.ManakovParams(
        alpha_lin=alpha_lin, beta2=beta2, gamma=float(param.gamma), Fs=float(Fs), Lspan=float(param.Lspan),
        hz=float(param.hz), maxNlinPhaseRot=float(param.maxNlinPhaseRot), tol=float(param.tol),
        n_spans=Nspans, maxIter=int(param.maxIter), nlprMethod=int(bool(param.nlprMethod)),
        direction=direction, amp_mode=amp_mode, noise_mode=noise_mode, edfa_gain_lin=gain_lin,
        edfa_noise_var=noise_var, seed=key, n_save=n_save, reserved=0,
    )


def _manakov_engine(Ei, param, direction, *, alpha_lin, beta2, Fs, noise_var=0.0, gain_lin=1.0):
    """Shared host driver of manakovSSF (direction=+1) and manakovDBP (direction=-1)."""
    torch = _cabi.require_cuda()
    lib = _cabi.lib()
    Ei = np.asarray(Ei)
    if Ei.ndim != 2 or Ei.shape[1] % 2 != 0:
        raise IndexError("Ei must have shape (N, 2K): interleaved x/y polarisation columns")
    N, C2 = Ei.shape
    K = C2 // 2
    Nspans = int(np.floor(param.Ltotal / param.Lspan))
    saveSpanN = param.saveSpanN
    if saveSpanN and K > 1:
        # the reference fails here as well (broadcast of (N, K) into (N, 1), channels.py:454)
        raise ValueError("saveSpanN is only supported for a single pol-pair (K=1); pass saveSpanN=[]")
    hits = [s for s in range(1, Nspans + 1) if s in saveSpanN] if saveSpanN else []

    host_in = _engine.as_host_complex(Ei)
    out_dtype = np.dtype(param.prec) if saveSpanN else host_in.dtype
    if out_dtype not in (np.dtype(np.complex64), np.dtype(np.complex128)):
        out_dtype = np.dtype(np.complex128)

    noise_mode, noise_host, key = (_cabi.NOISE_PHILOX, None, 0)
    if direction > 0:
        noise_mode, noise_host, key = _noise_setup(param.amp, getattr(param, "seed", None), (K, N), noise_var,
                                                   getattr(param, "noiseRNG", "reference"))
    q = _manakov_params(param, direction, Nspans, len(hits), alpha_lin=alpha_lin, beta2=beta2, Fs=Fs,
                        noise_var=noise_var, gain_lin=gain_lin, noise_mode=noise_mode, key=key)
    stats = _cabi.ManakovStats()
    plan = _engine.get_plan(N, 2 * K)
    nblk = max(1, len(hits))
    raw_out = np.empty((nblk, N, C2), dtype=out_dtype)
    save_arr = (C.c_int32 * len(hits))(*hits) if hits else None
    _cabi.check(
        lib.ocb_manakov_run_host(
            plan.handle, host_in.ctypes.data_as(C.c_void_p), _engine.dtype_tag(host_in.dtype),
            raw_out.ctypes.data_as(C.c_void_p), _engine.dtype_tag(out_dtype), C.byref(q),
            noise_host.ctypes.data_as(C.c_void_p) if noise_host is not None else None,
            save_arr, C.byref(stats), C.c_void_p(_cabi.stream_ptr(torch)),
        ),
        "ocb_manakov_run_host",
    )
    if stats.nonconverged:
        logg.warning(
            f"Warning: target SSFM error tolerance was not achieved in {param.maxIter} iterations "
            f"({stats.nonconverged} steps)"
        )
    param._b200_stats = {"steps": int(stats.steps), "iterations": int(stats.iterations),
                         "nonconverged": int(stats.nonconverged), "last_lim": float(stats.last_lim),
                         "z_last_step": float(stats.z_last_step)}
    if saveSpanN:
        Ech = np.zeros((N, C2 * len(saveSpanN)), dtype=out_dtype)
        for i in range(len(hits)):
            Ech[:, C2 * i:C2 * (i + 1)] = raw_out[i]
    else:
        Ech = raw_out[0]
        if Ech.dtype != Ei.dtype and np.iscomplexobj(Ei):
            Ech = Ech.astype(Ei.dtype)
    return Ech


def manakov_rows_device(rows, param, direction=+1, noise_rows=None):
    """Device-resident entry: propagate planar ``rows`` (torch complex64 CUDA tensor of shape (2K, N):
    x rows then y rows) IN PLACE through ``ocb_manakov_run`` with the final field only (no snapshots).
    ``param`` needs the same attributes as manakovSSF / manakovDBP (no defaults are filled in here).
    Used by the sharded drivers and by bench.py's HBM-resident timing.  Returns the stats dict."""
    torch = _cabi.require_cuda()
    lib = _cabi.lib()
    if rows.dtype != torch.complex64 or not rows.is_cuda or not rows.is_contiguous() or rows.dim() != 2:
        raise TypeError("rows must be a contiguous CUDA complex64 tensor of shape (2K, N)")
    R, N = rows.shape
    alpha_lin, beta2 = _fiber_constants(param.alpha, param.D, param.Fc)
    gain_lin, noise_var = 1.0, 0.0
    if direction > 0 and param.amp == "edfa":
        gain_lin, noise_var = _edfa_numbers(param.alpha * param.Lspan, param.NF, param.Fc, param.Fs)
    noise_mode = _cabi.NOISE_INJECTED if noise_rows is not None else _cabi.NOISE_PHILOX
    # Noise of this entry: ``noise_rows`` (a (K, N) complex64 CUDA tensor) reproduces the reference's seeded
    # CPU semantics (one realisation shared by x, y and every span); otherwise on-device Philox streams keyed by
    # ``param.seed`` — and, like manakovSSF, by a fresh random key when the seed is None, so that Monte-Carlo
    # calls without a seed are independent.  ``noiseRNG`` does not apply here (no host-side MT19937 stream).
    seed = getattr(param, "seed", None)
    key = (int(seed) if seed is not None
           else int(np.random.SeedSequence().generate_state(1, np.uint64)[0])) & 0xFFFFFFFFFFFFFFFF
    Nspans = int(np.floor(param.Ltotal / param.Lspan))
    q = _manakov_params(param, direction, Nspans, 0, alpha_lin=alpha_lin, beta2=beta2, Fs=param.Fs,
                        noise_var=noise_var, gain_lin=gain_lin, noise_mode=noise_mode, key=key)
    stats = _cabi.ManakovStats()
    with torch.cuda.device(rows.device):  # plan, kernels and stream all belong to the device that owns `rows`
        plan = _engine.get_plan(N, R, rows.device.index)
        _cabi.check(
            lib.ocb_manakov_run(plan.handle, C.c_void_p(rows.data_ptr()), C.byref(q),
                                C.c_void_p(noise_rows.data_ptr()) if noise_rows is not None else None,
                                None, None, C.byref(stats), C.c_void_p(_cabi.stream_ptr(torch))),
            "ocb_manakov_run",
        )
    return {"steps": int(stats.steps), "iterations": int(stats.iterations), "nonconverged": int(stats.nonconverged),
            "last_lim": float(stats.last_lim), "z_last_step": float(stats.z_last_step)}


def manakovSSF(Ei, param):
    """
    Run the Manakov split-step Fourier model (symmetric, dual-pol.) on the GPU.

    Parameters and defaults are those of the reference (channels.py:263-281): Ltotal [400],
    Lspan [80], hz [0.5], alpha [0.2], D [16], gamma [1.3], Fc [193.1e12], Fs, prec
    [np.complex128], amp ['edfa'], NF [4.5], maxIter [10], tol [1e-5], nlprMethod [True],
    maxNlinPhaseRot [2e-2], prgsBar [True], saveSpanN [[Ltotal//Lspan]], seed [None],
    returnParameters [False].  Extra (optional): ``noiseRNG`` = 'reference' (default: with a
    seed, the reference's single MT19937 realisation shared by x/y/all spans) or 'philox'
    (independent on-device streams).

    Returns ``Ech`` (and ``param`` when ``returnParameters``).
    """
    Fs = _require_Fs(param)

    # defaults are written back into the caller's object, like the reference (channels.py:305-322)
    param.Ltotal = getattr(param, "Ltotal", 400)
    param.Lspan = getattr(param, "Lspan", 80)
    param.hz = getattr(param, "hz", 0.5)
    param.alpha = getattr(param, "alpha", 0.2)
    param.D = getattr(param, "D", 16)
    param.gamma = getattr(param, "gamma", 1.3)
    param.Fc = getattr(param, "Fc", 193.1e12)
    param.prec = getattr(param, "prec", np.complex128)
    param.amp = getattr(param, "amp", "edfa")
    param.NF = getattr(param, "NF", 4.5)
    param.maxIter = getattr(param, "maxIter", 10)
    param.tol = getattr(param, "tol", 1e-5)
    param.nlprMethod = getattr(param, "nlprMethod", True)
    param.maxNlinPhaseRot = getattr(param, "maxNlinPhaseRot", 2e-2)
    param.seed = getattr(param, "seed", None)
    param.prgsBar = getattr(param, "prgsBar", True)
    param.saveSpanN = getattr(param, "saveSpanN", [param.Ltotal // param.Lspan])
    param.returnParameters = getattr(param, "returnParameters", False)

    alpha_lin, beta2 = _fiber_constants(param.alpha, param.D, param.Fc)
    gain_lin, noise_var = 1.0, 0.0
    if param.amp == "edfa":
        gain_lin, noise_var = _edfa_numbers(param.alpha * param.Lspan, param.NF, param.Fc, Fs)

    logg.info("Running Manakov SSF model on GPU (B200 native)...")
    Ech = _manakov_engine(Ei, param, +1, alpha_lin=alpha_lin, beta2=beta2, Fs=Fs,
                          noise_var=noise_var, gain_lin=gain_lin)
    return (Ech, param) if param.returnParameters else Ech


def ssfm(Ei, param=None):
    """
    Split-step Fourier method (symmetric, single-pol.) on the GPU.

    Parameters and defaults as in the reference (channels.py:123-136): Ltotal [400], Lspan [80],
    hz [0.5], alpha [0.2], D [16], gamma [1.3], Fc [193.1e12], Fs, prec [np.complex128],
    amp ['edfa'], NF [4.5], seed [None], prgsBar [True], returnParameters [False].
    """
    Fs = _require_Fs(param)

    param.Ltotal = getattr(param, "Ltotal", 400)
    param.Lspan = getattr(param, "Lspan", 80)
    param.hz = getattr(param, "hz", 0.5)
    param.alpha = getattr(param, "alpha", 0.2)
    param.D = getattr(param, "D", 16)
    param.gamma = getattr(param, "gamma", 1.3)
    param.Fc = getattr(param, "Fc", 193.1e12)
    param.prec = getattr(param, "prec", np.complex128)
    param.amp = getattr(param, "amp", "edfa")
    param.NF = getattr(param, "NF", 4.5)
    param.seed = getattr(param, "seed", None)
    param.prgsBar = getattr(param, "prgsBar", True)
    param.returnParameters = getattr(param, "returnParameters", False)

    torch = _cabi.require_cuda()
    lib = _cabi.lib()
    alpha_lin, beta2 = _fiber_constants(param.alpha, param.D, param.Fc)
    gain_lin, noise_var = 1.0, 0.0
    if param.amp == "edfa":
        gain_lin, noise_var = _edfa_numbers(param.alpha * param.Lspan, param.NF, param.Fc, Fs)

    Ei = np.asarray(Ei)
    N = len(Ei)
    host_in = _engine.as_host_complex(Ei.reshape(N))  # channels.py:208-210
    Nspans = int(np.floor(param.Ltotal / param.Lspan))  # channels.py:205
    Nsteps = int(np.floor(param.Lspan / param.hz))      # channels.py:206

    noise_mode, noise_host, key = _noise_setup(param.amp, param.seed, (N,), noise_var,
                                               getattr(param, "noiseRNG", "reference"))
    q = _cabi.NlseParams(
        alpha_lin=alpha_lin, beta2=beta2, gamma=float(param.gamma), Fs=float(Fs), hz=float(param.hz),
        n_spans=Nspans, n_steps=Nsteps, amp_mode=_AMP.get(param.amp, _cabi.AMP_NONE), noise_mode=noise_mode,
        edfa_gain_lin=gain_lin, edfa_noise_var=noise_var, seed=key,
    )
    plan = _engine.get_plan(N, 1)
    out = np.empty(N, dtype=host_in.dtype)
    _cabi.check(
        lib.ocb_nlse_run_host(
            plan.handle, host_in.ctypes.data_as(C.c_void_p), _engine.dtype_tag(host_in.dtype),
            out.ctypes.data_as(C.c_void_p), _engine.dtype_tag(out.dtype), C.byref(q),
            noise_host.ctypes.data_as(C.c_void_p) if noise_host is not None else None,
            C.c_void_p(_cabi.stream_ptr(torch)),
        ),
        "ocb_nlse_run_host",
    )
    return (out, param) if param.returnParameters else out
