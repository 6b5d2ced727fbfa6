"""Drop-in mirrors of ``optic.dsp.equalization.edc``, ``mimoAdaptEqualizer`` and ``manakovDBP``.

Signatures, parameter names/defaults, return shapes/dtypes and error behaviour follow the
reference (optic/dsp/equalization.py:36-122, 125-351, 976-1173).  All arithmetic runs on the GPU
through the C-ABI (``ocb_edc_run``, ``ocb_mimo_eq_run``, ``ocb_manakov_run_host``).
"""
from __future__ import annotations

import ctypes as C
import logging as logg

import numpy as np
import scipy.constants as const
from numpy.fft import fftfreq, fftshift, ifft

from . import _cabi, _engine
from .channels import _fiber_constants, _manakov_engine, _require_Fs
from .modulation import normalizedConstellation

_vp = C.c_void_p


def _ptr(t, byte_offset: int = 0):
    return _vp(t.data_ptr() + byte_offset)


def _to_device(torch, a: np.ndarray):
    """H2D copy of a numpy array into a torch tensor (device-memory container)."""
    return torch.from_numpy(np.ascontiguousarray(a)).to("cuda", non_blocking=False)


# ------------------------------------------------------------------------------------------------
def _edc_taps(param, Fs):
    """Time-domain taps of the EDC filter exactly as the reference builds them (equalization.py:90-105 and
    core.py:1016): returns (h complex64 of K taps, K, Nfft)."""
    L = getattr(param, "L", 50)
    D = getattr(param, "D", 16)
    Fc = getattr(param, "Fc", 193.1e12)
    Rs = getattr(param, "Rs", 32e9)
    NfilterCoeffs = getattr(param, "NfilterCoeffs", None)
    Nfft = getattr(param, "Nfft", None)

    _, beta2 = _fiber_constants(0.0, D, Fc)
    if NfilterCoeffs is None:  # equalization.py:96-97
        NfilterCoeffs = int(2 * np.ceil(6.67 * np.abs(beta2) * L * Rs**2 * (Fs / Rs)))
    if Nfft is None:  # equalization.py:100-101
        Nfft = 2 ** int(np.ceil(np.log2(NfilterCoeffs)))
    if Nfft < NfilterCoeffs:
        logg.error("FFT size is smaller than filter length")
        raise NameError("name 'd' is not defined")  # core.py:1009-1012 leaves d unbound
    K = int(NfilterCoeffs)
    if K < 1:
        raise ValueError("EDC filter needs at least one coefficient")
    w = 2 * np.pi * Fs * fftfreq(K)
    H = np.exp(-1j * (beta2 / 2) * (w**2) * L)        # equalization.py:103-105
    h = fftshift(ifft(H)).astype(np.complex64)        # core.py:1016 (time-domain taps, float64 math)
    return h, K, Nfft


def edc_rows_device(d_x, d_y, h):
    """Device-resident EDC / FIR stage: planar rows ``d_x`` (nModes, N, 2) float32 CUDA -> ``d_y`` (same shape) through
    ``ocb_edc_run`` with the complex64 taps ``h`` (host array).  No host synchronisation; returns the temporaries that
    must stay alive until the stream has consumed them."""
    torch = _cabi.require_cuda()
    lib = _cabi.lib()
    st = _vp(_cabi.stream_ptr(torch))
    nModes, Nsig = int(d_x.shape[0]), int(d_x.shape[1])
    K = len(h)
    d_h = _to_device(torch, np.ascontiguousarray(h).view(np.float32))
    ws_bytes = int(lib.ocb_edc_workspace_bytes(Nsig, nModes, K))
    d_ws = torch.empty(ws_bytes + 256, dtype=torch.uint8, device="cuda")
    ws_ptr = (d_ws.data_ptr() + 255) // 256 * 256
    _cabi.check(lib.ocb_edc_run(_ptr(d_x), _ptr(d_y), Nsig, nModes, _ptr(d_h), K, _vp(ws_ptr), ws_bytes, st),
                "ocb_edc_run")
    return d_h, d_ws


def edc(sigIn, param):
    """
    Electronic chromatic dispersion compensation (EDC) on the GPU.

    Parameters as in the reference (equalization.py:47-53): L [50 km], D [16 ps/nm/km],
    Fc [193.1e12 Hz], Fs, Rs [32e9], NfilterCoeffs [derived], Nfft [derived].
    The output is ``conv(x, h)[D : D+len(x)]`` with ``h = fftshift(ifft(H))`` exactly as
    ``blockwiseFFTConv(..., freqDomainFilter=True)`` computes it (core.py:1016-1044); the GPU picks
    its own overlap-save block size, which does not change the result of a linear convolution.
    """
    Fs = _require_Fs(param)
    sigIn = np.asarray(sigIn)
    try:
        nModes = sigIn.shape[1]
        input1D = False
    except IndexError:
        nModes = 1
        sigIn = sigIn.reshape(sigIn.size, nModes)
        input1D = True

    h, K, Nfft = _edc_taps(param, Fs)

    logg.info("Running CD compensation...")
    logg.info(f"CD filter length: {K} taps, FFT size: {Nfft}")

    torch = _cabi.require_cuda()
    lib = _cabi.lib()
    st = _vp(_cabi.stream_ptr(torch))
    Nsig = sigIn.shape[0]
    is_complex = np.iscomplexobj(sigIn)
    host = _engine.as_host_complex(sigIn)
    d_raw = _to_device(torch, host.view(np.float32 if host.dtype == np.complex64 else np.float64))
    d_x = torch.empty((nModes, Nsig, 2), dtype=torch.float32, device="cuda")
    d_y = torch.empty_like(d_x)
    _cabi.check(lib.ocb_pack_fields(_ptr(d_raw), _engine.dtype_tag(host.dtype), Nsig, nModes, 0, _ptr(d_x), st),
                "ocb_pack_fields")
    keep = edc_rows_device(d_x, d_y, h)
    _cabi.check(lib.ocb_unpack_fields(_ptr(d_y), Nsig, nModes, 0, _ptr(d_raw), _engine.dtype_tag(host.dtype), st),
                "ocb_unpack_fields")
    out = d_raw.cpu().numpy().view(host.dtype).reshape(Nsig, nModes)
    if is_complex:
        sigOut = out.astype(sigIn.dtype, copy=False)
    else:
        sigOut = out.real.astype(sigIn.dtype)  # equalization.py:110 + core.py:1043-1046
    if input1D:
        sigOut = sigOut.flatten()
    return sigOut


# ------------------------------------------------------------------------------------------------
class _EqSetup:
    """Parsed/validated arguments of one mimoAdaptEqualizer call (equalization.py:180-257)."""


def _parse_equalizer_args(sigIn, param, symbRef):
    s = _EqSetup()
    if symbRef is None:
        symbRef = []
    if param is None:
        param = []
    s.numIter = getattr(param, "numIter", 1)
    s.nTaps = getattr(param, "nTaps", 15)
    mu = getattr(param, "mu", [1e-3])
    s.lambdaRLS = getattr(param, "lambdaRLS", 0.99)
    s.SpS = getattr(param, "SpS", 2)
    H = getattr(param, "H", [])
    H_ = getattr(param, "H_", [])
    L = getattr(param, "L", [])
    s.storeCoeff = getattr(param, "storeCoeff", False)
    s.runWL = getattr(param, "runWL", False)
    s.alg = getattr(param, "alg", ["nlms"])
    constType = getattr(param, "constType", "qam")
    M = getattr(param, "M", 4)
    shapingFactor = getattr(param, "shapingFactor", 0)
    s.returnResults = getattr(param, "returnResults", False)
    s.prec = getattr(param, "prec", np.complex64)

    sigIn = np.asarray(sigIn)
    if not len(symbRef):
        symbRef = sigIn  # the reference copies; nothing here mutates either array
    symbRef = np.asarray(symbRef)
    try:
        if sigIn.shape[1] > sigIn.shape[0]:
            sigIn = sigIn.T
        s.input1D = False
    except IndexError:
        sigIn = sigIn.reshape(len(sigIn), 1)
        s.input1D = True
    try:
        if symbRef.shape[1] > symbRef.shape[0]:
            symbRef = symbRef.T
    except IndexError:
        symbRef = symbRef.reshape(len(symbRef), 1)
    s.nModes = int(sigIn.shape[1])

    # The casts to prec=complex64 (:222-223) and the zero padding of floor(nTaps/2) rows (:227-231) are
    # done on the device: the raw arrays are uploaded as they are (numpy's complex astype is slow).
    needs_ref = any(a in ("nlms", "da-rde", "rls") for a in (s.alg if isinstance(s.alg, list) else []))
    s.symbRef = _engine.as_host_complex(symbRef) if needs_ref else None
    s.sig = _engine.as_host_complex(sigIn)
    s.mu = np.atleast_1d(np.array(mu).astype(np.float32))

    Lpad = int(np.floor(s.nTaps / 2))
    s.Lpad = Lpad
    s.nPad = s.sig.shape[0] + 2 * Lpad

    s.constSymb = normalizedConstellation(M, constType, shapingFactor, np.complex64)  # :234-241
    s.totalNumSymb = int(np.fix((s.nPad - s.nTaps) / s.SpS + 1))  # :243

    if isinstance(L, np.ndarray):
        L = L.tolist()
    if not L:
        L = [s.totalNumSymb]
    s.L = [int(v) for v in (L if isinstance(L, (list, tuple)) else [L])]

    if isinstance(H, np.ndarray) and H.size:
        s.H = np.ascontiguousarray(H.astype(np.complex64))
    elif not isinstance(H, np.ndarray) and H:
        s.H = np.ascontiguousarray(np.asarray(H).astype(np.complex64))
    else:  # centre-spike initialisation (:249-255)
        s.H = np.zeros((s.nModes**2, s.nTaps), dtype=np.complex64)
        for i in range(s.nModes):
            s.H[i + i * s.nModes, int(np.floor(s.nTaps / 2))] = 1 + 0j
    if isinstance(H_, np.ndarray) and H_.size:
        s.H_ = np.ascontiguousarray(H_.astype(np.complex64))
    elif not isinstance(H_, np.ndarray) and H_:
        s.H_ = np.ascontiguousarray(np.asarray(H_).astype(np.complex64))
    else:
        s.H_ = np.zeros((s.nModes**2, s.nTaps), dtype=np.complex64)

    if not isinstance(s.alg, list):
        # the reference's scalar-alg branch cannot run (4-of-5 unpack + numba typing error, :320-339)
        raise TypeError("param.alg must be a list of algorithm names, e.g. ['cma', 'rde']")
    for a in s.alg:
        if a in ("rls", "dd-rls"):
            if s.nTaps > 64:
                raise NotImplementedError("'rls'/'dd-rls' on the GPU keep at most two matrix rows per lane: nTaps <= 64")
            if s.runWL:
                raise NotImplementedError("'rls'/'dd-rls' have no widely-linear update in the reference (rlsUp ignores H_)")
            continue
        if a not in _cabi.ALG_IDS:
            raise ValueError("Equalization algorithm not specified (or incorrectly specified).")
    if len(s.L) < len(s.alg) or len(s.mu) < len(s.alg):
        raise IndexError("list index out of range: L and mu need one entry per algorithm stage")
    # radii used by cma / rde (:453-456)
    s.Rcma = float(np.mean(np.abs(s.constSymb) ** 4) / np.mean(np.abs(s.constSymb) ** 2))
    s.Rrde = np.unique(np.abs(s.constSymb)).astype(np.float32)
    return s


def _run_equalizer_batch(setups):
    """Run a batch of independent streams (identical geometry) on the device, stage by stage."""
    torch = _cabi.require_cuda()
    lib = _cabi.lib()
    st = _vp(_cabi.stream_ptr(torch))
    s0 = setups[0]
    nS, nM, nT, SpS = len(setups), s0.nModes, s0.nTaps, s0.SpS
    nSamp = s0.nPad
    for s in setups[1:]:
        if (s.nModes, s.nTaps, s.SpS, s.nPad, s.alg, s.L, s.numIter, s.runWL) != \
           (nM, nT, SpS, nSamp, s0.alg, s0.L, s0.numIter, s0.runWL) or not np.array_equal(s.mu, s0.mu):
            raise ValueError("all streams of a batch must share geometry, stages and step sizes")
    total = s0.totalNumSymb
    has_ref = all(s.symbRef is not None for s in setups)
    Lref = min(s.symbRef.shape[0] for s in setups) if has_ref else 0

    def upload_c64(arr, dst_view):
        """raw H2D copy of a complex64/128 host array, converted to complex64 into dst_view (device)"""
        raw = torch.from_numpy(arr.view(np.float32 if arr.dtype == np.complex64 else np.float64)).to("cuda")
        _cabi.check(lib.ocb_cast_complex(_ptr(raw), _engine.dtype_tag(arr.dtype), _vp(dst_view.data_ptr()),
                                         _cabi.OCB_C64, arr.size, st), "ocb_cast_complex")
        return raw  # keep alive until the stream has consumed it

    def upload_stack(arrs, dst):
        """ONE pinned staging buffer and ONE H2D copy for the whole batch (the streams share shape and dtype), one
        cast kernel to complex64, then a device-side strided copy into the per-stream slices of ``dst``."""
        dt = arrs[0].dtype
        n0 = arrs[0].shape[0]
        if any(a.dtype != dt or a.shape != arrs[0].shape for a in arrs):
            return [upload_c64(a, dst[i]) for i, a in enumerate(arrs)]
        ft = np.float32 if dt == np.complex64 else np.float64
        stage = torch.empty((len(arrs), n0, nM, 2), dtype=torch.float32 if dt == np.complex64 else torch.float64,
                            pin_memory=True)
        view = stage.numpy()
        for i, a in enumerate(arrs):
            view[i] = a.view(ft).reshape(n0, nM, 2)
        raw = stage.to("cuda", non_blocking=True)
        if dt == np.complex64:
            dst.copy_(raw)
            return [stage, raw]
        c64 = torch.empty((len(arrs), n0, nM, 2), dtype=torch.float32, device="cuda")
        _cabi.check(lib.ocb_cast_complex(_ptr(raw), _cabi.OCB_C128, _ptr(c64), _cabi.OCB_C64, len(arrs) * n0 * nM, st),
                    "ocb_cast_complex")
        dst.copy_(c64)
        return [stage, raw, c64]

    d_x = torch.zeros((nS, nSamp, nM, 2), dtype=torch.float32, device="cuda")  # zero rows = the padding
    nsig = s0.sig.shape[0]
    keep = upload_stack([s.sig for s in setups], d_x[:, s0.Lpad:s0.Lpad + nsig])
    if has_ref:
        d_ref = torch.empty((nS, Lref, nM, 2), dtype=torch.float32, device="cuda")
        keep += upload_stack([np.ascontiguousarray(s.symbRef[:Lref]) for s in setups], d_ref)
    else:
        d_ref = None
    d_H = _to_device(torch, np.stack([s.H for s in setups]).view(np.float32))
    d_Hw = _to_device(torch, np.stack([s.H_ for s in setups]).view(np.float32)) if s0.runWL else None
    d_y, d_e, hiter = equalizer_stages_device(s0, nS, d_x, d_ref, Lref, d_H, d_Hw)
    return _equalizer_download(setups, d_y, d_e, d_H, d_Hw, hiter)


def equalizer_stages_device(s0, nS, d_x, d_ref, Lref, d_H, d_Hw):
    """Device-resident equalizer: run every stage of ``s0.alg`` on ``d_x`` (nS, nPad, nModes, 2) float32 CUDA (already
    zero-padded by floor(nTaps/2) rows on both ends) with taps ``d_H`` / ``d_Hw`` updated in place.  Returns
    (d_y (nS, totalNumSymb, nModes, 2), d_e (nS, nModes, totalNumSymb), Hiter tensor or None).  No host sync."""
    torch = _cabi.require_cuda()
    lib = _cabi.lib()
    st = _vp(_cabi.stream_ptr(torch))
    nM, nT, SpS = s0.nModes, s0.nTaps, s0.SpS
    nSamp = int(d_x.shape[1])
    total = s0.totalNumSymb
    d_c = _to_device(torch, s0.constSymb.view(np.float32))
    d_r = _to_device(torch, s0.Rrde)
    d_y = torch.zeros((nS, total, nM, 2), dtype=torch.float32, device="cuda")
    d_e = torch.zeros((nS, nM, total), dtype=torch.float32, device="cuda")

    hiter = None
    nStart = 0
    for stage, name in enumerate(s0.alg):
        Ls = s0.L[stage]
        nEnd = nStart + Ls
        if nEnd > total or (Ls > 0 and (nEnd - 1) * SpS + nT > nSamp):
            raise IndexError("training section runs past the end of the signal")
        if name in ("nlms", "da-rde", "rls") and nEnd > Lref:
            raise IndexError("reference symbol sequence shorter than the training section")
        reps = s0.numIter if stage == 0 else 1
        for rep in range(reps):
            # a stage is a pointer offset into the full buffers; streams are strided by full lengths
            d_hit = None
            if s0.storeCoeff:
                d_hit = torch.empty((nS, Ls, nM * nM, nT, 2), dtype=torch.float32, device="cuda")
            if name in ("rls", "dd-rls"):
                ws_bytes = int(lib.ocb_mimo_eq_rls_workspace_bytes(nS, nM, Ls, nT))
                d_ws = torch.empty(ws_bytes + 256, dtype=torch.uint8, device="cuda")
                ws_ptr = (d_ws.data_ptr() + 255) // 256 * 256
                lam = float(np.float32(s0.lambdaRLS))  # the reference casts lambda to prec (:225)
                _cabi.check(
                    lib.ocb_mimo_eq_rls_run(
                        _ptr(d_x, nStart * SpS * nM * 8), _ptr(d_ref, nStart * nM * 8) if d_ref is not None else None,
                        _ptr(d_H), _ptr(d_y, nStart * nM * 8), _ptr(d_e, nStart * 4),
                        _ptr(d_hit) if d_hit is not None else None,
                        nS, nSamp - nStart * SpS, nSamp * nM, Lref * nM, total * nM, nM * total, total,
                        Ls, nM, nT, SpS, int(name == "dd-rls"), lam, _ptr(d_c), len(s0.constSymb),
                        _vp(ws_ptr), ws_bytes, st,
                    ),
                    "ocb_mimo_eq_rls_run",
                )
                hiter = d_hit
                continue
            _cabi.check(
                lib.ocb_mimo_eq_run(
                    _ptr(d_x, nStart * SpS * nM * 8), _ptr(d_ref, nStart * nM * 8) if d_ref is not None else None, _ptr(d_H),
                    _ptr(d_Hw) if d_Hw is not None else None, _ptr(d_y, nStart * nM * 8), _ptr(d_e, nStart * 4),
                    _ptr(d_hit) if d_hit is not None else None,
                    nS, nSamp - nStart * SpS, nSamp * nM, Lref * nM, total * nM, nM * total, total,
                    Ls, nM, nT, SpS, _cabi.ALG_IDS[name], float(s0.mu[stage]),
                    _ptr(d_c), len(s0.constSymb), _ptr(d_r), len(s0.Rrde), s0.Rcma, int(bool(s0.runWL)), st,
                ),
                "ocb_mimo_eq_run",
            )
            hiter = d_hit
        nStart = nEnd
    return d_y, d_e, hiter


def _equalizer_download(setups, d_y, d_e, d_H, d_Hw, hiter):
    s0 = setups[0]
    nS, nM, nT, total = len(setups), s0.nModes, s0.nTaps, s0.totalNumSymb
    y = d_y.cpu().numpy().view(np.complex64).reshape(nS, total, nM)
    e = d_e.cpu().numpy()
    H = d_H.cpu().numpy().view(np.complex64).reshape(nS, nM * nM, nT)
    Hw = d_Hw.cpu().numpy().view(np.complex64).reshape(nS, nM * nM, nT) if d_Hw is not None else None
    if hiter is not None:
        hit = hiter.cpu().numpy().view(np.complex64).reshape(nS, -1, nM * nM, nT).transpose(0, 2, 3, 1)
    else:
        hit = H[..., None]
    results = []
    for i, s in enumerate(setups):
        sigOut = y[i].astype(s.prec, copy=False)
        errSq = e[i].astype(s.prec)  # the reference stores |err|² in a prec-typed array (:263)
        Hi = H[i].astype(s.prec, copy=False)
        Hit = np.ascontiguousarray(hit[i]).astype(s.prec, copy=False)
        if s.input1D:
            sigOut = sigOut.flatten()
        if s.returnResults:
            if s.runWL:
                results.append((sigOut, Hi, Hw[i].astype(s.prec, copy=False), errSq, Hit))
            else:
                results.append((sigOut, Hi, errSq, Hit))
        else:
            results.append(sigOut)
    return results


def mimoAdaptEqualizer(sigIn, param=None, symbRef=None):
    """
    General N x N MIMO adaptive equalizer (fractionally spaced FIR) on the GPU.

    Parameters as in the reference (equalization.py:138-153): numIter [1], nTaps [15], mu [[1e-3]],
    lambdaRLS [0.99], SpS [2], H [[]], H_ [[]], L [[]], storeCoeff [False], runWL [False],
    alg [['nlms']], constType ['qam'], M [4], shapingFactor [0], prgsBar [True],
    returnResults [False], prec [np.complex64].

    Algorithms: 'cma', 'rde', 'nlms', 'dd-lms', 'da-rde', 'rls', 'dd-rls', 'static' (list form only; one
    entry of ``L`` and ``mu`` per stage; 'rls'/'dd-rls': nTaps <= 64, forgetting factor ``lambdaRLS``, the inverse
    correlation matrices start from the identity in every stage).  Returns ``sigOut`` or, with ``returnResults``,
    ``(sigOut, H, errSq, Hiter)`` / ``(sigOut, H, H_, errSq, Hiter)`` in widely-linear mode.
    """
    logg.info("Running adaptive equalizer...")
    return _run_equalizer_batch([_parse_equalizer_args(sigIn, param, symbRef)])[0]


def mimoAdaptEqualizerBatch(sigIns, param=None, symbRefs=None):
    """Equalize several independent streams (WDM channels / Monte-Carlo realisations) in ONE
    launch per stage: one persistent warp per stream.  Every stream gets exactly the result of
    ``mimoAdaptEqualizer(sigIns[i], param, symbRefs[i])``."""
    if symbRefs is None:
        symbRefs = [None] * len(sigIns)
    return _run_equalizer_batch([_parse_equalizer_args(x, param, r) for x, r in zip(sigIns, symbRefs)])


# ------------------------------------------------------------------------------------------------
def manakovDBP(Ei, param):
    """
    Manakov SSF digital back-propagation (symmetric, dual-pol.) on the GPU.

    Parameters/defaults as in the reference (equalization.py:987-1003): Ltotal [400], Lspan [80],
    hz [0.5], alpha [0.2], D [16], gamma [1.3], Fc [193.1e12], Fs, prec [np.complex128],
    amp ['edfa'], maxIter [10], tol [1e-5], nlprMethod [True], maxNlinPhaseRot [2e-2],
    prgsBar [True], saveSpanN [[Ltotal//Lspan]], returnParameters [False].
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
    param.maxIter = getattr(param, "maxIter", 10)
    param.tol = getattr(param, "tol", 1e-5)
    param.nlprMethod = getattr(param, "nlprMethod", True)
    param.maxNlinPhaseRot = getattr(param, "maxNlinPhaseRot", 2e-2)
    param.prgsBar = getattr(param, "prgsBar", True)
    param.saveSpanN = getattr(param, "saveSpanN", [param.Ltotal // param.Lspan])
    param.returnParameters = getattr(param, "returnParameters", False)

    alpha_lin, beta2 = _fiber_constants(param.alpha, param.D, param.Fc)
    Ech = _manakov_engine(Ei, param, -1, alpha_lin=alpha_lin, beta2=beta2, Fs=Fs)
    return (Ech, param) if param.returnParameters else Ech
