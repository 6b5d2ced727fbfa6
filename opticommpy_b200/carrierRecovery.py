"""Drop-in mirrors of ``optic.dsp.carrierRecovery.bps`` and the ``cpr`` wrapper around it.

``bps`` (carrierRecovery.py:172-223) runs on the GPU through ``ocb_bps_run`` in float64, like the
reference.  ``cpr`` (carrierRecovery.py:37-169) keeps the reference's parameter names and defaults and
runs the whole wrapper — 4th-power FOE, power normalisation, bps, ``unwrap(4φ)/4``, de-rotation — in one
device call (``ocb_cpr_bps_run``); only ``alg='bps'`` (and its alias ``'bpsGPU'``, the reference's own GPU
selector) is part of this hot path.
"""
from __future__ import annotations

import ctypes as C
import logging as logg

import numpy as np

from . import _cabi
from .modulation import grayMapping

_vp = C.c_void_p


def bps(sigIn, N, constSymb, B, returnIndex=False):
    """
    Blind phase search (BPS) on the GPU.

    Parameters
    ----------
    sigIn : (L, nModes) complex array of received symbols.
    N : half-length of the 2N+1 averaging window.
    constSymb : complex constellation.
    B : number of test phases b*(pi/2)/B.
    returnIndex : also return the int32 argmin indices (extension, default False).

    Returns
    -------
    phaseEst : (L, nModes) float64, each entry one of the B test phases.
    """
    torch = _cabi.require_cuda()
    lib = _cabi.lib()
    from . import _engine
    x = _engine.as_host_complex(np.asarray(sigIn))  # uploaded raw; widened to complex128 on the device
    if x.ndim != 2:
        raise IndexError("bps expects a 2-D (symbols, modes) array")  # sigIn.shape[1], :197
    L, nModes = x.shape
    c = np.ascontiguousarray(np.asarray(constSymb).astype(np.complex128))
    st = _vp(_cabi.stream_ptr(torch))
    if x.dtype == np.complex128:
        d_x = torch.from_numpy(x.view(np.float64)).to("cuda")
    else:
        d_raw = torch.from_numpy(x.view(np.float32)).to("cuda")
        d_x = torch.empty((L, nModes, 2), dtype=torch.float64, device="cuda")
        _cabi.check(lib.ocb_cast_complex(_vp(d_raw.data_ptr()), _cabi.OCB_C64, _vp(d_x.data_ptr()), _cabi.OCB_C128,
                                         x.size, st), "ocb_cast_complex")
    d_c = torch.from_numpy(c.view(np.float64)).to("cuda")
    d_idx = torch.empty((L, nModes), dtype=torch.int32, device="cuda")
    d_ph = torch.empty((L, nModes), dtype=torch.float64, device="cuda")
    _cabi.check(
        lib.ocb_bps_run(_vp(d_x.data_ptr()), L, nModes, _vp(d_c.data_ptr()), len(c), int(B), int(N),
                        _vp(d_idx.data_ptr()), _vp(d_ph.data_ptr()), _vp(_cabi.stream_ptr(torch))),
        "ocb_bps_run",
    )
    phaseEst = d_ph.cpu().numpy()
    if returnIndex:
        return phaseEst, d_idx.cpu().numpy()
    return phaseEst


def cpr_bps_device(d_x, x_dtype_tag, L, nModes, constSymb, B, N, runFOE, Fs, foeM, want_fo=False):
    """Device-resident cpr(alg='bps'): ``d_x`` holds (L, nModes) interleaved complex samples on the GPU (complex64 or
    complex128 as float pairs, ``x_dtype_tag``).  Returns (d_y (L, nModes, 2) float64, d_phase (L, nModes) float64,
    fo ctypes array or None, temporaries to keep alive).  Synchronises only when the frequency offsets are requested."""
    torch = _cabi.require_cuda()
    lib = _cabi.lib()
    st = _vp(_cabi.stream_ptr(torch))
    c128 = np.ascontiguousarray(np.asarray(constSymb).astype(np.complex128))
    d_c = torch.from_numpy(c128.view(np.float64)).to("cuda")
    d_y = torch.empty((L, nModes, 2), dtype=torch.float64, device="cuda")
    d_ph = torch.empty((L, nModes), dtype=torch.float64, device="cuda")
    ws_bytes = int(lib.ocb_cpr_workspace_bytes(L, nModes))
    d_ws = torch.empty(ws_bytes + 256, dtype=torch.uint8, device="cuda")
    ws_ptr = (d_ws.data_ptr() + 255) // 256 * 256
    fo = (C.c_double * nModes)() if want_fo else None
    _cabi.check(
        lib.ocb_cpr_bps_run(_vp(d_x.data_ptr()), x_dtype_tag, L, nModes, _vp(d_c.data_ptr()), len(c128),
                            int(B), int(N // 2), int(bool(runFOE)), float(Fs), int(foeM), _vp(d_y.data_ptr()),
                            _vp(d_ph.data_ptr()), fo, _vp(ws_ptr), ws_bytes, st),
        "ocb_cpr_bps_run",
    )
    return d_y, d_ph, fo, (d_c, d_ws)


def cpr(sigIn, param=None, symbTx=None):
    """
    Carrier phase recovery (CPR) with the BPS algorithm on the GPU.

    Parameters as in the reference (carrierRecovery.py:48-58): alg ['bps'], M [4], constType
    ['qam'], shapingFactor [0], B [64], N [35], Ts [1/32e9], runFOE [True], returnPhases [False].
    Returns ``sigOut`` or ``(sigOut, phaseEst)``.
    """
    if param is None:
        param = []
    alg = getattr(param, "alg", "bps")
    M = getattr(param, "M", 4)
    constType = getattr(param, "constType", "qam")
    shapingFactor = getattr(param, "shapingFactor", 0)
    B = getattr(param, "B", 64)
    N = getattr(param, "N", 35)
    Ts = getattr(param, "Ts", 1 / 32e9)
    runFOE = getattr(param, "runFOE", True)
    returnPhases = getattr(param, "returnPhases", False)

    sigIn = np.asarray(sigIn)
    try:
        sigIn.shape[1]
        input1D = False
    except IndexError:
        sigIn = sigIn.reshape(len(sigIn), 1)
        input1D = True

    constSymb = grayMapping(M, constType)  # complex64, like the reference (:118-121)
    px = np.exp(-shapingFactor * np.abs(constSymb) ** 2)
    px = px / np.sum(px)
    constSymb /= np.sqrt(np.sum(np.abs(constSymb) ** 2 * px))

    if alg in ("ddpll", "viterbi"):
        raise NotImplementedError(f"cpr alg '{alg}' is outside the B200 hot path (SURVEY.md §2 row 3)")
    if alg not in ("bps", "bpsGPU"):
        logg.error("CPR algorithm incorrectly specified.")
        raise NameError("name 'phaseEst' is not defined")  # the reference falls through to :154

    # FOE + pnorm (:124-131), bps (:138), unwrap(4φ)/4 (:154) and pnorm(x·e^{jφ}) (:162) in one device call
    torch = _cabi.require_cuda()
    from . import _engine
    x = _engine.as_host_complex(sigIn)
    L, nModes = x.shape
    d_x = torch.from_numpy(x.view(np.float32 if x.dtype == np.complex64 else np.float64)).to("cuda")
    foeM = M if constType in ["psk", "apsk"] else 4
    if runFOE:
        logg.info("Running frequency offset compensation...")
    logg.info("Running BPS carrier phase recovery...")
    want_fo = runFOE and logg.getLogger().isEnabledFor(logg.INFO)
    d_y, d_ph, fo, _keep = cpr_bps_device(d_x, _engine.dtype_tag(x.dtype), L, nModes, constSymb, B, N, runFOE, 1 / Ts, foeM,
                                          want_fo=want_fo)
    if want_fo:
        logg.info(f"Estimated frequency offset (MHz): {np.round(np.array(fo[:]) / 1e6, 3)}")
    sigOut = d_y.cpu().numpy().view(np.complex128).reshape(L, nModes)
    phaseEst = d_ph.cpu().numpy()

    discard = phaseEst.shape[0] // 4
    if discard > 0 and logg.getLogger().isEnabledFor(logg.INFO):
        sigmaPhase = np.mean(np.var(np.diff(phaseEst[discard:-discard, :], axis=0), axis=0))
        logg.info(f"Estimated linewidth: {sigmaPhase / (2 * np.pi * Ts) / 1e3:.3f} kHz")

    if input1D:
        sigOut = sigOut.flatten()
        phaseEst = phaseEst.flatten()
    return (sigOut, phaseEst) if returnPhases else sigOut
