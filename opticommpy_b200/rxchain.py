"""Device-resident receiver chain: ``edc -> mimoAdaptEqualizer -> cpr`` without crossing PCIe between the stages.

The reference runs the three calls one after the other on host arrays
(examples/test_WDM_transmission.ipynb:1140-1146: ``edc`` -> ``mimoAdaptEqualizer`` -> ``cpr``).  ``rxChain`` takes the
same three ``parameters`` objects, uploads the received samples ONCE, keeps every intermediate on the GPU (the same
kernels the stand-alone mirrors call: ``ocb_edc_run``, ``ocb_mimo_eq_run`` / ``ocb_mimo_eq_rls_run``,
``ocb_cpr_bps_run``) and downloads the recovered symbols once.  Stage by stage the results are those of the
stand-alone calls fed with each other's outputs (tests/test_gpu_rxchain.py checks bit-identity), except that the
intermediate arrays are never rounded through the caller's dtype: the chain hands complex64 from stage to stage, which
is what the stand-alone calls produce for complex64 input as well.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _cabi, _engine
from .carrierRecovery import cpr_bps_device
from .channels import _require_Fs
from .equalization import (_edc_taps, _parse_equalizer_args, _ptr, _to_device, edc_rows_device,
                           equalizer_stages_device)
from .modulation import grayMapping

_vp = C.c_void_p


def rxChain(sigIn, paramEDC, paramEq, paramCPR, symbRef=None, returnAll=False, timing=None):
    """
    Chromatic-dispersion compensation, adaptive MIMO equalization and BPS carrier recovery in one device-resident pass.

    Parameters
    ----------
    sigIn : (N, nModes) complex array at ``paramEq.SpS`` samples per symbol.
    paramEDC, paramEq, paramCPR : the ``parameters`` objects of ``edc``, ``mimoAdaptEqualizer`` and ``cpr``.
    symbRef : reference symbols for data-aided equalizer stages (optional).
    returnAll : also return the equalizer taps, the squared error and the phase estimate.
    timing : optional dict that receives the per-stage device times [ms] (CUDA events on the launching stream).

    Returns
    -------
    sigOut : (totalNumSymb, nModes) complex128 recovered symbols (the output of ``cpr``);
    with ``returnAll``: ``(sigOut, H, errSq, phaseEst)``.
    """
    torch = _cabi.require_cuda()
    lib = _cabi.lib()
    st = _vp(_cabi.stream_ptr(torch))
    Fs = _require_Fs(paramEDC)
    sigIn = np.asarray(sigIn)
    if sigIn.ndim != 2:
        raise IndexError("rxChain expects a 2-D (samples, modes) array")
    h, K, _ = _edc_taps(paramEDC, Fs)
    s = _parse_equalizer_args(sigIn, paramEq, symbRef)   # geometry, stages, constellation, initial taps
    nM, Nsig = s.nModes, s.sig.shape[0]

    ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)] if timing is not None else None
    mark = (lambda i: ev[i].record()) if ev else (lambda i: None)

    # ---- upload once (raw dtype), planar rows for the EDC --------------------------------------------------------
    mark(0)
    host = s.sig
    # one host -> device copy of the raw samples (pageable memory is fine for a single transfer; allocating a pinned
    # staging buffer per call costs far more than it saves)
    d_raw = torch.from_numpy(host.view(np.float32 if host.dtype == np.complex64 else np.float64).reshape(host.shape + (2,))).to("cuda")
    d_rows = torch.empty((nM, Nsig, 2), dtype=torch.float32, device="cuda")
    _cabi.check(lib.ocb_pack_fields(_ptr(d_raw), _engine.dtype_tag(host.dtype), Nsig, nM, 0, _ptr(d_rows), st), "ocb_pack_fields")
    mark(1)
    d_edc = torch.empty_like(d_rows)
    keep = [edc_rows_device(d_rows, d_edc, h)]
    # ---- equalizer input: interleaved (sample, mode), zero-padded by floor(nTaps/2) rows at both ends -----------
    d_x = torch.zeros((1, s.nPad, nM, 2), dtype=torch.float32, device="cuda")
    _cabi.check(lib.ocb_unpack_fields(_ptr(d_edc), Nsig, nM, 0, _ptr(d_x, s.Lpad * nM * 8), _cabi.OCB_C64, st),
                "ocb_unpack_fields")
    mark(2)
    d_ref, Lref = None, 0
    if s.symbRef is not None:
        ref = np.ascontiguousarray(s.symbRef.astype(np.complex64))
        d_ref, Lref = _to_device(torch, ref.view(np.float32)).reshape(1, ref.shape[0], nM, 2), ref.shape[0]
    d_H = _to_device(torch, s.H[None].view(np.float32))
    d_Hw = _to_device(torch, s.H_[None].view(np.float32)) if s.runWL else None
    d_y, d_e, _ = equalizer_stages_device(s, 1, d_x, d_ref, Lref, d_H, d_Hw)
    mark(3)
    # ---- carrier recovery on the equalizer's complex64 output ------------------------------------------------------
    alg = getattr(paramCPR, "alg", "bps")
    if alg not in ("bps", "bpsGPU"):
        raise NotImplementedError("rxChain runs cpr with alg='bps' only")
    M = getattr(paramCPR, "M", 4)
    constType = getattr(paramCPR, "constType", "qam")
    shapingFactor = getattr(paramCPR, "shapingFactor", 0)
    c = grayMapping(M, constType)
    px = np.exp(-shapingFactor * np.abs(c) ** 2)
    px = px / np.sum(px)
    c /= np.sqrt(np.sum(np.abs(c) ** 2 * px))
    Ts = getattr(paramCPR, "Ts", 1 / 32e9)
    foeM = M if constType in ["psk", "apsk"] else 4
    L = s.totalNumSymb
    d_out, d_ph, _, keep2 = cpr_bps_device(d_y, _cabi.OCB_C64, L, nM, c, getattr(paramCPR, "B", 64), getattr(paramCPR, "N", 35),
                                           getattr(paramCPR, "runFOE", True), 1 / Ts, foeM)
    mark(4)
    sigOut = d_out.cpu().numpy().view(np.complex128).reshape(L, nM)   # the one device -> host copy (synchronises)
    if timing is not None:
        names = ["h2d_pack", "edc", "equalizer", "cpr_bps"]
        for i, n in enumerate(names):
            timing[n] = ev[i].elapsed_time(ev[i + 1])
        timing["h2d_bytes"] = int(host.nbytes)
        timing["d2h_bytes"] = int(sigOut.nbytes)
    if not returnAll:
        return sigOut
    H = d_H.cpu().numpy().view(np.complex64).reshape(nM * nM, s.nTaps)
    errSq = d_e.cpu().numpy()[0]
    return sigOut, H, errSq, d_ph.cpu().numpy()
