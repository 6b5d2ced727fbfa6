"""Drop-in mirrors of the Rx front-end glue in ``optic.dsp.core``: ``firFilter`` (optic/dsp/core.py:87-125),
``decimate`` (:435-491) and ``pnorm`` (:702-717) — SURVEY.md §8f rank 3, the calls between the fiber model and ``edc`` in the reference
notebooks.  numpy in, numpy out; the arithmetic runs on the GPU through the C-ABI (``ocb_edc_run`` — an
overlap-save linear convolution — and ``ocb_decimate_run``).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _cabi, _engine

_vp = C.c_void_p


def _ptr(t):
    return _vp(t.data_ptr())


def _as_columns(x):
    x = np.asarray(x)
    if x.ndim == 1:
        return x.reshape(len(x), 1), True
    return x, False


def _upload_rows(torch, lib, st, cols):
    """(N, nModes) host array -> planar complex64 rows[nModes][N] on the device (+ the complex host view used)."""
    host = _engine.as_host_complex(cols)
    d_raw = torch.from_numpy(host.view(np.float32 if host.dtype == np.complex64 else np.float64)).to("cuda")
    N, nModes = cols.shape
    d_x = torch.empty((nModes, N, 2), dtype=torch.float32, device="cuda")
    _cabi.check(lib.ocb_pack_fields(_ptr(d_raw), _engine.dtype_tag(host.dtype), N, nModes, 0, _ptr(d_x), st), "ocb_pack_fields")
    return d_x, host.dtype


def _download_rows(torch, lib, st, d_rows, N, nModes, host_dtype):
    d_out = torch.empty((N, nModes, 2), dtype=torch.float32 if host_dtype == np.complex64 else torch.float64, device="cuda")
    _cabi.check(lib.ocb_unpack_fields(_ptr(d_rows), N, nModes, 0, _ptr(d_out), _engine.dtype_tag(host_dtype), st),
                "ocb_unpack_fields")
    return d_out.cpu().numpy().view(host_dtype).reshape(N, nModes)


def firFilter(h, x):
    """
    FIR filtering with delay compensation: ``y[:, n] = scipy.signal.fftconvolve(x[:, n], h, mode="same")``
    for every column (core.py:115-119).  ``h``: filter taps (real or complex), ``x``: (N,) or (N, nModes).
    The output has the shape and dtype of ``x`` (the reference writes into ``x.copy()``).
    """
    x = np.asarray(x)
    cols, input1D = _as_columns(x)
    h = np.asarray(h).reshape(-1)
    N, nModes = cols.shape
    K = int(h.size)
    if K < 1 or N < 1:
        raise ValueError("firFilter needs a non-empty filter and signal")
    if K > N:
        raise ValueError("firFilter: filters longer than the signal are not supported on the GPU path")
    torch = _cabi.require_cuda()
    lib = _cabi.lib()
    st = _vp(_cabi.stream_ptr(torch))
    d_x, host_dtype = _upload_rows(torch, lib, st, cols)
    d_y = torch.empty_like(d_x)
    d_h = torch.from_numpy(np.ascontiguousarray(h.astype(np.complex64)).view(np.float32)).to("cuda")
    ws_bytes = int(lib.ocb_edc_workspace_bytes(N, nModes, K))
    d_ws = torch.empty(ws_bytes + 256, dtype=torch.uint8, device="cuda")
    ws_ptr = (d_ws.data_ptr() + 255) // 256 * 256
    _cabi.check(lib.ocb_edc_run(_ptr(d_x), _ptr(d_y), N, nModes, _ptr(d_h), K, _vp(ws_ptr), ws_bytes, st), "ocb_edc_run")
    out = _download_rows(torch, lib, st, d_y, N, nModes, host_dtype)
    if np.iscomplexobj(x):
        y = out.astype(x.dtype, copy=False)
    else:
        y = out.real.astype(x.dtype)  # the reference assigns into a copy of x, which keeps x's dtype
    return y.flatten() if input1D else y


def decimate(sigIn, param):
    """
    Decimate a signal at its maximum-variance sampling instant (core.py:435-491).

    ``param.SpSin`` / ``param.SpSout``: samples per symbol of the input / output.  Per column the phase
    ``p`` in ``[0, SpSin)`` with the largest ``var(sigIn[p::SpSin])`` is found (first one on ties), the column
    is rolled by ``-p`` and every ``int(SpSin / SpSout)``-th sample is kept.
    """
    sigIn = np.asarray(sigIn)
    cols, input1D = _as_columns(sigIn)
    N, nModes = cols.shape
    SpSin = int(param.SpSin)
    decFactor = int(param.SpSin / param.SpSout)
    if N % SpSin != 0:
        raise ValueError(f"cannot reshape array of size {N} into shape ({SpSin})")  # numpy's reshape(-1, SpSin) error
    torch = _cabi.require_cuda()
    lib = _cabi.lib()
    st = _vp(_cabi.stream_ptr(torch))
    d_x, host_dtype = _upload_rows(torch, lib, st, cols)
    Nout = (N + decFactor - 1) // decFactor
    d_y = torch.empty((nModes, Nout, 2), dtype=torch.float32, device="cuda")
    d_delay = torch.empty(nModes, dtype=torch.int32, device="cuda")
    ws_bytes = int(lib.ocb_decimate_workspace_bytes(nModes, SpSin))
    d_ws = torch.empty(ws_bytes + 256, dtype=torch.uint8, device="cuda")
    ws_ptr = (d_ws.data_ptr() + 255) // 256 * 256
    _cabi.check(lib.ocb_decimate_run(_ptr(d_x), _ptr(d_y), N, nModes, SpSin, decFactor, _ptr(d_delay), _vp(ws_ptr),
                                     ws_bytes, st), "ocb_decimate_run")
    out = _download_rows(torch, lib, st, d_y, Nout, nModes, host_dtype)
    if np.iscomplexobj(sigIn):
        sigOut = out.astype(sigIn.dtype, copy=False)
    else:
        sigOut = out.real.astype(sigIn.dtype)
    return sigOut.flatten() if input1D else sigOut


def pnorm(x):
    """
    Normalise the average power: ``x / sqrt(mean(|x|^2))`` with the mean over the WHOLE array (core.py:717).
    Complex (or real) array of any shape in, complex128 (float64 for real input) array of the same shape out.
    """
    x = np.asarray(x)
    host = np.ascontiguousarray(x.astype(np.complex128))
    if host.size == 0:
        return host if np.iscomplexobj(x) else host.real
    torch = _cabi.require_cuda()
    lib = _cabi.lib()
    st = _vp(_cabi.stream_ptr(torch))
    d = torch.from_numpy(host.view(np.float64)).to("cuda")
    d_ws = torch.empty(4096 * 8 + 256, dtype=torch.uint8, device="cuda")
    ws_ptr = (d_ws.data_ptr() + 255) // 256 * 256
    _cabi.check(lib.ocb_pnorm_run(_ptr(d), host.size, _vp(ws_ptr), 4096 * 8, st), "ocb_pnorm_run")
    out = d.cpu().numpy().view(np.complex128).reshape(x.shape)
    return out if np.iscomplexobj(x) else out.real


def delay_rows_device(d_rows, delay, Fs, NFFT=1024):
    """Device-resident ``delaySignal`` (optic/dsp/core.py:880-922; default NFFT = 1024, a 512-tap filter) on planar complex64 rows (nRows, N, 2) float32 CUDA:
    the reference pads the signal by ceil(|delay Fs|) zeros, filters it with the time-domain taps of
    ``H = exp(-j 2 pi f delay)`` sampled on ``fftfreq(NFFT // 2)`` through blockwiseFFTConv (= a linear convolution,
    delay-compensated by (K - 1) // 2), rolls the result by -1 and keeps the first N samples.  Here the convolution is
    ``ocb_edc_run`` with the same taps."""
    torch = _cabi.require_cuda()
    from .equalization import edc_rows_device
    nRows, N = int(d_rows.shape[0]), int(d_rows.shape[1])
    padLen = int(np.ceil(np.abs(delay * Fs)))
    if NFFT is None:
        NFFT = 2 ** int(np.ceil(np.log2(N + padLen)))
    freq = np.fft.fftfreq(NFFT // 2, d=1 / Fs)
    H = np.exp(-1j * 2 * np.pi * freq * delay)
    h = np.fft.fftshift(np.fft.ifft(H)).astype(np.complex64)          # core.py:1016
    d_pad = torch.zeros((nRows, N + padLen, 2), dtype=torch.float32, device="cuda")
    d_pad[:, :N] = d_rows
    d_y = torch.empty_like(d_pad)
    keep = edc_rows_device(d_pad, d_y, h)
    out = torch.roll(d_y, -1, dims=1)[:, :N].contiguous()             # core.py:920-922
    return out, keep


def delaySignal(sig, delay, Fs=1, NFFT=1024):
    """
    Apply a time delay to a 1-D signal with an FFT-convolved fractional-delay filter (core.py:880-922, NFFT [1024]).
    Complex in, complex out; real in, real out.
    """
    sig = np.asarray(sig)
    cols, _ = _as_columns(sig)
    torch = _cabi.require_cuda()
    lib = _cabi.lib()
    st = _vp(_cabi.stream_ptr(torch))
    d_x, host_dtype = _upload_rows(torch, lib, st, cols)
    d_y, _keep = delay_rows_device(d_x, delay, Fs, NFFT)
    out = _download_rows(torch, lib, st, d_y, cols.shape[0], 1, host_dtype)[:, 0]
    if np.iscomplexobj(sig):
        return out.astype(sig.dtype, copy=False)
    return out.real.astype(sig.dtype)


def symbolSync(rx, tx, SpS, mode="amp"):
    """
    Symbol synchronizer (core.py:552-675): align the transmitted sequence ``tx`` to the received one ``rx`` (``SpS``
    samples per symbol).  'amp': correlation of the mean-removed magnitudes picks, for every received mode, the
    transmitted column and its delay; 'real': correlations of real / imaginary parts also resolve pi/2 rotations and
    conjugation.  Returns ``tx`` permuted, rotated and rolled.  The correlations (cuFFT, float64), the sequences and the
    final gather run on the device; only the nModes x nModes peak scalars come back for the decisions.
    """
    rx = np.asarray(rx)
    tx = np.asarray(tx)
    input1D = rx.ndim == 1
    if input1D:
        rx = rx.reshape(len(rx), 1)
    if tx.ndim == 1:
        tx = tx.reshape(len(tx), 1)
    nModes = rx.shape[1]
    if SpS > 1:
        class _P:
            pass
        pd = _P()
        pd.SpSin, pd.SpSout = SpS, 1
        rx = decimate(rx, pd)
    torch = _cabi.require_cuda()
    lib = _cabi.lib()
    st = _vp(_cabi.stream_ptr(torch))
    tx128 = np.ascontiguousarray(tx.astype(np.complex128))
    rx128 = np.ascontiguousarray(rx.astype(np.complex128))
    Lt, nT = tx128.shape
    Lr = rx128.shape[0]
    d_tx = torch.from_numpy(tx128.view(np.float64)).to("cuda")
    d_rx = torch.from_numpy(rx128.view(np.float64)).to("cuda")

    def seq(d, nCols, L, kind):
        out = torch.empty((nCols, L), dtype=torch.float64, device="cuda")
        _cabi.check(lib.ocb_sync_sequence_run(_ptr(d), nCols, L, kind, _ptr(out), st), "ocb_sync_sequence_run")
        return out

    def peaks(a, b):
        nA, nB = int(a.shape[0]), int(b.shape[0])
        ws_bytes = int(lib.ocb_xcorr_workspace_bytes(nA, Lt, nB, Lr))
        d_ws = torch.empty(ws_bytes + 256, dtype=torch.uint8, device="cuda")
        ws_ptr = (d_ws.data_ptr() + 255) // 256 * 256
        idx = (C.c_int64 * (nA * nB))()
        val = (C.c_double * (nA * nB))()
        _cabi.check(lib.ocb_xcorr_peak_run(_ptr(a), nA, Lt, _ptr(b), nB, Lr, idx, val, _vp(ws_ptr), ws_bytes, st),
                    "ocb_xcorr_peak_run")
        return np.array(idx[:]).reshape(nA, nB), np.array(val[:]).reshape(nA, nB)

    delay = np.zeros(nModes, dtype=np.int64)
    rot = np.ones(nModes, dtype=np.complex128)
    conj = np.zeros(nModes, dtype=np.int32)
    if mode == "amp":
        idx, val = peaks(seq(d_tx, nT, Lt, 0), seq(d_rx, nModes, Lr, 0))
        corr = np.abs(val)                                            # corrMatrix[m, n] (:609)
        swap = np.argmax(corr, axis=0)
        for k in range(nModes):
            delay[k] = idx[swap[k], k] - Lt + 1                       # finddelay (:696-697)
    elif mode == "real":
        t_re, t_im = seq(d_tx, nT, Lt, 1), seq(d_tx, nT, Lt, 2)
        r_re, r_im = seq(d_rx, nModes, Lr, 1), seq(d_rx, nModes, Lr, 2)
        i_rr, v_rr = peaks(t_re, r_re)                                # correlate(Re tx_m, Re rx_n)
        i_ir, v_ir = peaks(t_im, r_re)                                # correlate(Im tx_m, Re rx_n)
        _, v_ri = peaks(t_re, r_im)                                   # correlate(Re tx_m, Im rx_n)
        _, v_ii = peaks(t_im, r_im)                                   # correlate(Im tx_m, Im rx_n)
        corr = np.maximum(np.abs(v_rr), np.abs(v_ir))                 # :633
        rot_mn = np.ones((nT, nModes), dtype=np.complex128)
        for m in range(nT):
            for n in range(nModes):
                if abs(v_rr[m, n]) > abs(v_ir[m, n]):                 # :636-645 (pi/2 rotations)
                    rot_mn[m, n] = 1 if v_rr[m, n] > 0 else -1
                else:
                    rot_mn[m, n] = -1j if v_ir[m, n] > 0 else 1j
        swap = np.argmax(corr, axis=0)
        for k in range(nModes):
            r = rot_mn[k, swap[k]]                                    # :652, indexed as in the reference
            rot[k] = r
            m = swap[k]
            # Re(r z) and Im(r z) of the swapped column are +-Re z or +-Im z: reuse the peaks of those correlations
            if r in (1, -1):
                delay[k] = i_rr[m, k] - Lt + 1                        # finddelay(Re(r tx), Re rx): |.| ignores the sign
                cii_peak = r.real * v_ii[m, k]                        # Im(r z) = r Im z
            else:
                delay[k] = i_ir[m, k] - Lt + 1                        # Re(-j z) = Im z, Re(j z) = -Im z
                cii_peak = (-1.0 if r == -1j else 1.0) * v_ri[m, k]   # Im(-j z) = -Re z, Im(j z) = Re z
            if cii_peak < 0:                                          # :658-662
                conj[k] = 1
    else:
        swap = np.arange(nModes)
    d_out = torch.empty((Lt, nModes, 2), dtype=torch.float64, device="cuda")
    d_swap = torch.from_numpy(np.asarray(swap, dtype=np.int32)).to("cuda")
    d_rot = torch.from_numpy(rot.view(np.float64)).to("cuda")
    d_conj = torch.from_numpy(conj).to("cuda")
    d_delay = torch.from_numpy(delay).to("cuda")
    _cabi.check(lib.ocb_sync_apply_run(_ptr(d_tx), _ptr(d_out), Lt, nT, _ptr(d_swap), _ptr(d_rot), _ptr(d_conj),
                                       _ptr(d_delay), st), "ocb_sync_apply_run")
    out = d_out.cpu().numpy().view(np.complex128).reshape(Lt, nModes)
    if np.iscomplexobj(tx):
        out = out.astype(tx.dtype, copy=False)
    return out.flatten() if input1D else out
