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
