"""Drop-in mirror of ``optic.comm.metrics.fastBERcalc`` (optic/comm/metrics.py:110-195) on the GPU.

One call per waveform: rotation estimate, power normalisation, SNR estimate, hard decisions of both
sequences and the bit / symbol error counts all happen on the device (``ocb_ber_count``); three
scalars per mode come back.  ``fastBERcalc`` also accepts CUDA ``torch`` tensors (complex64/complex128,
shape (L, nModes)) so that a Monte-Carlo sweep can keep its waveforms on the device and gather only the
scalars.
"""
from __future__ import annotations

import ctypes as C
import logging as logg

import numpy as np

from . import _cabi
from .modulation import grayMapping

_vp = C.c_void_p


def _columns_host(x):
    """(L, nModes) C-contiguous complex copy with the reference's orientation rule (metrics.py:158-167)."""
    from . import _engine
    x = np.asarray(x)
    if x.ndim == 1:
        x = x.reshape(-1, 1)
    elif x.shape[1] > x.shape[0]:
        x = x.T
    return _engine.as_host_complex(x)


def _columns_device(torch, x):
    if x.dim() == 1:
        x = x.reshape(-1, 1)
    elif x.shape[1] > x.shape[0]:
        x = x.T
    if x.dtype not in (torch.complex64, torch.complex128):
        x = x.to(torch.complex128)
    return x.contiguous()


def fastBERcalc(rx, tx, M, constType, px=None, returnCounts=False):
    """
    Monte Carlo BER/SER/SNR calculation on the GPU.

    Parameters as in the reference: ``rx``/``tx`` symbol sequences (1-D or (L, nModes); a wide array is
    transposed), modulation order ``M``, ``constType`` in 'qam', 'psk', 'apsk', 'pam', 'ook', optional prior
    ``px``.  Returns ``(BER, SER, SNR)`` float64 arrays of length nModes; with ``returnCounts`` (extension)
    also the int64 ``(2, nModes)`` bit / symbol error counts.
    """
    if M != 2 and constType == "ook":
        logg.warning("OOK has only 2 symbols, but M != 2. Changing M to 2.")
        M = 2
    if px is None:
        px = []
    if len(px) == 0:
        px = 1 / M * np.ones(M)
    constSymb = grayMapping(M, constType)
    Es = np.sum(np.abs(constSymb) ** 2 * px)  # metrics.py:151

    torch = _cabi.require_cuda()
    lib = _cabi.lib()
    from . import _engine
    if hasattr(rx, "data_ptr"):  # CUDA tensors stay where they are
        d_rx, d_tx = _columns_device(torch, rx), _columns_device(torch, tx)
        if d_rx.dtype != d_tx.dtype:
            d_rx, d_tx = d_rx.to(torch.complex128), d_tx.to(torch.complex128)
        tag = _cabi.OCB_C64 if d_rx.dtype == torch.complex64 else _cabi.OCB_C128
        shape_rx, shape_tx = tuple(d_rx.shape), tuple(d_tx.shape)
    else:
        hrx, htx = _columns_host(rx), _columns_host(tx)
        if hrx.dtype != htx.dtype:
            hrx, htx = hrx.astype(np.complex128), htx.astype(np.complex128)
        fview = np.float32 if hrx.dtype == np.complex64 else np.float64
        d_rx = torch.from_numpy(hrx.view(fview)).to("cuda")
        d_tx = torch.from_numpy(htx.view(fview)).to("cuda")
        tag = _engine.dtype_tag(hrx.dtype)
        shape_rx, shape_tx = hrx.shape, htx.shape
    if shape_rx != shape_tx:
        raise ValueError("rx and tx must have the same shape")
    L, nModes = int(shape_tx[0]), int(shape_tx[1])
    c128 = np.ascontiguousarray(constSymb.astype(np.complex128))
    d_c = torch.from_numpy(c128.view(np.float64)).to("cuda")
    ws_bytes = int(lib.ocb_ber_workspace_bytes(nModes))
    d_ws = torch.empty(ws_bytes // 8 + 1, dtype=torch.float64, device="cuda")
    ber, ser, snr = (C.c_double * nModes)(), (C.c_double * nModes)(), (C.c_double * nModes)()
    counts = (C.c_int64 * (2 * nModes))()
    _cabi.check(
        lib.ocb_ber_count(_vp(d_rx.data_ptr()), _vp(d_tx.data_ptr()), tag, L, nModes, _vp(d_c.data_ptr()), len(c128),
                          int(constType in ("qam", "psk")), float(np.sqrt(Es)), ber, ser, snr, counts,
                          _vp(d_ws.data_ptr()), ws_bytes, _vp(_cabi.stream_ptr(torch))),
        "ocb_ber_count",
    )
    out = (np.array(ber[:]), np.array(ser[:]), np.array(snr[:]))
    if returnCounts:
        return out + (np.array(counts[:], dtype=np.int64).reshape(2, nModes),)
    return out
