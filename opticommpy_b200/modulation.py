"""Host-side constellation construction used by the equalizer and the carrier recovery.

The ORDER of the returned points defines the symbol/decision indices, so it reproduces
``optic.comm.modulation.grayMapping`` (optic/comm/modulation.py:64-118; ``qamConst`` :143-174,
``pskConst`` :177-197, ``apskConst`` :200-269): point ``i`` of the raster constellation is stored
at position ``gray(i) = i ^ (i >> 1)`` and the result is complex64.
"""
from __future__ import annotations

import numpy as np


def _qam_raster(M: int) -> np.ndarray:
    side = int(np.sqrt(M))
    if side * side != M:
        raise ValueError("square QAM needs M to be a perfect square")
    levels = np.arange(-(side - 1), side, 2).astype(np.float64)
    grid = np.empty((side, side), dtype=np.complex128)
    for r in range(side):
        re = levels[::-1] if (r % 2 == 1) else levels  # boustrophedon rows keep neighbours adjacent
        grid[r, :] = re + 1j * levels[side - 1 - r]
    return grid.reshape(-1)


def _psk_raster(M: int) -> np.ndarray:
    return np.exp(1j * (np.arange(M) * (2 * np.pi / M)))


_APSK_RING_BITS = {16: 1, 32: 2, 64: 2, 128: 3, 256: 3, 512: 4, 1024: 4}


def _apsk_raster(M: int) -> np.ndarray:
    m1 = _APSK_RING_BITS[M]
    rings = 1 << m1
    per_ring = 1 << int(np.log2(M) - m1)
    pts = np.zeros(M, dtype=np.complex64)
    for k in range(rings):
        radius = np.sqrt(-np.log(1 - (k + 0.5) * per_ring / M))
        ring = _psk_raster(per_ring)
        if k % 2 == 0:
            ring = ring[::-1]
        pts[k * per_ring:(k + 1) * per_ring] = radius * ring
    return pts * np.exp(1j * (np.pi / per_ring))


def grayMapping(M, constType):
    """Constellation of order ``M`` sorted by the integer value of each point's Gray label."""
    if constType == "ook":
        M = 2
        raster = np.arange(0, 2).astype(np.float64)
    elif constType == "pam":
        raster = np.arange(-(M - 1), M, 2).astype(np.float64)
    elif constType == "qam":
        raster = _qam_raster(M)
    elif constType == "psk":
        raster = _psk_raster(M)
    elif constType == "apsk":
        raster = _apsk_raster(M)
    else:
        raise ValueError(f"unknown constellation type {constType!r}")
    dtype = np.float32 if constType in ("pam", "ook") else np.complex64
    out = np.zeros(M, dtype=dtype)
    idx = np.arange(M)
    out[idx ^ (idx >> 1)] = raster.astype(dtype)
    return out


def normalizedConstellation(M, constType, shapingFactor=0.0, prec=None):
    """Unit-power constellation under a Maxwell-Boltzmann pmf (equalization.py:234-241,
    carrierRecovery.py:118-121)."""
    c = grayMapping(M, constType)
    if prec is not None:
        c = c.astype(prec)
    px = np.exp(-shapingFactor * np.abs(c) ** 2)
    px = px / np.sum(px)
    c /= np.sqrt(np.sum(np.abs(c) ** 2 * px))
    return c


def _device_decisions(symb, const, want_idx, want_bits):
    """Run ``ocb_min_euclid`` on ``symb`` (any real/complex array) against ``const``; int64 outputs."""
    import ctypes as C

    from . import _cabi, _engine
    torch = _cabi.require_cuda()
    lib = _cabi.lib()
    x = _engine.as_host_complex(np.asarray(symb).reshape(-1))
    c = np.ascontiguousarray(np.asarray(const).reshape(-1).astype(np.complex128))
    n, M = x.shape[0], c.shape[0]
    nbits = int(np.log2(M))
    d_x = torch.from_numpy(x.view(np.float32 if x.dtype == np.complex64 else np.float64)).to("cuda")
    d_c = torch.from_numpy(c.view(np.float64)).to("cuda")
    d_idx = torch.empty(n, dtype=torch.int64, device="cuda") if want_idx else None
    d_bits = torch.empty(n * nbits, dtype=torch.int64, device="cuda") if want_bits else None
    vp = C.c_void_p
    _cabi.check(
        lib.ocb_min_euclid(vp(d_x.data_ptr()), _engine.dtype_tag(x.dtype), n, vp(d_c.data_ptr()), M,
                           vp(d_idx.data_ptr() if want_idx else None), vp(d_bits.data_ptr() if want_bits else None),
                           vp(_cabi.stream_ptr(torch))),
        "ocb_min_euclid",
    )
    return (d_idx.cpu().numpy() if want_idx else None), (d_bits.cpu().numpy() if want_bits else None)


def minEuclid(symb, const):
    """Index of the closest constellation symbol for every entry of ``symb`` (1-D), on the GPU.

    Mirror of ``optic.comm.modulation.minEuclid`` (modulation.py:271-299): int64 indices, first index on
    exact ties.  Distances are evaluated in float64 for every input dtype.
    """
    return _device_decisions(symb, const, True, False)[0]


def demodulateGray(symb, M, constType):
    """Hard-decision demodulation to bits (modulation.py:369-408) on the GPU: log2(M) bits per symbol, most
    significant first, as an int64 array like the reference's ``dtype="int"``."""
    if M != 2 and constType == "ook":
        M = 2
    return _device_decisions(symb, grayMapping(M, constType), False, True)[1]
