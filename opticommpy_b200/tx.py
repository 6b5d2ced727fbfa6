"""Drop-in mirror of ``optic.models.tx.simpleWDMTx`` (optic/models/tx.py:42-228) — SURVEY.md §8f rank 4, the input
generator of the fiber model — with the per-sample work on the GPU: upsampling, pulse shaping (overlap-save FIR), amplitude
normalisation, IQ modulator, power normalisation, per-channel frequency shift and the sum over the WDM channels
(``ocb_upsample_run``, ``ocb_edc_run``, ``ocb_wdm_tx_combine_run``).  What stays on the host is what the reference draws
from numpy's legacy generator — the symbol indices (optic/comm/sources.py:167-211) and the laser phase-noise walk
(optic/dsp/core.py:791-826) — and the filter taps (``pulseShape``, core.py:211-269), so that a seeded call transmits the
reference's symbols exactly.

``wdm_tx_rows_device(param)`` returns the field as planar complex64 CUDA rows, ready for ``channels.manakov_rows_device``.
"""
from __future__ import annotations

import ctypes as C
import logging as logg

import numpy as np

from . import _cabi
from .modulation import _apsk_raster, _psk_raster, _qam_raster
from .utils import parameters

_vp = C.c_void_p


def _ptr(t):
    return _vp(t.data_ptr())


def symbolSource(param):
    """
    Random symbol sequence of a modulation format (optic/comm/sources.py:112-211): nSymbols [1000], M [4],
    constType ['qam' | 'pam' | 'psk' | 'apsk'], dist ['uniform' | 'maxwell-boltzmann'], shapingFactor [0], px [None], seed [None].
    The constellation is scaled to unit average power under the pmf and sampled with the legacy generator
    (``np.random.seed(seed)`` + ``np.random.choice(..., p=px)``), i.e. the reference's symbols for the same seed.
    """
    nSymbols = getattr(param, "nSymbols", 1000)
    M = getattr(param, "M", 4)
    constType = getattr(param, "constType", "qam")
    dist = getattr(param, "dist", "uniform")
    shapingFactor = getattr(param, "shapingFactor", 0.0)
    px = getattr(param, "px", None)
    seed = getattr(param, "seed", None)
    rs = np.random.RandomState(seed) if seed is not None else np.random
    if constType == "qam":
        const = _qam_raster(M)
    elif constType == "pam":
        const = np.arange(-(M - 1), M, 2)
    elif constType == "psk":
        const = _psk_raster(M)
    elif constType == "apsk":
        const = _apsk_raster(M)
    else:
        raise ValueError("Invalid constellation type. Supported types are 'qam', 'pam', 'psk', and 'apsk'.")
    if px is None:
        if dist == "uniform":
            px = np.full(M, 1.0 / M)
        elif dist == "maxwell-boltzmann":
            px = np.exp(-shapingFactor * np.abs(const) ** 2)
            px = (px / np.sum(px)).flatten()
        else:
            raise ValueError("Invalid probability distribution.")
    const = const / np.sqrt(np.sum(px * np.abs(const) ** 2))
    return rs.choice(const, nSymbols, p=px)


def _rrc(t, alpha):
    """Root-raised-cosine taps for unit symbol period (core.py:128-173), vectorised over the three branches."""
    t = np.asarray(t, dtype=np.float64)
    out = np.empty_like(t)
    centre = t == 0
    edge = np.abs(t) == 1 / (4 * alpha) if alpha > 0 else np.zeros_like(centre)
    rest = ~(centre | edge)
    out[centre] = 1 + alpha * (4 / np.pi - 1)
    out[edge] = (alpha / np.sqrt(2)) * ((1 + 2 / np.pi) * np.sin(np.pi / (4 * alpha)) + (1 - 2 / np.pi) * np.cos(np.pi / (4 * alpha))) if alpha > 0 else 0
    tr = t[rest]
    out[rest] = (np.sin(np.pi * tr * (1 - alpha)) + 4 * alpha * tr * np.cos(np.pi * tr * (1 + alpha))) / (np.pi * tr * (1 - (4 * alpha * tr) ** 2))
    return out


def _rc(t, alpha):
    """Raised-cosine taps for unit symbol period (core.py:176-208)."""
    t = np.asarray(t, dtype=np.float64)
    out = np.empty_like(t)
    edge = np.abs(t) == 1 / (2 * alpha) if alpha > 0 else np.zeros(t.shape, dtype=bool)
    out[edge] = np.pi / 4 * np.sinc(1 / (2 * alpha)) if alpha > 0 else 0
    tr = t[~edge]
    out[~edge] = np.sinc(tr) * np.cos(np.pi * alpha * tr) / (1 - 4 * alpha ** 2 * tr ** 2)
    return out


def pulseShape(param):
    """
    Pulse-shaping filter taps, normalised to unit sum (optic/dsp/core.py:211-269): pulseType ['rrc' | 'rc' | 'rect' | 'nrz'],
    SpS [2], nFilterTaps [256], rollOff [0.1].  Host-side constants of the transmitter (float64).
    """
    pulseType = getattr(param, "pulseType", "rrc")
    SpS = getattr(param, "SpS", 2)
    nFilterTaps = getattr(param, "nFilterTaps", 256)
    rollOff = getattr(param, "rollOff", 0.1)
    if pulseType == "rect":
        pulse = np.concatenate((np.zeros(int(SpS / 2)), np.ones(SpS), np.zeros(int(SpS / 2))))
    elif pulseType == "nrz":
        t = np.linspace(-2, 2, SpS)
        pulse = np.convolve(np.ones(SpS), 2 / np.sqrt(np.pi) * np.exp(-(t ** 2)), mode="full")
    elif pulseType == "rrc":
        pulse = _rrc(np.linspace(-nFilterTaps // 2, nFilterTaps // 2, nFilterTaps) * (1 / SpS), rollOff)
    elif pulseType == "rc":
        pulse = _rc(np.linspace(-nFilterTaps // 2, nFilterTaps // 2, nFilterTaps) * (1 / SpS), rollOff)
    else:
        raise ValueError(f"pulseShape: pulse type {pulseType!r} is not supported on this path")
    return pulse / np.sum(pulse)


def phaseNoise(lw, Nsamples, Ts, seed=None):
    """Random-walk laser phase noise (optic/dsp/core.py:791-826): phi[0] = 0, steps N(0, 2 pi lw Ts) from the legacy
    generator (the reference's scalar draws and one vector draw consume the stream identically)."""
    rs = np.random.RandomState(seed) if seed is not None else np.random
    steps = rs.normal(0, np.sqrt(2 * np.pi * lw * Ts), max(int(Nsamples) - 1, 0))
    return np.concatenate((np.zeros(1), np.cumsum(steps)))[:Nsamples]


def _defaults(param):
    for name, val in (("M", 16), ("constType", "qam"), ("Rs", 32e9), ("SpS", 16), ("probDist", "uniform"), ("shapingFactor", 0),
                      ("seed", None), ("nBits", 60000), ("pulseType", "rrc"), ("nFilterTaps", 1024), ("pulseRollOff", 0.01),
                      ("mzmScale", 0.5), ("powerPerChannel", -3), ("nChannels", 5), ("Fc", 193.1e12), ("laserLinewidth", 0),
                      ("wdmGridSpacing", 50e9), ("nPolModes", 1), ("prgsBar", True)):
        setattr(param, name, getattr(param, name, val))  # tx.py:86-104 (defaults are written back)


def wdm_tx_rows_device(param):
    """The transmitter with its output left on the GPU: returns ``(rows, symbTxWDM, param)`` with ``rows`` a planar
    complex64 CUDA tensor (nPolModes, N), N = nSymbols * SpS, and ``symbTxWDM`` the (nSymbols, nPolModes, nChannels)
    complex128 host array of transmitted symbols.  Parameters and defaults as ``simpleWDMTx``."""
    torch = _cabi.require_cuda()
    lib = _cabi.lib()
    st = _vp(_cabi.stream_ptr(torch))
    _defaults(param)
    nCh, nPol, SpS = int(param.nChannels), int(param.nPolModes), int(param.SpS)
    Fs = 1 / ((1 / param.Rs) / SpS)
    bits_per_symbol = int(np.log2(param.M))
    nSym = int(param.nBits / np.log2(param.M))
    if param.probDist not in ("uniform", "maxwell-boltzmann"):
        raise ValueError("Invalid probability distribution.")
    # pmf of the constellation in Gray order (tx.py:111-119): exposed like the reference does
    from .modulation import grayMapping
    c_gray = grayMapping(param.M, param.constType)
    if param.probDist == "uniform":
        param.pmf = np.ones(param.M) / param.M
    else:
        px = np.exp(-param.shapingFactor * np.abs(c_gray) ** 2)
        param.pmf = px / np.sum(px)

    q = parameters()
    q.pulseType, q.nFilterTaps, q.rollOff, q.SpS = param.pulseType, param.nFilterTaps, param.pulseRollOff, SpS
    pulse = pulseShape(q)
    grid = np.arange(-np.floor(nCh / 2), np.floor(nCh / 2) + 1, 1) * param.wdmGridSpacing
    if nCh % 2 == 0:
        grid = grid + param.wdmGridSpacing / 2  # tx.py:141-147 (one entry more than channels for even counts, like the reference)
    if type(param.powerPerChannel) == list:
        assert len(param.powerPerChannel) == nCh, "list length of power per channel does not match number of channels."
        Pch = 10 ** (np.array(param.powerPerChannel, dtype=np.float64) / 10) * 1e-3
    else:
        Pch = 10 ** (param.powerPerChannel / 10) * 1e-3 * np.ones(nCh)

    N = nSym * SpS
    K = int(pulse.size)
    if K > N:
        raise ValueError("simpleWDMTx: the pulse-shaping filter is longer than the signal")
    symb = np.zeros((nSym, nPol, nCh), dtype=np.complex128)
    src = parameters()
    src.nSymbols, src.M, src.constType = param.nBits // bits_per_symbol, param.M, param.constType
    src.dist, src.shapingFactor = param.probDist, param.shapingFactor
    seed = param.seed
    lo_host = None
    for ch in range(nCh):
        logg.info("channel %d\t fc : %3.4f THz" % (ch, (param.Fc + grid[ch]) / 1e12))
        for m in range(nPol):
            src.seed = seed
            symb[:, m, ch] = symbolSource(src)
            if param.seed is not None:
                seed += 1  # tx.py:186-187: a new seed for every pol / channel
            if m == 0 and param.laserLinewidth:  # tx.py:199-203: one laser per channel, drawn with param.seed each time
                if lo_host is None:
                    lo_host = np.empty((nCh, N), dtype=np.complex64)
                lo_host[ch] = np.exp(1j * phaseNoise(param.laserLinewidth, N, 1 / Fs, seed=param.seed))

    R = nCh * nPol
    sym_rows = np.ascontiguousarray(symb.transpose(2, 1, 0).reshape(R, nSym).astype(np.complex64))  # row = ch * nPol + mode
    d_sym = torch.from_numpy(sym_rows.view(np.float32)).to("cuda")
    d_up = torch.empty((R, N, 2), dtype=torch.float32, device="cuda")
    _cabi.check(lib.ocb_upsample_run(_ptr(d_sym), R, nSym, SpS, _ptr(d_up), st), "ocb_upsample_run")
    d_sh = torch.empty_like(d_up)
    d_h = torch.from_numpy(np.ascontiguousarray(pulse.astype(np.complex64)).view(np.float32)).to("cuda")
    ws_bytes = int(lib.ocb_edc_workspace_bytes(N, R, K))
    d_ws = torch.empty(ws_bytes + 256, dtype=torch.uint8, device="cuda")
    ws_ptr = (d_ws.data_ptr() + 255) // 256 * 256
    _cabi.check(lib.ocb_edc_run(_ptr(d_up), _ptr(d_sh), N, R, _ptr(d_h), K, _vp(ws_ptr), ws_bytes, st), "ocb_edc_run")  # firFilter
    d_lo = torch.from_numpy(lo_host.view(np.float32)).to("cuda") if lo_host is not None else None
    prm = _cabi.WdmTxParams(mzmScale=float(param.mzmScale), Vpi=2.0, VbI=-2.0, VbQ=-2.0, Vphi=1.0, ERI=60.0, ERQ=60.0)  # iqm defaults
    rows = torch.empty((nPol, N), dtype=torch.complex64, device="cuda")
    ws2 = int(lib.ocb_wdm_tx_workspace_bytes(nCh, nPol))
    d_ws2 = torch.empty(ws2 + 256, dtype=torch.uint8, device="cuda")
    ws2_ptr = (d_ws2.data_ptr() + 255) // 256 * 256
    pw = (C.c_double * nCh)(*[float(v) for v in Pch])
    fr = (C.c_double * nCh)(*[float(v) for v in grid[:nCh]])
    _cabi.check(lib.ocb_wdm_tx_combine_run(_ptr(d_sh), _ptr(d_lo) if d_lo is not None else None, nCh, nPol, N, C.byref(prm), pw, fr,
                                           float(Fs), _ptr(rows), _vp(ws2_ptr), ws2, st), "ocb_wdm_tx_combine_run")
    param.wdmFreqGrid = grid
    return rows, symb, param


def simpleWDMTx(param):
    """
    Simple WDM transmitter (optic/models/tx.py:42-228) on the GPU.

    param: M [16], constType ['qam'], Rs [32e9], SpS [16], probDist ['uniform'], shapingFactor [0], seed [None],
    nBits [60000], pulseType ['rrc'], nFilterTaps [1024], pulseRollOff [0.01], mzmScale [0.5], powerPerChannel [-3 dBm,
    scalar or list], nChannels [5], Fc [193.1e12], laserLinewidth [0], wdmGridSpacing [50e9], nPolModes [1], prgsBar [True].

    Returns ``(sigTxWDM, symbTxWDM, param)``: the (N, nPolModes) complex WDM field (complex128 array holding complex64
    precision), the (nSymbols, nPolModes, nChannels) symbols per carrier and ``param`` with the defaults,
    ``pmf`` and ``wdmFreqGrid`` written back.
    """
    rows, symb, param = wdm_tx_rows_device(param)
    sig = rows.cpu().numpy().T.astype(np.complex128)
    return np.ascontiguousarray(sig), symb, param
