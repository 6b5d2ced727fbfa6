"""Device-resident compositions of the hot-path stages for the sharded workloads of BASELINE.json
(configs[3]: per-channel digital back-propagation of a WDM field; configs[4]: Monte-Carlo sweep over ASE noise seeds
with on-device error counting).  Every stage is one of this package's own device entry points; the WDM field is
uploaded once per process and only scalars (or the final per-channel fields) leave the GPU.

    channel pipeline (examples/test_WDM_transmission.ipynb, cells "receiver" .. "DBP"):
        manakovSSF -> pdmCoherentReceiver (CW LO on the channel = down-shift) -> firFilter (matched filter)
                   -> decimate -> manakovDBP | edc -> mimoAdaptEqualizer -> cpr -> fastBERcalc

Reference call sites restated: optic/models/devices.py:574-668, optic/dsp/core.py:87-125, 435-491,
optic/dsp/equalization.py:36-122, 125-351, 976-1173, optic/dsp/carrierRecovery.py:37-169, optic/comm/metrics.py:111-195.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _cabi
from .carrierRecovery import cpr_bps_device
from .channels import manakov_rows_device
from .devices import pdm_frontend_rows_device
from .equalization import (_edc_taps, _parse_equalizer_args, _ptr, _to_device, edc_rows_device,
                           equalizer_stages_device)
from .metrics import fastBERcalc
from .modulation import grayMapping

_vp = C.c_void_p


class Bag:
    def __init__(self, **kw):
        self.__dict__.update(kw)

    def copy(self):
        return Bag(**self.__dict__)


def upload_field(x):
    """(N, 2) complex host field -> planar complex64 CUDA tensor (2, N) (x row, y row), one H2D copy."""
    torch = _cabi.require_cuda()
    x = np.asarray(x)
    return torch.from_numpy(np.ascontiguousarray(x.T.astype(np.complex64))).cuda()


def channel_frontend_device(rows, ch_freq, Fs, pulse, SpSin, SpSout, lo_power_w=1e-2):
    """One WDM channel out of the field ``rows`` ((2, N) complex64 CUDA): coherent detection with a noiseless CW LO tuned
    to ``ch_freq`` (the down-shift), matched filter ``pulse`` (host taps) and decimation SpSin -> SpSout at the
    maximum-variance sampling instant.  Returns a (2, N * SpSout / SpSin) complex64 CUDA tensor."""
    torch = _cabi.require_cuda()
    lib = _cabi.lib()
    st = _vp(_cabi.stream_ptr(torch))
    N = int(rows.shape[1])
    d_E = torch.view_as_real(rows).contiguous()
    d_S, keep = pdm_frontend_rows_device(d_E, Bag(Fs=Fs), d_Elo=None, lo_power_w=lo_power_w, lo_freq_shift=ch_freq)
    d_mf = torch.empty_like(d_S)
    keep2 = edc_rows_device(d_S, d_mf, np.asarray(pulse).astype(np.complex64))        # firFilter = 'same' convolution
    dec = int(SpSin // SpSout)
    Nout = (N + dec - 1) // dec
    d_y = torch.empty((2, Nout, 2), dtype=torch.float32, device="cuda")
    d_delay = torch.empty(2, dtype=torch.int32, device="cuda")
    ws_bytes = int(lib.ocb_decimate_workspace_bytes(2, int(SpSin)))
    d_ws = torch.empty(ws_bytes + 256, dtype=torch.uint8, device="cuda")
    ws_ptr = (d_ws.data_ptr() + 255) // 256 * 256
    _cabi.check(lib.ocb_decimate_run(_ptr(d_mf), _ptr(d_y), N, 2, int(SpSin), dec, _ptr(d_delay), _vp(ws_ptr), ws_bytes, st),
                "ocb_decimate_run")
    return torch.view_as_complex(d_y)


def pnorm_rows_device(rows):
    """In-place power normalisation of a complex64 CUDA tensor over all of its entries (optic/dsp/core.py:702-717)."""
    torch = _cabi.require_cuda()
    lib = _cabi.lib()
    st = _vp(_cabi.stream_ptr(torch))
    n = rows.numel() // (1 if rows.is_complex() else 2)   # complex elements (float-pair tensors carry a trailing 2)
    d128 = torch.empty((n, 2), dtype=torch.float64, device="cuda")
    _cabi.check(lib.ocb_cast_complex(_ptr(rows), _cabi.OCB_C64, _ptr(d128), _cabi.OCB_C128, n, st), "ocb_cast_complex")
    d_ws = torch.empty(4096 + 32, dtype=torch.float64, device="cuda")
    _cabi.check(lib.ocb_pnorm_run(_ptr(d128), n, _ptr(d_ws), 4096 * 8, st), "ocb_pnorm_run")
    _cabi.check(lib.ocb_cast_complex(_ptr(d128), _cabi.OCB_C128, _ptr(rows), _cabi.OCB_C64, n, st), "ocb_cast_complex")
    return rows


def dbp_channel_device(rows_wdm, ch_freq, Fs, pulse, SpSin, paramDBP, SpSout=2):
    """cfg4 unit: channel front end + manakovDBP on the device.  ``paramDBP`` as in manakovDBP with ``Fs`` of the
    decimated signal.  Returns ((2, N') complex64 CUDA tensor, DBP stats)."""
    ch = channel_frontend_device(rows_wdm, ch_freq, Fs, pulse, SpSin, SpSout).contiguous()
    stats = manakov_rows_device(ch, paramDBP, -1)
    return ch, stats


class RxRecipe:
    """Seed-independent part of the Monte-Carlo receiver: taps, constellation, equalizer setup, aligned reference."""

    def __init__(self, Fs, Rs, SpSin, pulse, Ltotal, D, Fc, symbRef, M=16, nTaps=15, mu=(5e-3, 5e-4), train_frac=0.2, B=64,
                 Nbps=25):
        self.Fs, self.Rs, self.SpSin, self.pulse, self.M = Fs, Rs, int(SpSin), np.asarray(pulse), M
        self.Fs2 = 2 * Rs
        self.h_edc, _, _ = _edc_taps(Bag(L=Ltotal, D=D, Fc=Fc, Rs=Rs, Fs=self.Fs2), self.Fs2)
        self.symbRef = np.ascontiguousarray(np.asarray(symbRef).astype(np.complex64))       # (nSymb, 2), aligned
        self.nTaps, self.mu, self.train_frac, self.B, self.Nbps = nTaps, list(mu), train_frac, B, Nbps
        c = grayMapping(M, "qam")
        self.const = c / np.sqrt(np.mean(np.abs(c) ** 2))
        self.d_ref = None


def rx_symbols_device(rows_wdm, ch_freq, recipe, timing=None):
    """cfg5 unit after the fiber: channel front end -> EDC -> pnorm -> mimoAdaptEqualizer(nlms -> dd-lms) -> cpr(bps).
    Returns the recovered symbols as an (L, 2) complex128 CUDA tensor."""
    torch = _cabi.require_cuda()
    lib = _cabi.lib()
    st = _vp(_cabi.stream_ptr(torch))
    r = recipe
    ch = channel_frontend_device(rows_wdm, ch_freq, r.Fs, r.pulse, r.SpSin, 2)
    d_in = torch.view_as_real(ch).contiguous()
    d_edc = torch.empty_like(d_in)
    keep = edc_rows_device(d_in, d_edc, r.h_edc)
    pnorm_rows_device(d_edc)
    N2 = int(d_edc.shape[1])
    nsym = N2 // 2
    ntrain = int(r.train_frac * nsym)
    pq = Bag(nTaps=r.nTaps, SpS=2, M=r.M, constType="qam", alg=["nlms", "dd-lms"], mu=r.mu, L=[ntrain, nsym - ntrain],
             prgsBar=False)
    s = _parse_equalizer_args(np.zeros((N2, 2), dtype=np.complex64), pq, r.symbRef)
    d_x = torch.zeros((1, s.nPad, 2, 2), dtype=torch.float32, device="cuda")
    _cabi.check(lib.ocb_unpack_fields(_ptr(d_edc), N2, 2, 0, _ptr(d_x, s.Lpad * 2 * 8), _cabi.OCB_C64, st), "ocb_unpack_fields")
    if r.d_ref is None:
        r.d_ref = _to_device(torch, r.symbRef.view(np.float32)).reshape(1, r.symbRef.shape[0], 2, 2)
    d_H = _to_device(torch, s.H[None].view(np.float32))
    d_y, d_e, _ = equalizer_stages_device(s, 1, d_x, r.d_ref, r.symbRef.shape[0], d_H, None)
    L = s.totalNumSymb
    d_out, d_ph, _, keep2 = cpr_bps_device(d_y, _cabi.OCB_C64, L, 2, r.const, r.B, r.Nbps, False, r.Rs, 4)
    return torch.view_as_complex(d_out.reshape(L, 2, 2))


def ber_scalars_device(d_sym, recipe, discard=2000):
    """(BER, SER, SNR[dB]) averaged over the two polarisations, counted on the device (ocb_ber_count)."""
    torch = _cabi.require_cuda()
    L = min(int(d_sym.shape[0]), recipe.symbRef.shape[0])
    ref = torch.view_as_complex(recipe.d_ref.reshape(-1, 2, 2))[:L].to(torch.complex128)
    a, b = discard, L - discard
    ber, ser, snr = fastBERcalc(d_sym[a:b].contiguous(), ref[a:b].contiguous(), recipe.M, "qam")
    return float(np.mean(ber)), float(np.mean(ser)), float(np.mean(snr))


def dbp_channels_device(rows_wdm, ch_freqs, Fs, pulse, SpSin, paramDBP, units=None, workers=None, SpSout=2):
    """cfg4 on one GPU: ``dbp_channel_device`` for the channels ``units`` (indices into ``ch_freqs``; default: all) with
    several channels in flight (``sharding.run_concurrent``: one host thread, CUDA stream and plan per worker; at N = 2^17
    one back-propagation alone uses a fraction of the SMs).  Returns ``{k: ((2, N') complex64 CUDA tensor, DBP stats)}``."""
    from .sharding import run_concurrent
    units = list(range(len(ch_freqs))) if units is None else list(units)
    return run_concurrent(lambda k: dbp_channel_device(rows_wdm, float(ch_freqs[k]), Fs, pulse, SpSin, paramDBP, SpSout),
                          units, workers)


def monte_carlo_ber_device(rows0, seeds, param, ch_freq, recipe, workers=None, discard=2000):
    """cfg5 on one GPU: for every ASE-noise seed in ``seeds`` propagate a private copy of the noiseless transmit field
    ``rows0`` ((2, N) complex64 CUDA) with ``param`` (``param.seed`` replaced by the seed: on-device Philox streams), run the
    receiver of ``recipe`` on channel ``ch_freq`` and count errors on the device; several seeds in flight
    (``sharding.run_concurrent``).  Returns ``{seed: (float64 CUDA tensor [BER, SER, SNR dB], executed SSFM steps)}``."""
    torch = _cabi.require_cuda()
    from .sharding import run_concurrent
    if recipe.d_ref is None:  # created once, before the workers start
        recipe.d_ref = _to_device(torch, recipe.symbRef.view(np.float32)).reshape(1, recipe.symbRef.shape[0], 2, 2)

    def unit(seed):
        r = rows0.clone()
        p = Bag(**param.__dict__)
        p.seed = int(seed)
        st = manakov_rows_device(r, p, +1)
        d_sym = rx_symbols_device(r, float(ch_freq), recipe)
        return torch.tensor(ber_scalars_device(d_sym, recipe, discard), dtype=torch.float64, device=r.device), st["steps"]

    return run_concurrent(unit, list(seeds), workers)
