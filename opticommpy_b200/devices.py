"""Drop-in mirror of ``optic.models.devices.pdmCoherentReceiver`` (optic/models/devices.py:574-668) — SURVEY.md §8f rank 3,
the front end between the fiber model and the receiver DSP — plus the device-resident entry the sharded DBP pipeline
uses.  numpy in, numpy out; the arithmetic runs on the GPU (``ocb_pdm_frontend_run``, and ``ocb_edc_run`` for the
polarisation delay / IQ skew, which the reference realises as FFT-convolved fractional-delay filters).

Only ideal photodiodes (``paramPD.ideal = True``, the setting of examples/test_WDM_transmission.ipynb) are part of this
path: the shot / thermal noise of the non-ideal model is drawn from NumPy's global legacy stream interleaved over eight
photodiodes (devices.py:329-351) and is outside the hot path.
"""
from __future__ import annotations

import ctypes as C
import logging as logg

import numpy as np

from . import _cabi, _engine
from .core import _ptr, _vp, delay_rows_device


def _iq_coeffs(ampImb_dB, phaseImb):
    """k1, k2 of iqMixing (optic/dsp/core.py:951-958)."""
    a = 10 ** (ampImb_dB / 20) - 1
    k1 = (1 - a) * np.exp(1j * phaseImb / 2) / 2 + (1 + a) * np.exp(-1j * phaseImb / 2) / 2
    k2 = (1 - a) * np.exp(-1j * phaseImb / 2) / 2 - (1 + a) * np.exp(1j * phaseImb / 2) / 2
    return complex(k1), complex(k2)


def pdm_frontend_rows_device(d_Es, paramFE, d_Elo=None, lo_power_w=0.0, lo_freq_shift=0.0, R=1.0):
    """Device-resident front end: ``d_Es`` planar rows (2, N, 2) float32 CUDA (x, y) -> (2, N, 2) detected baseband rows.
    ``d_Elo``: (N, 2) float32 LO field, or None for a noiseless CW LO of ``lo_power_w`` watts shifted by
    ``lo_freq_shift`` Hz generated inside the kernel (the channel down-shift).  ``paramFE`` as in pdmCoherentReceiver."""
    torch = _cabi.require_cuda()
    lib = _cabi.lib()
    st = _vp(_cabi.stream_ptr(torch))
    Fs = paramFE.Fs
    N = int(d_Es.shape[1])
    polRotation = getattr(paramFE, "polRotation", 0)
    pdl = getattr(paramFE, "pdl", 0)
    polDelay = getattr(paramFE, "polDelay", 0)
    k1x, k2x = _iq_coeffs(getattr(paramFE, "ampImbX", 0), getattr(paramFE, "phaseImbX", 0))
    k1y, k2y = _iq_coeffs(getattr(paramFE, "ampImbY", 0), getattr(paramFE, "phaseImbY", 0))
    skew = (getattr(paramFE, "timeSkewX", 0), getattr(paramFE, "timeSkewY", 0))
    iq = (C.c_double * 8)(k1x.real, k1x.imag, k2x.real, k2x.imag, k1y.real, k1y.imag, k2y.real, k2y.imag)
    keep = []
    if polDelay != 0:
        # The delay acts on the rotated polarisations (devices.py:651-653): rotate first with a unit front end ...
        c, s = np.cos(polRotation), np.sin(polRotation)
        Ec = torch.view_as_complex(d_Es)
        rot = torch.stack([Ec[0] * c + Ec[1] * s, -Ec[0] * s + Ec[1] * c])
        d_rot = torch.view_as_real(rot).contiguous()
        dx, k1 = delay_rows_device(d_rot[0:1], -polDelay / 2, Fs)
        dy, k2 = delay_rows_device(d_rot[1:2], polDelay / 2, Fs)
        d_Es = torch.cat([dx, dy]).contiguous()
        keep += [k1, k2]
        polRotation = 0.0  # ... the kernel then applies PDL, hybrid, photodiodes and IQ imbalance
    d_S = torch.empty_like(d_Es)
    _cabi.check(lib.ocb_pdm_frontend_run(_ptr(d_Es), _ptr(d_Elo) if d_Elo is not None else None, _ptr(d_S), N,
                                         float(polRotation), float(pdl), float(R), float(lo_power_w), float(lo_freq_shift),
                                         float(Fs), iq, st), "ocb_pdm_frontend_run")
    for p, sk in enumerate(skew):
        if sk == 0:
            # Reference quirk (reproduced): iqMixing ALWAYS passes I and Q through delaySignal (core.py:962-965), and a
            # zero delay is not the identity there — the 512-tap "delta" sits one tap off the compensated group delay
            # and the final roll(-1) wraps the leading zero to the end — so the LAST sample of each output is zero.
            d_S[p, N - 1] = 0
        else:  # iqMixing skew (core.py:962-965): I delayed by -sk/2, Q by +sk/2, each as a real signal
            s = d_S[p]
            re = torch.stack([s[:, 0], torch.zeros_like(s[:, 0])], dim=1)[None].contiguous()
            im = torch.stack([s[:, 1], torch.zeros_like(s[:, 1])], dim=1)[None].contiguous()
            dre, k1 = delay_rows_device(re, -sk / 2, Fs)
            dim, k2 = delay_rows_device(im, sk / 2, Fs)
            d_S[p, :, 0] = dre[0, :, 0]
            d_S[p, :, 1] = dim[0, :, 0]
            keep += [k1, k2]
    return d_S, keep


def pdmCoherentReceiver(Es, Elo, paramFE, paramPD=None):
    """
    Polarization multiplexed coherent optical front-end on the GPU (ideal photodiodes).

    Parameters as in the reference (devices.py:585-607): ``Es`` (N, 2) signal field (or (N,): x polarisation only),
    ``Elo`` (N,) LO field, ``paramFE``: Fs, polRotation, pdl, polDelay, phaseImbX/Y, ampImbX/Y, timeSkewX/Y;
    ``paramPD``: photodiode parameters — ``ideal`` must be True, ``R`` [1 A/W] is honoured.
    Returns ``S`` (N, 2): the down-converted x and y signals.
    """
    Es = np.asarray(Es)
    Elo = np.asarray(Elo)
    assert len(Es) == len(Elo), "Es and Elo need to have the same length"
    try:
        Fs = paramFE.Fs
    except AttributeError:
        logg.error("Simulation sampling frequency (Fs) not provided.")
        raise NameError("name 'Fs' is not defined") from None  # the reference dies the same way at :617
    if paramPD is None or not getattr(paramPD, "ideal", False):
        raise NotImplementedError("pdmCoherentReceiver on the GPU models ideal photodiodes only: set paramPD.ideal = True "
                                  "(the noisy photodiode draws from NumPy's global legacy stream, devices.py:329-351)")
    R = getattr(paramPD, "R", 1)
    assert R > 0, "PD responsivity should be a positive scalar"
    if Es.ndim == 1:  # pbs (devices.py:246-250): a single-polarisation field enters on x
        Es = np.stack([Es, np.zeros_like(Es)], axis=1)
    elif Es.shape[1] > 2:
        logg.error("E need to be a (N,2) or a (N,) np.array")
    torch = _cabi.require_cuda()
    lib = _cabi.lib()
    st = _vp(_cabi.stream_ptr(torch))
    N = len(Es)
    host = _engine.as_host_complex(Es)
    d_raw = torch.from_numpy(host.view(np.float32 if host.dtype == np.complex64 else np.float64)).to("cuda")
    d_rows = torch.empty((2, N, 2), dtype=torch.float32, device="cuda")
    _cabi.check(lib.ocb_pack_fields(_ptr(d_raw), _engine.dtype_tag(host.dtype), N, 2, 0, _ptr(d_rows), st), "ocb_pack_fields")
    d_lo = torch.from_numpy(np.ascontiguousarray(Elo.astype(np.complex64)).view(np.float32)).to("cuda")
    d_S, _keep = pdm_frontend_rows_device(d_rows, paramFE, d_Elo=d_lo, R=R)
    d_out = torch.empty((N, 2, 2), dtype=torch.float64, device="cuda")
    _cabi.check(lib.ocb_unpack_fields(_ptr(d_S), N, 2, 0, _ptr(d_out), _cabi.OCB_C128, st), "ocb_unpack_fields")
    return d_out.cpu().numpy().view(np.complex128).reshape(N, 2)  # complex128 like the reference's sI + 1j*sQ


def basicLaserModel(param=None):
    """
    Laser field with random-walk phase noise and RIN (optic/models/devices.py:729-791): P [10 dBm], lw [1 kHz],
    RIN_var [1e-20], Fs, Ns [1000], seed [None], freqShift [0 Hz].  Returns the (Ns,) complex128 field
    ``sqrt(P + dP) exp(j (2 pi freqShift t + phi_pn))``.

    A host-side input generator, kept for drop-in completeness of the receiver notebooks: both noise processes come from
    numpy's legacy generator (seed for the phase walk, seed + 73 for the RIN), so a seeded call returns the reference's
    local-oscillator field.  A noiseless CW LO needs no array at all: ``pdm_frontend_rows_device(..., d_Elo=None)``
    generates it inside the front-end kernel.
    """
    from .tx import phaseNoise
    try:
        Fs = param.Fs
    except AttributeError:
        logg.error("Simulation sampling frequency (Fs) not provided.")
        raise NameError("name 'Fs' is not defined") from None  # the reference dies on the unbound name (devices.py:779)
    P = getattr(param, "P", 10)
    lw = getattr(param, "lw", 1e3)
    RIN_var = getattr(param, "RIN_var", 1e-20)
    Ns = int(getattr(param, "Ns", 1000))
    seed = getattr(param, "seed", None)
    freqShift = getattr(param, "freqShift", 0)
    pn = phaseNoise(lw, Ns, 1 / Fs, seed)
    if seed is None:
        s = np.sqrt(RIN_var / 2)
        deltaP = np.random.normal(0, s, pn.shape) + 1j * np.random.normal(0, s, pn.shape)  # core.py:758-763, global stream
    else:
        deltaP = _engine.legacy_complex_noise(pn.shape, RIN_var, seed + 73)
    fo = 2 * np.pi * freqShift * np.arange(Ns) / Fs if freqShift != 0 else 0
    return np.sqrt(10 ** (P / 10) * 1e-3 + deltaP) * np.exp(1j * (fo + pn))
