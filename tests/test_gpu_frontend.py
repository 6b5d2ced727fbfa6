"""GPU parity of the Rx front-end glue (opticommpy_b200.core.firFilter / decimate) against
the reference's own outputs (tests/golden/ref_frontend.npz) and the oracle."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "ref_frontend.npz")


@pytest.fixture(scope="module")
def g():
    return np.load(GOLD)


class Bag:
    def __init__(self, **kw):
        self.__dict__.update(kw)


def rel(a, b):
    return np.linalg.norm(np.asarray(a) - np.asarray(b)) / np.linalg.norm(b)


def test_fir_filter_golden(g):
    from opticommpy_b200.core import firFilter
    x = g["fir_in"]
    y = firFilter(g["fir_h_rrc"], x)
    assert y.shape == x.shape and y.dtype == x.dtype
    assert rel(y, g["fir_rrc"]) < 2e-6  # complex64 arithmetic on the device
    y = firFilter(g["fir_h_even"], x[:, 0])
    assert y.shape == g["fir_even_1d"].shape and rel(y, g["fir_even_1d"]) < 2e-6
    y = firFilter(g["fir_h_rrc"], x.real.copy())
    assert y.dtype == np.float64 and rel(y, g["fir_real_in"]) < 2e-6
    y = firFilter(g["fir_h_rrc"].astype(np.float32), x.astype(np.complex64))
    assert y.dtype == np.complex64 and rel(y, g["fir_c64"]) < 2e-6


def test_fir_filter_long_signal_vs_oracle():
    """2^20 samples x 2 modes through the overlap-save path, checked on slices against the direct convolution."""
    from opticommpy_b200.core import firFilter
    from oracle import frontend_oracle as fe
    rng = np.random.default_rng(4)
    x = rng.normal(size=(1 << 20, 2)) + 1j * rng.normal(size=(1 << 20, 2))
    h = rng.normal(size=401) * np.hanning(401)
    y = firFilter(h, x)
    for lo in (0, 4096 - 300, (1 << 19) - 77, (1 << 20) - 5000):  # start, a block boundary, middle, end
        seg = slice(max(lo - 400, 0), min(lo + 5400, 1 << 20))
        ref = fe.fir_filter(h, x[seg])
        a = lo - seg.start
        take = slice(a + (400 if seg.start > 0 else 0), a + 4000)
        assert rel(y[seg][take], ref[take]) < 2e-6


def test_decimate_golden(g):
    from opticommpy_b200.core import decimate
    s = g["dec_in"]
    y = decimate(s, Bag(SpSin=16, SpSout=2))
    assert y.shape == g["dec_16_2"].shape
    assert np.array_equal(y.astype(np.complex64), g["dec_16_2"].astype(np.complex64))  # pure selection of input samples
    y = decimate(s[:, 1], Bag(SpSin=16, SpSout=1))
    assert y.shape == g["dec_16_1_1d"].shape
    assert np.array_equal(y.astype(np.complex64), g["dec_16_1_1d"].astype(np.complex64))
    y = decimate(s[:4000], Bag(SpSin=4, SpSout=2))
    assert np.array_equal(y.astype(np.complex64), g["dec_4_2"].astype(np.complex64))
    with pytest.raises(ValueError):
        decimate(s[:4001], Bag(SpSin=4, SpSout=2))


def test_pnorm_vs_oracle():
    from opticommpy_b200.core import pnorm
    from oracle import rxdsp_oracle as ro
    rng = np.random.default_rng(3)
    x = 3.7 * (rng.normal(size=(100001, 2)) + 1j * rng.normal(size=(100001, 2)))
    y = pnorm(x)
    assert y.shape == x.shape and y.dtype == np.complex128
    assert rel(y, ro.pnorm(x)) < 1e-14
    assert np.mean(np.abs(y) ** 2) == pytest.approx(1.0, rel=1e-13)
    r = pnorm(x.real.copy()[:, 0])
    assert r.dtype == np.float64 and np.mean(r ** 2) == pytest.approx(1.0, rel=1e-13)


def test_symbol_sync_golden(g):
    """symbolSync on the device against the unmodified reference: the output is the transmit sequence permuted /
    rotated / rolled, so equal decisions give bit-identical arrays."""
    from opticommpy_b200.core import symbolSync
    y = symbolSync(g["sync_rx"], g["sync_tx_amp"], 2, "amp")
    assert y.dtype == g["sync_amp"].dtype and np.array_equal(y, g["sync_amp"])
    y = symbolSync(g["sync_rx"], g["sync_tx_real"], 2, "real")
    assert np.array_equal(y, g["sync_real"])
    tx0 = np.roll(g["sync_tx_amp"][:, 1], 11)
    y = symbolSync(g["sync_rx"][:, 0], np.roll(tx0, 9), 2, "amp")
    assert y.shape == g["sync_amp_1d"].shape and np.array_equal(y, g["sync_amp_1d"])


def test_symbol_sync_large_vs_oracle():
    """2^16 symbols x 2 modes at 1 SpS, different rx / tx lengths, all four rotations."""
    from opticommpy_b200.core import symbolSync
    from oracle import frontend_oracle as fo
    rng = np.random.default_rng(5)
    n = 1 << 16
    c = np.array([a + 1j * b for a in (-3, -1, 1, 3) for b in (-3, -1, 1, 3)]) / np.sqrt(10)
    tx = c[rng.integers(0, 16, size=(n, 2))]
    rx = tx + 0.1 * (rng.normal(size=tx.shape) + 1j * rng.normal(size=tx.shape))
    for rotx, roty in ((1, -1), (1j, -1j)):
        txs = np.stack([rotx * np.roll(tx[:, 1], 1234), np.conj(roty * np.roll(tx[:, 0], -777))], axis=1)
        assert np.array_equal(symbolSync(rx[: n - 100], txs, 1, "real"), fo.symbol_sync(rx[: n - 100], txs, 1, "real"))
        assert np.array_equal(symbolSync(rx, txs, 1, "amp"), fo.symbol_sync(rx, txs, 1, "amp"))


def test_pdm_coherent_receiver_golden(g):
    """pdmCoherentReceiver (ideal photodiodes) and delaySignal against the unmodified reference; complex64 arithmetic on
    the device: rel. L2 <= 2e-6 (1e-5 where a 512-tap fractional-delay filter is involved)."""
    from opticommpy_b200.core import delaySignal
    from opticommpy_b200.devices import pdmCoherentReceiver
    Es, Elo, Fs = g["fe_Es"], g["fe_Elo"], 64e9
    pd = Bag(B=32e9, Fs=Fs, ideal=True, seed=1011)
    y = pdmCoherentReceiver(Es, Elo, Bag(Fs=Fs, polRotation=np.pi / 3, pdl=0, polDelay=0), pd)
    assert y.shape == g["fe_rot"].shape and y.dtype == np.complex128 and rel(y, g["fe_rot"]) < 2e-6
    y = pdmCoherentReceiver(Es, Elo, Bag(Fs=Fs, polRotation=0.4, pdl=1.5, polDelay=0, phaseImbX=3 * np.pi / 180,
                                         phaseImbY=-2 * np.pi / 180, ampImbX=0.5, ampImbY=-0.3),
                            Bag(B=32e9, Fs=Fs, ideal=True, R=0.8))
    assert rel(y, g["fe_imb"]) < 2e-6
    y = pdmCoherentReceiver(Es, Elo, Bag(Fs=Fs, polRotation=np.pi / 3, pdl=0, polDelay=3 / 32e9), pd)
    assert rel(y, g["fe_delay"]) < 1e-5
    y = pdmCoherentReceiver(Es, Elo, Bag(Fs=Fs, polRotation=0.2, timeSkewX=4e-12, timeSkewY=-6e-12), pd)
    assert rel(y, g["fe_skew"]) < 1e-5
    y = pdmCoherentReceiver(Es[:, 0].copy(), Elo, Bag(Fs=Fs, polRotation=0.2, timeSkewX=4e-12, timeSkewY=-6e-12), pd)
    assert rel(y, g["fe_1pol"]) < 1e-5
    y = delaySignal(Es[:, 0].copy(), 7.3e-12, Fs)
    assert y.dtype == np.complex128 and rel(y, g["delay_c"]) < 1e-5
    y = delaySignal(Es[:, 1].real.copy(), -2.6e-11, Fs)
    assert np.isrealobj(y) and rel(y, g["delay_r"]) < 1e-5
    with pytest.raises(NotImplementedError):
        pdmCoherentReceiver(Es, Elo, Bag(Fs=Fs), Bag(Fs=Fs, ideal=False))
    with pytest.raises(AssertionError):
        pdmCoherentReceiver(Es, Elo[:-1], Bag(Fs=Fs), pd)


def test_cw_lo_downshift_equals_explicit_lo():
    """The on-the-fly CW LO of the device entry (the WDM channel down-shift) equals an explicit noiseless LO field."""
    import torch
    from opticommpy_b200.devices import pdm_frontend_rows_device
    rng = np.random.default_rng(1)
    n, Fs, f0, P = 1 << 16, 512e9, -75e9, 1e-2
    E = ((rng.normal(size=(2, n)) + 1j * rng.normal(size=(2, n))) * 0.02).astype(np.complex64)
    d_E = torch.view_as_real(torch.from_numpy(E).cuda()).contiguous()
    lo = (np.sqrt(P) * np.exp(2j * np.pi * f0 * np.arange(n) / Fs)).astype(np.complex64)
    d_lo = torch.view_as_real(torch.from_numpy(lo).cuda()).contiguous()
    fe_ = Bag(Fs=Fs, polRotation=0.3)
    a, _ = pdm_frontend_rows_device(d_E, fe_, d_Elo=d_lo)
    b, _ = pdm_frontend_rows_device(d_E, fe_, d_Elo=None, lo_power_w=P, lo_freq_shift=f0)
    a, b = a.cpu().numpy(), b.cpu().numpy()
    assert np.linalg.norm(a - b) / np.linalg.norm(a) < 2e-6
