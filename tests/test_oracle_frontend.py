"""The numpy restatement of the Rx front-end glue (oracle/frontend_oracle.py) against outputs of the unmodified
reference (tests/golden/ref_frontend.npz, produced by tests/golden/make_golden_frontend.py)."""
import os

import numpy as np
import pytest

from oracle import frontend_oracle as fe

GOLD = os.path.join(os.path.dirname(__file__), "golden", "ref_frontend.npz")


@pytest.fixture(scope="module")
def g():
    return np.load(GOLD)


def rel(a, b):
    return np.linalg.norm(np.asarray(a) - np.asarray(b)) / np.linalg.norm(b)


def test_fir_filter_matches_reference(g):
    x = g["fir_in"]
    assert rel(fe.fir_filter(g["fir_h_rrc"], x), g["fir_rrc"]) < 1e-12
    y = fe.fir_filter(g["fir_h_even"], x[:, 0])
    assert y.shape == g["fir_even_1d"].shape and rel(y, g["fir_even_1d"]) < 1e-12
    y = fe.fir_filter(g["fir_h_rrc"], x.real.copy())
    assert y.dtype == g["fir_real_in"].dtype and rel(y, g["fir_real_in"]) < 1e-12
    y = fe.fir_filter(g["fir_h_rrc"].astype(np.float32), x.astype(np.complex64))
    assert y.dtype == np.complex64 and rel(y, g["fir_c64"]) < 1e-5


def test_decimate_matches_reference(g):
    s = g["dec_in"]
    y, d = fe.decimate(s, 16, 2)
    assert np.array_equal(y, g["dec_16_2"]) and d[0] != d[1]  # the two modes sit at different sampling phases
    y, _ = fe.decimate(s[:, 1], 16, 1)
    assert y.shape == g["dec_16_1_1d"].shape and np.array_equal(y, g["dec_16_1_1d"])
    y, _ = fe.decimate(s[:4000], 4, 2)
    assert np.array_equal(y, g["dec_4_2"])


@pytest.mark.parametrize("K", [1, 2, 7, 64, 257])
def test_fir_oracle_equals_scipy_same_mode(K):
    """Independent of the golden file: the oracle's centred slice is scipy's mode='same' for any filter length."""
    from scipy import signal
    rng = np.random.default_rng(K)
    x = rng.normal(size=(1000, 2)) + 1j * rng.normal(size=(1000, 2))
    h = rng.normal(size=K) + 1j * rng.normal(size=K)
    ref = np.stack([signal.fftconvolve(x[:, n], h, mode="same") for n in range(2)], axis=1)
    assert rel(fe.fir_filter(h, x), ref) < 1e-12


def test_decimate_oracle_picks_the_eye_opening():
    """Triangular transitions at 8 SpS peak at one phase per symbol; after a circular shift by 3 samples the
    maximum-variance phase moves by 3 and the decimated sequence is the symbol sequence again."""
    rng = np.random.default_rng(1)
    lv = np.array([-3.0, -1.0, 1.0, 3.0])
    sym = rng.choice(lv, size=500) + 1j * rng.choice(lv, size=500)
    tri = np.convolve(np.repeat(sym, 8), np.ones(8) / 8)[: 8 * 500]   # full-amplitude sample at phase 7 of every symbol
    y0, d0 = fe.decimate(tri, 8, 1)
    assert d0 == [7] and np.allclose(y0, sym, atol=1e-12)
    y3, d3 = fe.decimate(np.roll(tri, 3), 8, 1)
    assert d3 == [(7 + 3) % 8]
    assert np.allclose(y3[1:], sym[:-1], atol=1e-12)   # phase 2 of symbol slot k holds symbol k-1


def test_symbol_sync_matches_reference(g):
    """symbolSync (core.py:552-675) restated in the oracle — test infrastructure for the next §8f row."""
    y = fe.symbol_sync(g["sync_rx"], g["sync_tx_amp"], 2, "amp")
    assert np.array_equal(y, g["sync_amp"])
    y = fe.symbol_sync(g["sync_rx"], g["sync_tx_real"], 2, "real")
    assert np.allclose(y, g["sync_real"], atol=1e-12)
    tx0 = np.roll(g["sync_tx_amp"][:, 1], 11)          # the generator built column 1 as roll(tx0, -11)
    y = fe.symbol_sync(g["sync_rx"][:, 0], np.roll(tx0, 9), 2, "amp")
    assert y.shape == g["sync_amp_1d"].shape and np.array_equal(y, g["sync_amp_1d"])


def test_pdm_frontend_and_delay_match_reference(g):
    """pdmCoherentReceiver with ideal photodiodes and delaySignal: oracle against the unmodified reference."""
    Es, Elo, Fs = g["fe_Es"], g["fe_Elo"], 64e9
    y = fe.pdm_coherent_receiver_ideal(Es, Elo, Fs, polRotation=np.pi / 3)
    assert np.allclose(y, g["fe_rot"], rtol=0, atol=1e-12 * np.abs(g["fe_rot"]).max())
    y = fe.pdm_coherent_receiver_ideal(Es, Elo, Fs, polRotation=0.4, pdl=1.5, phaseImb=(3 * np.pi / 180, -2 * np.pi / 180),
                                       ampImb=(0.5, -0.3), R=0.8)
    assert np.allclose(y, g["fe_imb"], rtol=0, atol=1e-12 * np.abs(g["fe_imb"]).max())
    y = fe.pdm_coherent_receiver_ideal(Es, Elo, Fs, polRotation=np.pi / 3, polDelay=3 / 32e9)
    assert np.allclose(y, g["fe_delay"], rtol=0, atol=1e-10 * np.abs(g["fe_delay"]).max())
    y = fe.pdm_coherent_receiver_ideal(Es, Elo, Fs, polRotation=0.2, timeSkew=(4e-12, -6e-12))
    assert np.allclose(y, g["fe_skew"], rtol=0, atol=1e-10 * np.abs(g["fe_skew"]).max())
    y = fe.pdm_coherent_receiver_ideal(Es[:, 0], Elo, Fs, polRotation=0.2, timeSkew=(4e-12, -6e-12))
    assert np.allclose(y, g["fe_1pol"], rtol=0, atol=1e-10 * np.abs(g["fe_skew"]).max())
    assert np.allclose(fe.delay_signal(Es[:, 0], 7.3e-12, Fs), g["delay_c"], rtol=0, atol=1e-12)
    yr = fe.delay_signal(Es[:, 1].real, -2.6e-11, Fs)
    assert np.isrealobj(yr) and np.allclose(yr, g["delay_r"], rtol=0, atol=1e-12)
