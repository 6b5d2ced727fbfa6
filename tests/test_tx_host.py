"""Host-side pieces of the transmitter mirror (opticommpy_b200.tx): the symbol draw, the pulse taps and the phase-noise walk
must equal the unmodified reference's (tests/golden/ref_tx.npz) — they define WHICH symbols a seeded call transmits."""
import os

import numpy as np
import pytest

from conftest import ROOT, Bag
from opticommpy_b200.tx import phaseNoise, pulseShape, symbolSource


@pytest.fixture(scope="module")
def gtx():
    with np.load(os.path.join(ROOT, "tests", "golden", "ref_tx.npz")) as z:
        return {k: z[k] for k in z.files}


def test_symbol_source(gtx):
    for name, M, ct, dist, sf in (("qam16", 16, "qam", "uniform", 0.0), ("qam64mb", 64, "qam", "maxwell-boltzmann", 0.05),
                                  ("psk8", 8, "psk", "uniform", 0.0), ("pam4", 4, "pam", "uniform", 0.0)):
        got = symbolSource(Bag(nSymbols=3000, M=M, constType=ct, dist=dist, shapingFactor=sf, seed=99))
        assert np.allclose(got, gtx[f"src_{name}"], rtol=0, atol=1e-15), name
    a = symbolSource(Bag(nSymbols=50, M=16, seed=None))
    b = symbolSource(Bag(nSymbols=50, M=16, seed=None))
    assert a.shape == (50,) and not np.array_equal(a, b)            # unseeded draws continue the global stream
    with pytest.raises(ValueError):
        symbolSource(Bag(nSymbols=5, M=16, constType="star"))


def test_pulse_shape_and_phase_noise(gtx):
    assert np.allclose(pulseShape(Bag(pulseType="rrc", SpS=16, nFilterTaps=1024, rollOff=0.01)), gtx["pulse_rrc16"], rtol=1e-12, atol=1e-16)
    assert np.allclose(pulseShape(Bag(pulseType="rrc", SpS=8, nFilterTaps=257, rollOff=0.1)), gtx["pulse_rrc8"], rtol=1e-12, atol=1e-16)
    assert np.array_equal(pulseShape(Bag(pulseType="rect", SpS=8)), gtx["pulse_rect"])
    assert abs(np.sum(pulseShape(Bag(pulseType="rc", SpS=4, nFilterTaps=64, rollOff=0.25))) - 1) < 1e-12
    with pytest.raises(ValueError):
        pulseShape(Bag(pulseType="duobinary"))
    assert np.allclose(phaseNoise(100e3, 4096, 1 / 512e9, seed=5), gtx["pn"], rtol=1e-12, atol=1e-15)
    assert np.array_equal(phaseNoise(0.0, 7, 1e-12, seed=1), np.zeros(7))


def test_basic_laser_model(gtx):
    from opticommpy_b200.devices import basicLaserModel
    got = basicLaserModel(Bag(P=10, lw=100e3, RIN_var=1e-20, Fs=512e9, Ns=5000, seed=789, freqShift=37.5e9 - 128e6))
    assert got.dtype == np.complex128 and np.allclose(got, gtx["laser_pn"], rtol=1e-12, atol=1e-15)
    cw = basicLaserModel(Bag(P=7, lw=0.0, RIN_var=0, Fs=64e9, Ns=1000, seed=1))
    assert np.allclose(cw, gtx["laser_cw"], rtol=1e-14, atol=0)
    with pytest.raises(NameError):
        basicLaserModel(Bag(P=0))
