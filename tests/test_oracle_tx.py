"""The transmitter oracle (oracle/tx_oracle.py) against golden vectors of the unmodified reference
(tests/golden/ref_tx.npz, made by tests/golden/make_golden_tx.py): symbolSource, pulseShape, phaseNoise, iqm, simpleWDMTx."""
import os

import numpy as np
import pytest

from conftest import ROOT, rel_l2
from oracle import tx_oracle as to


@pytest.fixture(scope="module")
def gtx():
    with np.load(os.path.join(ROOT, "tests", "golden", "ref_tx.npz")) as z:
        return {k: z[k] for k in z.files}


def test_symbol_source_draws_equal_the_reference(gtx):
    for name, M, ct, dist, sf in (("qam16", 16, "qam", "uniform", 0.0), ("qam64mb", 64, "qam", "maxwell-boltzmann", 0.05),
                                  ("psk8", 8, "psk", "uniform", 0.0), ("pam4", 4, "pam", "uniform", 0.0)):
        got = to.symbol_source(3000, M, ct, 99, sf, dist)
        assert np.allclose(got, gtx[f"src_{name}"], rtol=0, atol=1e-15), name


def test_pulse_phase_noise_iqm(gtx):
    assert np.allclose(to.pulse_shape("rrc", 16, 1024, 0.01), gtx["pulse_rrc16"], rtol=1e-12, atol=1e-16)
    assert np.allclose(to.pulse_shape("rrc", 8, 257, 0.1), gtx["pulse_rrc8"], rtol=1e-12, atol=1e-16)
    assert np.array_equal(to.pulse_shape("rect", 8, 0, 0.0), gtx["pulse_rect"])
    assert np.allclose(to.phase_noise(100e3, 4096, 1 / 512e9, 5), gtx["pn"], rtol=1e-12, atol=1e-15)
    assert rel_l2(to.iqm(np.exp(1j * gtx["pn"][:2000]), gtx["iqm_u"]), gtx["iqm_out"]) < 1e-14


def test_simple_wdm_tx_equals_the_reference(gtx):
    cases = {
        "dp5": dict(M=16, Rs=32e9, SpS=8, nBits=4 * 2048, nFilterTaps=1024, pulseRollOff=0.01, powerPerChannel=-2.0, nChannels=5,
                    wdmGridSpacing=37.5e9, nPolModes=2, seed=321),
        "sp3": dict(M=4, Rs=10e9, SpS=16, nBits=2 * 1024, nFilterTaps=256, pulseRollOff=0.1, powerPerChannel=[-1.0, 0.0, 1.5],
                    nChannels=3, wdmGridSpacing=25e9, nPolModes=1, seed=17, laserLinewidth=100e3),
        "dp4": dict(M=16, Rs=32e9, SpS=4, nBits=4 * 1024, nFilterTaps=128, pulseRollOff=0.2, powerPerChannel=0.0, nChannels=4,
                    wdmGridSpacing=40e9, nPolModes=2, seed=5, mzmScale=0.25),
    }
    for tag, kw in cases.items():
        sig, symb, grid = to.simple_wdm_tx(**kw)
        assert np.allclose(grid, gtx[f"tx_{tag}_grid"]), tag
        assert np.allclose(symb, gtx[f"tx_{tag}_symb"], rtol=0, atol=1e-15), tag
        assert rel_l2(sig, gtx[f"tx_{tag}_sig"]) < 1e-12, tag
