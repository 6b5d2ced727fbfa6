"""Pin oracle/metrics_oracle.py (decisions, Gray demapping, BER/SER/SNR) against the unmodified reference's
outputs in tests/golden/ref_metrics.npz.  CPU-only.  Indices, bits and error counts must be identical;
SNR agrees to rounding (complex64 inputs: to the reference's own float32 rounding)."""
import numpy as np
import pytest

from oracle import metrics_oracle as mo
from opticommpy_b200.modulation import grayMapping


def test_min_euclid_and_bits(golden_metrics):
    g = golden_metrics
    c16 = grayMapping(16, "qam")
    np.testing.assert_array_equal(mo.min_euclid(g["me_qam16_in"], c16), g["me_qam16_idx"])
    np.testing.assert_array_equal(mo.min_euclid(g["me_ties_in"], c16), g["me_ties_idx"])
    np.testing.assert_array_equal(mo.demodulate_gray(g["me_qam16_in"], c16), g["dg_qam16_bits"])
    c8 = grayMapping(8, "psk")
    np.testing.assert_array_equal(mo.min_euclid(g["me_psk8_in"], c8), g["me_psk8_idx"])
    np.testing.assert_array_equal(mo.demodulate_gray(g["me_psk8_in"], c8), g["dg_psk8_bits"])
    np.testing.assert_array_equal(mo.demodulate_gray(g["me_pam4_in"], grayMapping(4, "pam")), g["dg_pam4_bits"])


@pytest.mark.parametrize("case,M,ct,rtol", [("a", 16, "qam", 1e-12), ("b", 64, "qam", 1e-12), ("c", 8, "psk", 1e-12),
                                            ("d", 16, "qam", 1e-5), ("h", 16, "apsk", 1e-12)])
def test_fast_ber_calc(golden_metrics, case, M, ct, rtol):
    g = golden_metrics
    ber, ser, snr = mo.fast_ber_calc(g[f"ber_{case}_rx"], g[f"ber_{case}_tx"], grayMapping(M, ct), ct)
    ref = g[f"ber_{case}"]
    np.testing.assert_array_equal(ber, ref[0])
    np.testing.assert_array_equal(ser, ref[1])
    np.testing.assert_allclose(snr, ref[2], rtol=rtol)


def test_fast_ber_calc_variants(golden_metrics):
    g = golden_metrics
    c = grayMapping(16, "qam")
    out = np.array(mo.fast_ber_calc(g["ber_a_rx"], g["ber_a_tx"], c, "qam", g["ber_e_px"]))
    np.testing.assert_array_equal(out[:2], g["ber_e"][:2])
    np.testing.assert_allclose(out[2], g["ber_e"][2], rtol=1e-12)
    out = np.array(mo.fast_ber_calc(g["ber_a_rx"].T, g["ber_a_tx"].T, c, "qam"))  # wide orientation
    np.testing.assert_array_equal(out[:2], g["ber_f"][:2])
    out = np.array(mo.fast_ber_calc(g["ber_a_tx"], g["ber_a_tx"], c, "qam"))  # error-free
    assert np.all(out[:2] == 0) and np.all(out[2] > 100)
