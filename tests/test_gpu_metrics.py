"""GPU parity of the hard decisions and error counting (minEuclid, demodulateGray, fastBERcalc) vs the
reference's golden vectors (tests/golden/ref_metrics.npz) and the CPU oracle.  Indices, bits and error
counts: bit-exact.  SNR estimate: rtol 1e-12 for complex128 inputs; 1e-5 for complex64 inputs (the reference
then computes its means in float32, the device in float64)."""
import numpy as np
import pytest

from conftest import Bag

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def api():
    from opticommpy_b200 import _cabi
    _cabi.require_cuda()
    from opticommpy_b200 import metrics, modulation
    return Bag(minEuclid=modulation.minEuclid, demod=modulation.demodulateGray, gray=modulation.grayMapping,
               ber=metrics.fastBERcalc)


def test_min_euclid_golden(api, golden_metrics):
    g = golden_metrics
    c16 = api.gray(16, "qam")
    idx = api.minEuclid(g["me_qam16_in"], c16)
    assert idx.dtype == np.int64
    np.testing.assert_array_equal(idx, g["me_qam16_idx"])
    np.testing.assert_array_equal(api.minEuclid(g["me_ties_in"], c16), g["me_ties_idx"])  # first index on ties
    np.testing.assert_array_equal(api.minEuclid(g["me_psk8_in"], api.gray(8, "psk")), g["me_psk8_idx"])
    # the reference's own unit cases (tests/test_modulation.py:79-92)
    np.testing.assert_array_equal(api.minEuclid(np.array([1 + 1j, 2 + 2j, 3 + 3j]), np.array([1 + 1j, 3 + 3j, 2 + 2j])),
                                  [0, 2, 1])
    rng = np.random.default_rng(0)
    ind = rng.integers(0, 16, 200)
    np.testing.assert_array_equal(api.minEuclid(c16[ind] + 1e-3 * rng.normal(size=200), c16), ind)
    np.testing.assert_array_equal(api.minEuclid(g["me_qam16_in"].astype(np.complex64), c16), g["me_qam16_idx"])


def test_demodulate_gray_golden(api, golden_metrics):
    g = golden_metrics
    bits = api.demod(g["me_qam16_in"], 16, "qam")
    assert bits.dtype == np.int64 and bits.shape == g["dg_qam16_bits"].shape
    np.testing.assert_array_equal(bits, g["dg_qam16_bits"])
    np.testing.assert_array_equal(api.demod(g["me_psk8_in"], 8, "psk"), g["dg_psk8_bits"])
    np.testing.assert_array_equal(api.demod(g["me_pam4_in"], 4, "pam"), g["dg_pam4_bits"])


@pytest.mark.parametrize("case,M,ct,rtol", [("a", 16, "qam", 1e-12), ("b", 64, "qam", 1e-12), ("c", 8, "psk", 1e-12),
                                            ("d", 16, "qam", 1e-5), ("h", 16, "apsk", 1e-12)])
def test_fast_ber_calc_golden(api, golden_metrics, case, M, ct, rtol):
    g = golden_metrics
    ber, ser, snr = api.ber(g[f"ber_{case}_rx"], g[f"ber_{case}_tx"], M, ct)
    ref = g[f"ber_{case}"]
    np.testing.assert_array_equal(ber, ref[0])
    np.testing.assert_array_equal(ser, ref[1])
    np.testing.assert_allclose(snr, ref[2], rtol=rtol)


def test_fast_ber_calc_variants(api, golden_metrics):
    g = golden_metrics
    out = np.array(api.ber(g["ber_a_rx"], g["ber_a_tx"], 16, "qam", g["ber_e_px"]))  # shaped prior
    np.testing.assert_array_equal(out[:2], g["ber_e"][:2])
    np.testing.assert_allclose(out[2], g["ber_e"][2], rtol=1e-12)
    out = np.array(api.ber(g["ber_a_rx"].T.copy(), g["ber_a_tx"].T.copy(), 16, "qam"))  # wide orientation
    np.testing.assert_array_equal(out[:2], g["ber_f"][:2])
    ber, ser, snr = api.ber(g["ber_a_tx"], g["ber_a_tx"], 16, "qam")  # noiseless (tests/test_metrics.py:40-49)
    assert np.all(ber == 0) and np.all(ser == 0) and np.all(snr > 100)
    # device-resident inputs give the same numbers
    import torch
    d_rx, d_tx = torch.from_numpy(g["ber_a_rx"]).cuda(), torch.from_numpy(g["ber_a_tx"]).cuda()
    out = np.array(api.ber(d_rx, d_tx, 16, "qam"))
    np.testing.assert_array_equal(out[:2], g["ber_a"][:2])
    np.testing.assert_allclose(out[2], g["ber_a"][2], rtol=1e-12)


def test_fast_ber_calc_at_scale_known_counts_and_oracle(api):
    """2^21 symbols x 2 modes (cfg3 size): errors are planted at known positions with known Gray-label
    distances, so the exact counts are known; a 2^16-symbol slice is also compared with the oracle."""
    from oracle import metrics_oracle as mo
    rng = np.random.default_rng(5)
    L, M = 1 << 21, 16
    c = api.gray(M, "qam")
    cn = (c / np.sqrt(np.mean(np.abs(c) ** 2))).astype(np.complex128)
    itx = rng.integers(0, M, size=(L, 2))
    irx = itx.copy()
    pos = rng.choice(L, size=3000, replace=False)
    mask = rng.integers(1, M, size=(3000, 2))
    irx[pos] ^= mask
    tx = cn[itx]
    rx = (cn[irx] + 0.02 * (rng.normal(size=tx.shape) + 1j * rng.normal(size=tx.shape))) * (0.8 * np.exp(0.1j))
    ber, ser, snr, counts = api.ber(rx, tx, M, "qam", returnCounts=True)
    pop = np.array([bin(v).count("1") for v in range(M)])
    np.testing.assert_array_equal(counts[0], pop[mask].sum(axis=0))
    np.testing.assert_array_equal(counts[1], [3000, 3000])
    np.testing.assert_array_equal(ber, counts[0] / (L * 4))
    np.testing.assert_array_equal(ser, counts[1] / L)
    n = 1 << 16
    o_ber, o_ser, o_snr, o_counts = mo.fast_ber_calc(rx[:n], tx[:n], c, "qam", return_counts=True)
    g_ber, g_ser, g_snr, g_counts = api.ber(rx[:n], tx[:n], M, "qam", returnCounts=True)
    np.testing.assert_array_equal(g_counts, o_counts)
    np.testing.assert_allclose(g_snr, o_snr, rtol=1e-12)
    # BER decreases monotonically with the SNR (tests/test_metrics.py:62-70)
    sym = cn[rng.integers(0, M, size=50000)]
    w = rng.normal(size=50000) + 1j * rng.normal(size=50000)
    bers = [api.ber(sym + 10 ** (-s / 20) / np.sqrt(2) * w, sym, M, "qam")[0][0] for s in (8, 11, 14, 17)]
    assert np.all(np.diff(bers) <= 0) and bers[0] > 0
