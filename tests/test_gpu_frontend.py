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
