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
