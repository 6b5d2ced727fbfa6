"""The float64 restatement of the RLS stages (oracle/rxdsp_oracle.py:rls_stage) against outputs of the unmodified
reference (tests/golden/ref_rls.npz, produced by tests/golden/make_golden_rls.py; the reference runs in complex64)."""
import os

import numpy as np
import pytest

from oracle import rxdsp_oracle as ro

GOLD = os.path.join(os.path.dirname(__file__), "golden", "ref_rls.npz")
CASES = {
    "rls": dict(alg=("rls",), mu=(1e-3,), L=(1500,), lambdaRLS=0.99),
    "nlms_rls": dict(alg=("nlms", "rls"), mu=(5e-3, 1e-3), L=(500, 1000), lambdaRLS=0.995),
    "rls_store": dict(alg=("rls",), mu=(1e-3,), L=(400,), lambdaRLS=0.98, storeCoeff=True),
    "rls35": dict(alg=("nlms", "rls"), mu=(5e-3, 1e-3), L=(300, 1200), lambdaRLS=0.995, nTaps=35),
}


@pytest.fixture(scope="module")
def g():
    return np.load(GOLD)


def rel(a, b):
    return np.linalg.norm(np.asarray(a) - np.asarray(b)) / np.linalg.norm(b)


@pytest.mark.parametrize("tag", sorted(CASES))
def test_rls_oracle_matches_reference(g, tag):
    from opticommpy_b200.modulation import grayMapping
    kw = dict(CASES[tag])
    y, H, _, err, Hiter = ro.mimo_adapt_equalizer(g["in"], g["ref"], grayMapping(16, "qam"), nTaps=kw.pop("nTaps", 11), SpS=2, **kw)
    assert y.shape == g[f"{tag}_y"].shape
    assert rel(y, g[f"{tag}_y"]) < 2e-4          # complex64 matrix recursion in the reference vs float64 here
    assert rel(H, g[f"{tag}_H"]) < 2e-4
    n = sum(CASES[tag]["L"])
    assert rel(err[:, :n], g[f"{tag}_err"].real[:, :n]) < 2e-3
    if CASES[tag].get("storeCoeff"):
        assert Hiter.shape == g[f"{tag}_Hiter"].shape and rel(Hiter, g[f"{tag}_Hiter"]) < 2e-4


def test_rls_oracle_converges_on_a_clean_channel():
    """Sanity of the restated recursion itself: on a noiseless 2x2 mixing channel RLS drives the error to ~0
    within a few hundred symbols, and decision-directed RLS started from those taps keeps it there."""
    from opticommpy_b200.modulation import grayMapping
    rng = np.random.default_rng(8)
    c = grayMapping(4, "qam").astype(np.complex128)
    c /= np.sqrt(np.mean(np.abs(c) ** 2))
    sym = c[rng.integers(0, 4, size=(1200, 2))]
    th = 0.6
    x = np.repeat(sym, 2, axis=0) @ np.array([[np.cos(th), -np.sin(th)], [np.sin(th), np.cos(th)]]).T
    y, H, _, err, _ = ro.mimo_adapt_equalizer(x, sym, grayMapping(4, "qam"), nTaps=5, SpS=2, alg=("rls", "dd-rls"),
                                              mu=(1e-3, 1e-3), L=(600, 600), lambdaRLS=0.99)
    assert np.max(err[:, 300:600]) < 1e-4      # trained stage (squared error; the input is rounded to complex64)
    assert np.max(err[:, 700:1200]) < 1e-4     # decision-directed stage stays locked
    assert np.max(np.abs(y[300:1200] - sym[300:1200])) < 1e-2
