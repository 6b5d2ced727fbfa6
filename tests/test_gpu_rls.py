"""GPU parity of the 'rls' / 'dd-rls' equalizer stages: 'rls' against the reference's own
outputs (tests/golden/ref_rls.npz), 'dd-rls' against the oracle (the reference runs it on an uninitialised matrix,
equalization.py:447-451, so it has no golden vector)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "ref_rls.npz")
CASES = {
    "rls": dict(alg=["rls"], mu=[1e-3], L=[1500], lambdaRLS=0.99),
    "nlms_rls": dict(alg=["nlms", "rls"], mu=[5e-3, 1e-3], L=[500, 1000], lambdaRLS=0.995),
    "rls_store": dict(alg=["rls"], mu=[1e-3], L=[400], lambdaRLS=0.98, storeCoeff=True),
    "rls35": dict(alg=["nlms", "rls"], mu=[5e-3, 1e-3], L=[300, 1200], lambdaRLS=0.995, nTaps=35),  # two matrix rows per lane
}


class Bag:
    def __init__(self, **kw):
        self.__dict__.update(kw)


@pytest.fixture(scope="module")
def g():
    return np.load(GOLD)


def rel(a, b):
    return np.linalg.norm(np.asarray(a) - np.asarray(b)) / np.linalg.norm(b)


@pytest.mark.parametrize("tag", sorted(CASES))
def test_rls_golden(g, tag):
    from opticommpy_b200.equalization import mimoAdaptEqualizer
    p = Bag(**{**dict(nTaps=11, SpS=2, M=16, constType="qam", prgsBar=False, returnResults=True), **CASES[tag]})
    y, H, err, Hiter = mimoAdaptEqualizer(g["in"], p, g["ref"])
    assert y.dtype == np.complex64 and y.shape == g[f"{tag}_y"].shape
    assert rel(y, g[f"{tag}_y"]) < 5e-4      # both sides run the matrix recursion in complex64
    assert rel(H, g[f"{tag}_H"]) < 5e-4
    n = sum(CASES[tag]["L"])
    assert rel(err.real[:, :n], g[f"{tag}_err"].real[:, :n]) < 5e-3
    assert Hiter.shape == g[f"{tag}_Hiter"].shape
    assert rel(Hiter, g[f"{tag}_Hiter"]) < 5e-4
    c = np.unique(np.round(g["ref"], 6))
    dec = lambda z: np.argmin(np.abs(z[..., None] - c), axis=-1)
    assert np.array_equal(dec(y[200:n]), dec(g[f"{tag}_y"][200:n]))  # identical hard decisions once converged


def test_dd_rls_vs_oracle(g):
    from opticommpy_b200.equalization import mimoAdaptEqualizer
    from opticommpy_b200.modulation import grayMapping
    from oracle import rxdsp_oracle as ro
    kw = dict(alg=["nlms", "dd-rls"], mu=[5e-3, 1e-3], L=[500, 1000], lambdaRLS=0.99)
    y = mimoAdaptEqualizer(g["in"], Bag(nTaps=11, SpS=2, M=16, constType="qam", prgsBar=False, **kw), g["ref"])
    ref = ro.mimo_adapt_equalizer(g["in"], g["ref"], grayMapping(16, "qam"), nTaps=11, SpS=2, alg=("nlms", "dd-rls"),
                                  mu=(5e-3, 1e-3), L=(500, 1000), lambdaRLS=0.99)[0]
    assert rel(y, ref) < 5e-4


def test_rls_limits(g):
    from opticommpy_b200.equalization import mimoAdaptEqualizer
    with pytest.raises(NotImplementedError):
        mimoAdaptEqualizer(g["in"], Bag(nTaps=65, SpS=2, M=16, alg=["rls"], mu=[1e-3], L=[100], prgsBar=False), g["ref"])


def test_dd_rls_45_taps_vs_oracle(g):
    """'dd-rls' beyond 32 taps (pinned to the oracle only, like every dd-rls case)."""
    from opticommpy_b200.equalization import mimoAdaptEqualizer
    from opticommpy_b200.modulation import grayMapping
    from oracle import rxdsp_oracle as ro
    kw = dict(alg=["nlms", "dd-rls"], mu=[5e-3, 1e-3], L=[400, 800], lambdaRLS=0.995)
    y = mimoAdaptEqualizer(g["in"], Bag(nTaps=45, SpS=2, M=16, constType="qam", prgsBar=False, **kw), g["ref"])
    ref = ro.mimo_adapt_equalizer(g["in"], g["ref"], grayMapping(16, "qam"), nTaps=45, SpS=2, alg=("nlms", "dd-rls"),
                                  mu=(5e-3, 1e-3), L=(400, 800), lambdaRLS=0.995)[0]
    assert rel(y[:1200], ref[:1200]) < 2e-3
