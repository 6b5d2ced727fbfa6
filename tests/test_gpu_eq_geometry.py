"""Latency-mode equalizer kernel (one warp per output mode, staged chunks, rotating window register sets) against the
float64 CPU oracle where its bookkeeping can go wrong: stage lengths around the chunk (126) and unroll (3) boundaries and
very short stages, every algorithm of the family, 1 / 2 / 4 modes, 1 or 2 samples per symbol; outputs, taps and squared
errors within the stated complex64 tolerance, identical hard decisions on the converged part."""
import numpy as np
import pytest

from conftest import Bag, rel_l2

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eq():
    from opticommpy_b200 import _cabi
    _cabi.require_cuda()
    from opticommpy_b200.equalization import mimoAdaptEqualizer
    return mimoAdaptEqualizer


def _case(golden, nmodes):
    x = golden["eq_in"]
    ref = golden["eq_ref"]
    if nmodes == 1:
        return x[:, :1].copy(), ref[:, :1].copy()
    if nmodes == 4:
        rng = np.random.default_rng(4)
        x4 = np.concatenate([x, x[::-1] * np.exp(0.5j)], axis=1) + 0.01 * rng.normal(size=(len(x), 4))
        return x4, np.concatenate([ref, ref[::-1]], axis=1)
    return x, ref


@pytest.mark.parametrize("L", [1, 2, 3, 5, 125, 126, 127, 128, 129, 131, 252, 253, 259, 1000])
def test_lengths_around_chunk_and_unroll_boundaries(eq, golden, L):
    from oracle import rxdsp_oracle as ro
    x, _ = _case(golden, 2)
    p = Bag(nTaps=31, SpS=2, M=16, constType="qam", alg=["cma"], mu=[2e-3], L=[L], prgsBar=False, returnResults=True)
    y, H, err, _ = eq(x, p)
    yo, Ho, _, eo, _ = ro.mimo_adapt_equalizer(x, None, golden["const_qam16"], nTaps=31, SpS=2, alg=["cma"], mu=[2e-3], L=[L])
    assert rel_l2(y[:L], yo[:L]) < 1e-4 and np.all(y[L:] == 0)
    assert rel_l2(H, Ho) < 1e-4
    assert rel_l2(err[:L], eo[:L]) < 1e-3


@pytest.mark.parametrize("alg,mu,nmodes,ntaps,sps", [
    (["cma", "rde"], [1e-3, 1e-3], 2, 31, 2),
    (["nlms", "dd-lms"], [5e-3, 1e-3], 2, 32, 2),
    (["da-rde", "rde"], [1e-3, 5e-4], 2, 15, 2),
    (["nlms", "static"], [5e-3, 0.0], 2, 7, 1),
    (["cma"], [1e-3], 1, 21, 2),
    (["nlms"], [2e-3], 4, 9, 2),
])
def test_algorithms_modes_geometries_vs_oracle(eq, golden, alg, mu, nmodes, ntaps, sps):
    from oracle import rxdsp_oracle as ro
    x, ref = _case(golden, nmodes)
    nsym = len(x) // sps if sps == 2 else 1500
    L = [nsym // 2 - 3, nsym - nsym // 2 - 6][: len(alg)] if len(alg) == 2 else [nsym - 9]
    need_ref = any(a in ("nlms", "da-rde") for a in alg)
    p = Bag(nTaps=ntaps, SpS=sps, M=16, constType="qam", alg=alg, mu=mu, L=L, prgsBar=False, returnResults=True)
    y, H, err, _ = eq(x, p, ref) if need_ref else eq(x, p)
    yo, Ho, _, eo, _ = ro.mimo_adapt_equalizer(x, ref if need_ref else None, golden["const_qam16"], nTaps=ntaps, SpS=sps, alg=alg,
                                               mu=mu, L=L)
    n = sum(L)
    assert rel_l2(y[:n], yo[:n]) < 1e-4
    assert rel_l2(H, Ho) < 2e-4
    # identical hard decisions on the converged part
    c = golden["const_qam16"].astype(np.complex128)
    c = c / np.sqrt(np.mean(np.abs(c) ** 2))
    tail = slice(n // 2, n)
    d = np.argmin(np.abs(y[tail, :, None] - c), axis=-1)
    do = np.argmin(np.abs(yo[tail, :, None] - c), axis=-1)
    assert np.array_equal(d, do)
