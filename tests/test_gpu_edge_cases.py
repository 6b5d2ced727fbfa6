"""Edge cases of the drop-in boundary on the GPU: degenerate step/span counts, odd and non-power-of-two
lengths (cuFFT-driven engine), several pol-pairs, tiny windows, training sections shorter than the
signal, and full-size (2^20) properties that do not need the oracle."""
import numpy as np
import pytest

from conftest import Bag, rel_l2

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def api():
    from opticommpy_b200 import _cabi
    _cabi.require_cuda()
    from opticommpy_b200 import carrierRecovery, channels, equalization
    return Bag(ssfm=channels.ssfm, manakovSSF=channels.manakovSSF, manakovDBP=equalization.manakovDBP,
               edc=equalization.edc, eq=equalization.mimoAdaptEqualizer, bps=carrierRecovery.bps)


def rnd(seed, shape, scale=0.03):
    rng = np.random.default_rng(seed)
    return (rng.normal(size=shape) + 1j * rng.normal(size=shape)) * scale


def test_zero_spans_and_zero_steps(api):
    x = rnd(0, (1000, 2))
    out = api.manakovSSF(x, Bag(Fs=64e9, Ltotal=40, Lspan=80, amp="ideal", saveSpanN=[], prgsBar=False))
    assert np.allclose(out, x.astype(np.complex64), rtol=1e-6)  # floor(40/80) = 0 spans: field unchanged
    y = rnd(1, 777)
    out = api.ssfm(y, Bag(Fs=64e9, Ltotal=80, Lspan=80, hz=100.0, amp=None, prgsBar=False))
    assert np.allclose(out, y.astype(np.complex64), rtol=1e-6)  # floor(80/100) = 0 steps per span


@pytest.mark.parametrize("n", [999, 1000, 3 * 5 * 7 * 11])
def test_odd_and_composite_lengths_vs_oracle(api, n):
    from oracle import fiber_oracle as fo
    x = rnd(n, (n, 2))
    kw = dict(Fs=64e9, Ltotal=40, Lspan=20, hz=5.0, amp="ideal", nlprMethod=False)
    ref = fo.manakov(x, fo.FiberConfig(**kw))
    out = api.manakovSSF(x, Bag(saveSpanN=[], prgsBar=False, **kw))
    assert rel_l2(out, ref) < 2e-5
    y = x[:, 0]
    ref = fo.nlse_ssfm(y, fo.FiberConfig(**kw))
    assert rel_l2(api.ssfm(y, Bag(prgsBar=False, **kw)), ref) < 2e-5


def test_three_pol_pairs_global_semantics(api):
    """K = 3 pairs in one call: the step size and the convergence test are global over the pairs,
    like the reference (channels.py:394, 517)."""
    from oracle import fiber_oracle as fo
    x = rnd(5, (2048, 6))
    x[:, 2:4] *= 2.0  # one strong pair drives the adaptive step of all
    kw = dict(Fs=64e9, Ltotal=10, Lspan=10, hz=1.0, amp=None, nlprMethod=True, maxNlinPhaseRot=1e-2)
    st = {}
    ref = fo.manakov(x, fo.FiberConfig(**kw), stats=st)
    p = Bag(saveSpanN=[], prgsBar=False, **kw)
    out = api.manakovSSF(x, p)
    assert rel_l2(out, ref) < 5e-5
    assert abs(p._b200_stats["steps"] - st["steps"]) <= 1


def test_full_size_properties_2e20(api):
    """cfg2 size (N = 2^20, 512 GSa/s): lossless fiber conserves power; DBP undoes SSF."""
    x = rnd(11, (1 << 20, 2), 0.04)
    X = np.fft.fft(x, axis=0)
    X[np.abs(np.fft.fftfreq(1 << 20)) > 0.4] = 0
    x = np.fft.ifft(X, axis=0)
    kw = dict(Fs=512e9, Ltotal=8, Lspan=4, hz=0.08, alpha=0.0, amp=None, nlprMethod=False, saveSpanN=[], prgsBar=False)
    p = Bag(**kw)
    y = api.manakovSSF(x, p)
    assert p._b200_stats["steps"] == 2 * 50
    assert np.sum(np.abs(y) ** 2) == pytest.approx(np.sum(np.abs(x) ** 2), rel=1e-4)
    assert rel_l2(y, x) > 1e-2  # the fiber did something
    back = api.manakovDBP(y, Bag(**kw))
    assert rel_l2(back, x) < 2e-4


def test_edc_small_cases(api):
    from oracle import rxdsp_oracle as ro
    s = rnd(3, 300, 1.0)
    out = api.edc(s, Bag(L=20, D=16, Fc=193.1e12, Fs=64e9, Rs=32e9))
    assert out.shape == s.shape and rel_l2(out, ro.edc(s, 20, 16, 193.1e12, 64e9, 32e9)) < 1e-5
    out = api.edc(s, Bag(L=20, Fs=64e9, NfilterCoeffs=7, Nfft=8))  # odd tap count, explicit sizes
    assert rel_l2(out, ro.edc(s, 20, 16, 193.1e12, 64e9, 32e9, NfilterCoeffs=7, Nfft=8)) < 1e-5


def test_equalizer_partial_training_and_odd_geometry(api, golden):
    from oracle import rxdsp_oracle as ro
    x = golden["eq_in"]
    # sum(L) < totalNumSymb: the tail of the output stays zero
    p = Bag(nTaps=15, SpS=2, M=16, alg=["cma"], mu=[2e-3], L=[1200], prgsBar=False)
    y = api.eq(x, p)
    assert y.shape == (3000, 2) and np.all(y[1200:] == 0) and np.all(y[:1200] != 0)
    # even tap count, 1 sample per symbol, single tap per lane overflow (33 taps -> 2 taps per lane)
    for ntaps, sps in [(8, 1), (33, 2), (1, 1)]:
        p = Bag(nTaps=ntaps, SpS=sps, M=16, alg=["cma", "rde"], mu=[1e-3, 1e-3], L=[500, 500], prgsBar=False)
        y = api.eq(x, p)
        yo, *_ = ro.mimo_adapt_equalizer(x, None, golden["const_qam16"], nTaps=ntaps, SpS=sps, alg=["cma", "rde"],
                                         mu=[1e-3, 1e-3], L=[500, 500])
        assert rel_l2(y[:1000], yo[:1000]) < 1e-4, (ntaps, sps)


def test_bps_degenerate_windows(api, golden):
    from oracle import rxdsp_oracle as ro
    r, c = golden["bps_in"][:300], golden["bps_const"]
    for N, B in [(0, 1), (0, 64), (200, 8), (3, 5)]:  # window longer than the signal; odd B
        ph, idx = api.bps(r, N, c, B, returnIndex=True)
        pho, idxo = ro.bps(r, N, c, B)
        assert np.array_equal(idx, idxo) and np.array_equal(ph, pho), (N, B)
