"""GPU parity of the receiver DSP chain (edc -> mimoAdaptEqualizer -> bps/cpr) vs the reference's
golden vectors and the CPU oracle.  Equalizer outputs: relative L2 <= 1e-4 (complex64 on both
sides, different summation order); BPS phase indices: bit-exact."""
import numpy as np
import pytest

from conftest import Bag, rel_l2
from test_oracle_golden import EQ_CASES

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def api():
    from opticommpy_b200 import _cabi
    _cabi.require_cuda()
    from opticommpy_b200 import carrierRecovery, equalization
    return Bag(edc=equalization.edc, eq=equalization.mimoAdaptEqualizer, eqb=equalization.mimoAdaptEqualizerBatch,
               bps=carrierRecovery.bps, cpr=carrierRecovery.cpr)


def test_edc_golden(api, golden):
    s = golden["edc_in"]
    out = api.edc(s, Bag(L=100, D=16, Fc=193.1e12, Fs=64e9, Rs=32e9))
    assert out.shape == s.shape and out.dtype == s.dtype
    assert rel_l2(out, golden["edc_100km"]) < 1e-5
    out = api.edc(s[:, 0], Bag(L=60, D=17, Fc=193.4e12, Fs=64e9, Rs=32e9, Nfft=256))
    assert out.shape == golden["edc_1d_nfft256"].shape
    assert rel_l2(out, golden["edc_1d_nfft256"]) < 1e-5
    out = api.edc(s.astype(np.complex64), Bag(L=100, D=16, Fc=193.1e12, Fs=64e9, Rs=32e9))
    assert out.dtype == np.complex64 and rel_l2(out, golden["edc_c64"]) < 1e-5
    with pytest.raises(NameError):  # FFT shorter than the filter (core.py:1009-1012)
        api.edc(s, Bag(L=100, Fs=64e9, Nfft=16))


def test_edc_at_scale_vs_oracle_and_inversion(api):
    """2^20 samples x 2 modes, 800 km (448 taps): (i) parity with the oracle's overlap-save (default
    NFFT=512, i.e. a different block size than the GPU's), (ii) the reference's own invariant
    edc(linearFiberChannel(x)) ~ x after re-alignment (tests/test_channels.py:106-151)."""
    from oracle import fiber_oracle as fo
    from oracle import rxdsp_oracle as ro
    rng = np.random.default_rng(0)
    n = 1 << 20
    X = np.fft.fft(rng.normal(size=(n, 2)) + 1j * rng.normal(size=(n, 2)), axis=0)
    X[np.abs(np.fft.fftfreq(n)) > 0.14] = 0  # 32 GBd at 4 SpS, like the reference's test signal
    x = np.fft.ifft(X, axis=0)
    Fs = 128e9
    y = fo.linear_fiber(x, 400, 0.0, 16, 193.1e12, Fs)
    out = api.edc(y, Bag(L=400, D=16, Fc=193.1e12, Fs=Fs, Rs=32e9))
    ref = ro.edc(y, 400, 16, 193.1e12, Fs, 32e9)
    assert rel_l2(out, ref) < 1e-5
    core = slice(4000, n - 4000)
    best = min(rel_l2(np.roll(out, lag, axis=0)[core], x[core]) for lag in range(-4, 5))
    assert best ** 2 < 0.02                                    # residual power < 2 %
    assert best ** 2 < rel_l2(y[core], x[core]) ** 2 / 100    # and < 1/100 of the uncompensated one


@pytest.mark.parametrize("tag", sorted(EQ_CASES))
def test_equalizer_golden(api, golden, tag):
    kw = EQ_CASES[tag]
    p = Bag(nTaps=15, SpS=2, M=16, constType="qam", prgsBar=False, returnResults=True, **kw)
    y, H, err, Hiter = api.eq(golden["eq_in"], p, golden["eq_ref"])
    assert y.dtype == np.complex64 and y.shape == golden[f"eq_{tag}_y"].shape
    assert rel_l2(y, golden[f"eq_{tag}_y"]) < 1e-4
    assert rel_l2(H, golden[f"eq_{tag}_H"]) < 1e-4
    assert err.dtype == np.complex64 and err.shape == golden[f"eq_{tag}_err"].shape
    if "static" not in kw["alg"]:
        assert rel_l2(err.real, golden[f"eq_{tag}_err"].real) < 1e-3
    assert Hiter.shape == golden[f"eq_{tag}_Hiter"].shape
    assert rel_l2(Hiter, golden[f"eq_{tag}_Hiter"]) < 1e-4
    # hard decisions on the equalised symbols are identical to the reference's
    c = golden["const_qam16"] / np.sqrt(np.mean(np.abs(golden["const_qam16"]) ** 2))
    dec = lambda z: np.argmin(np.abs(z[..., None] - c), axis=-1)
    tail = slice(1500, None)
    assert np.array_equal(dec(y[tail]), dec(golden[f"eq_{tag}_y"][tail]))


def test_equalizer_1d_and_errors(api, golden):
    p = Bag(nTaps=7, SpS=2, M=4, constType="qam", prgsBar=False, alg=["cma"], mu=[2e-3])
    y = api.eq(golden["eq_in"][:, 0], p)
    assert y.ndim == 1 and rel_l2(y, golden["eq_1d_y"]) < 1e-4
    with pytest.raises(ValueError):
        api.eq(golden["eq_in"], Bag(alg=["bogus"], mu=[1e-3], prgsBar=False))
    with pytest.raises(TypeError):
        api.eq(golden["eq_in"], Bag(alg="cma", mu=1e-3, prgsBar=False))


def test_equalizer_batch_equals_single(api, golden):
    """Independent streams in one launch give exactly the single-stream results."""
    rng = np.random.default_rng(2)
    xs = [golden["eq_in"] * np.exp(1j * rng.uniform(0, 2 * np.pi)) for _ in range(5)]
    p = Bag(nTaps=15, SpS=2, M=16, constType="qam", prgsBar=False, alg=["cma", "rde"], mu=[5e-3, 2e-3], L=[1000, 2000])
    ys = api.eqb(xs, p)
    for x, y in zip(xs, ys):
        assert np.array_equal(y, api.eq(x, p))


def test_equalizer_vs_oracle_wl_and_4modes(api, golden):
    from oracle import rxdsp_oracle as ro
    rng = np.random.default_rng(9)
    x = golden["eq_in"]
    x4 = np.concatenate([x, x[::-1] * np.exp(0.5j)], axis=1) + 0.01 * (rng.normal(size=(len(x), 4)))
    p = Bag(nTaps=9, SpS=2, M=16, constType="qam", prgsBar=False, alg=["cma"], mu=[1e-3], returnResults=True)
    y, H, err, _ = api.eq(x4, p)
    yo, Ho, _, eo, _ = ro.mimo_adapt_equalizer(x4, None, golden["const_qam16"], nTaps=9, SpS=2, alg=["cma"], mu=[1e-3])
    assert rel_l2(y, yo) < 1e-4 and rel_l2(H, Ho) < 1e-4
    # widely-linear mode, 45 taps (2 taps per lane)
    p = Bag(nTaps=45, SpS=2, M=16, constType="qam", prgsBar=False, alg=["nlms"], mu=[2e-3], runWL=True,
            returnResults=True)
    y, H, Hw, err, _ = api.eq(x, p, golden["eq_ref"])
    yo, Ho, Hwo, eo, _ = ro.mimo_adapt_equalizer(x, golden["eq_ref"], golden["const_qam16"], nTaps=45, SpS=2,
                                                 alg=["nlms"], mu=[2e-3], runWL=True)
    assert rel_l2(y, yo) < 1e-4 and rel_l2(H, Ho) < 1e-4 and rel_l2(Hw, Hwo) < 1e-3


def test_bps_bit_exact(api, golden):
    r, c = golden["bps_in"], golden["bps_const"]
    assert np.array_equal(api.bps(r, 12, c, 64), golden["bps_N12_B64"])
    assert np.array_equal(api.bps(r, 0, c, 16), golden["bps_N0_B16"])
    assert np.array_equal(api.bps(r[:500], 5, golden["const_psk8"], 32), golden["bps_N5_B32_psk"])


def test_bps_vs_oracle_large(api, golden):
    """2^17 symbols x 2 modes, window 25, B=64 (cfg3 geometry): indices identical to the oracle."""
    from oracle import rxdsp_oracle as ro
    rng = np.random.default_rng(4)
    c = golden["bps_const"]
    n = 1 << 17
    sym = c[rng.integers(0, 16, size=(n, 2))]
    pn = np.cumsum(rng.normal(scale=np.sqrt(2 * np.pi * 100e3 / 32e9), size=(n, 2)), axis=0)
    r = sym * np.exp(1j * pn) + 0.07 * (rng.normal(size=sym.shape) + 1j * rng.normal(size=sym.shape))
    ph, idx = api.bps(r, 12, c, 64, returnIndex=True)
    pho, idxo = ro.bps(r, 12, c, 64)
    assert np.array_equal(idx, idxo) and np.array_equal(ph, pho)


def test_cpr_golden(api, golden):
    out, ph = api.cpr(golden["bps_in"], Bag(alg="bps", M=16, constType="qam", N=25, B=64, runFOE=False,
                                            returnPhases=True))
    assert np.allclose(ph, golden["cpr_nofoe_ph"], atol=1e-12)
    assert rel_l2(out, golden["cpr_nofoe_out"]) < 1e-12
    out, ph = api.cpr(golden["cpr_foe_in"], Bag(alg="bps", M=16, constType="qam", N=35, B=64, runFOE=True,
                                                returnPhases=True, Ts=1 / 32e9))
    assert np.allclose(ph, golden["cpr_foe_ph"], atol=1e-12)
    assert rel_l2(out, golden["cpr_foe_out"]) < 1e-12
    out1 = api.cpr(golden["bps_in"][:, 0], Bag(alg="bpsGPU", M=16, N=25, runFOE=False))
    assert out1.ndim == 1


def test_cfg3_chain_vs_reference_golden(api):
    """cfg3 geometry (BASELINE configs[2]): edc(800 km) -> 2x2 mimoAdaptEqualizer(CMA -> RDE, nTaps = 31, the case the
    look-ahead kernel serves) -> cpr/bps(B = 64, window 25) on 2^17 symbols x 2 pol against the UNMODIFIED reference
    (tests/golden/make_golden_cfg3.py).  Every stage is fed the previous stage's own output, like a user's chain.
    Hard decisions after carrier recovery must be identical to the reference's."""
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    from cfg3_signal import make_signal
    with np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_cfg3.npz")) as z:
        g = {k: z[k] for k in z.files}
    with np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_vectors.npz")) as z:
        c = z["const_qam16"]
    c = c / np.sqrt(np.mean(np.abs(c) ** 2))
    nsym = int(g["nsym"])
    x, _ = make_signal(nsym, c, seed=int(g["seed"]))
    y1 = api.edc(x, Bag(L=800, D=16, Fc=193.1e12, Fs=64e9, Rs=32e9))
    assert rel_l2(y1[::8], g["edc_sub"]) < 1e-5
    p = Bag(nTaps=31, SpS=2, M=16, constType="qam", alg=["cma", "rde"], mu=list(g["mu"]),
            L=[int(0.2 * nsym), int(0.8 * nsym)], prgsBar=False, returnResults=True, prec=np.complex64)
    y2, H, err, _ = api.eq(y1, p)
    # 131072 adaptive updates in complex64 on both sides (numba fastmath vs CUDA, different summation order)
    assert rel_l2(H, g["eq_H"]) < 1e-3
    assert rel_l2(y2[::8], g["eq_y_sub"]) < 1e-3
    y3, ph = api.cpr(y2, Bag(alg="bps", M=16, constType="qam", N=25, B=64, runFOE=False, returnPhases=True))
    dec = lambda z: np.argmin(np.abs(z[..., None] - c), axis=-1).astype(np.uint8)
    d = dec(y3)
    mism = np.flatnonzero((d != g["cpr_dec"]).any(axis=1))
    # identical decisions except where the reference's own sample sits on a decision boundary (none expected)
    assert mism.size == 0, f"{mism.size} symbols decided differently, first at {mism[:5]}"
    assert np.mean(np.isclose(ph, g["cpr_ph"], rtol=0, atol=1e-9)) > 0.999  # a discrete grid: equal unless a metric tie flips
