"""Generate tests/golden/ref_vectors.npz by running the UNMODIFIED reference (OptiCommPy v0.11.0,
/root/reference) on seeded inputs.  Run in the build container only:

    NUMBA_CACHE_DIR=/tmp/nbcache PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

The reference tree is imported read-only (matplotlib & co. are stubbed; numba's cache is pointed
at a scratch directory so nothing is written under /root/reference).  The resulting vectors pin
the CPU oracle (oracle/) and, through it and directly, the CUDA path.
"""
import os
import sys
from unittest.mock import MagicMock

os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/nbcache")
sys.dont_write_bytecode = True
for _m in ["matplotlib", "matplotlib.pyplot", "matplotlib.mlab", "matplotlib.cm", "matplotlib.colors",
           "matplotlib.animation", "mpl_scatter_density", "simple_pid", "prettytable"]:
    sys.modules[_m] = MagicMock()
REF = os.environ.get("OPTICOMMPY_REF", "/root/reference")
sys.path.insert(0, REF)

import numpy as np  # noqa: E402

import optic.models.channels as ch  # noqa: E402
from optic.comm.modulation import grayMapping  # noqa: E402
from optic.dsp.carrierRecovery import bps, cpr  # noqa: E402
from optic.dsp.core import gaussianComplexNoise, pnorm  # noqa: E402
from optic.dsp.equalization import edc, manakovDBP, mimoAdaptEqualizer  # noqa: E402
from optic.models.devices import edfa  # noqa: E402
from optic.utils import parameters  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_vectors.npz")
G = {}


def field(seed, n, cols, power_w):
    rng = np.random.default_rng(seed)
    x = rng.normal(size=(n, cols)) + 1j * rng.normal(size=(n, cols))
    # band-limit a little so the spectrum is not white up to Nyquist
    X = np.fft.fft(x, axis=0)
    f = np.fft.fftfreq(n)
    X[np.abs(f) > 0.35] = 0
    x = np.fft.ifft(X, axis=0)
    x *= np.sqrt(power_w / np.mean(np.sum(np.abs(x) ** 2, axis=1)) * (cols / 2 if cols > 1 else 1))
    return x


class Counter:
    """Counts executed steps / iterations of manakovSSF without touching the reference source."""

    def __init__(self):
        self.iters = 0
        self.iffts = 0
        self._cc, self._ifft = ch.convergenceCondition, ch.ifft

    def __enter__(self):
        def cc(*a):
            self.iters += 1
            return self._cc(*a)

        def ifft(*a, **k):
            self.iffts += 1
            return self._ifft(*a, **k)

        ch.convergenceCondition, ch.ifft = cc, ifft
        return self

    def __exit__(self, *exc):
        ch.convergenceCondition, ch.ifft = self._cc, self._ifft

    @property
    def steps(self):  # per step: 2 iffts for the first half step + 2 per iteration
        return (self.iffts - 2 * self.iters) // 2


def P(**kw):
    p = parameters()
    for k, v in kw.items():
        setattr(p, k, v)
    return p


# ---- 1. scalar NLSE ssfm ---------------------------------------------------------------------------
x = field(1, 1024, 1, 4e-3)[:, 0]
G["ssfm_in"] = x
G["ssfm_ideal"] = ch.ssfm(x, P(Fs=64e9, Ltotal=160, Lspan=80, hz=2.0, amp="ideal", prgsBar=False))
G["ssfm_none"] = ch.ssfm(x, P(Fs=64e9, Ltotal=80, Lspan=80, hz=0.5, amp=None, gamma=2.0, prgsBar=False))
G["ssfm_edfa_seed7"] = ch.ssfm(x, P(Fs=64e9, Ltotal=80, Lspan=80, hz=4.0, amp="edfa", seed=7, prgsBar=False))

# ---- 2. manakovSSF ------------------------------------------------------------------------------------
e1 = field(2, 1024, 2, 8e-3)
G["mk_in"] = e1
with Counter() as c:
    G["mk_fixed_ideal"] = ch.manakovSSF(e1, P(Fs=64e9, Ltotal=160, Lspan=80, hz=4.0, amp="ideal", nlprMethod=False,
                                              saveSpanN=[], prgsBar=False))
G["mk_fixed_ideal_counts"] = np.array([c.steps, c.iters])
with Counter() as c:
    G["mk_fixed_degenerate"] = ch.manakovSSF(e1, P(Fs=64e9, Ltotal=80, Lspan=80, hz=0.8, amp=None, nlprMethod=False,
                                                   saveSpanN=[], prgsBar=False))
G["mk_fixed_degenerate_counts"] = np.array([c.steps, c.iters])  # 101 steps
with Counter() as c:
    G["mk_adaptive_edfa"] = ch.manakovSSF(e1, P(Fs=64e9, Ltotal=40, Lspan=20, hz=0.5, amp="edfa", seed=11,
                                                nlprMethod=True, maxNlinPhaseRot=2e-2, maxIter=5, prgsBar=False))
G["mk_adaptive_edfa_counts"] = np.array([c.steps, c.iters])
G["mk_savespans"] = ch.manakovSSF(e1, P(Fs=64e9, Ltotal=240, Lspan=80, hz=8.0, amp="ideal", nlprMethod=False,
                                        saveSpanN=[1, 3], prgsBar=False))
e2 = field(3, 512, 4, 6e-3)
G["mk_in_k2"] = e2
G["mk_k2"] = ch.manakovSSF(e2, P(Fs=64e9, Ltotal=80, Lspan=80, hz=4.0, amp="ideal", nlprMethod=False,
                                 saveSpanN=[], prgsBar=False))
with Counter() as c:
    G["mk_k2_adaptive"] = ch.manakovSSF(e2, P(Fs=64e9, Ltotal=20, Lspan=20, hz=4.0, amp=None, nlprMethod=True,
                                              saveSpanN=[], prgsBar=False))
G["mk_k2_adaptive_counts"] = np.array([c.steps, c.iters])

# ---- 3. manakovDBP ------------------------------------------------------------------------------------
G["dbp_of_fixed_ideal"] = manakovDBP(G["mk_fixed_ideal"], P(Fs=64e9, Ltotal=160, Lspan=80, hz=4.0, amp="ideal",
                                                            nlprMethod=False, saveSpanN=[], prgsBar=False))
G["dbp_adaptive"] = manakovDBP(e1, P(Fs=64e9, Ltotal=40, Lspan=20, hz=1.0, amp="edfa", nlprMethod=True,
                                     maxNlinPhaseRot=1e-2, saveSpanN=[], prgsBar=False))

# ---- 4. EDFA + noise stream ---------------------------------------------------------------------------
G["noise_seed5"] = gaussianComplexNoise((2, 64), 3.0e-7, 5)
G["edfa_seed9"] = edfa(e1[:256, 0], P(G=16.0, NF=4.5, Fc=193.1e12, Fs=64e9, seed=9))

# ---- 5. EDC --------------------------------------------------------------------------------------------
s = field(4, 4096, 2, 1.0)
G["edc_in"] = s
G["edc_100km"] = edc(s, P(L=100, D=16, Fc=193.1e12, Fs=64e9, Rs=32e9))
G["edc_1d_nfft256"] = edc(s[:, 0], P(L=60, D=17, Fc=193.4e12, Fs=64e9, Rs=32e9, Nfft=256))
G["edc_c64"] = edc(s.astype(np.complex64), P(L=100, D=16, Fc=193.1e12, Fs=64e9, Rs=32e9))

# ---- 6. adaptive equalizer ------------------------------------------------------------------------------
rng = np.random.default_rng(5)
M = 16
const = grayMapping(M, "qam")
const_n = pnorm(const)
nsym = 3000
sym = const_n[rng.integers(0, M, size=(nsym, 2))]
up = np.zeros((2 * nsym, 2), dtype=complex)
up[0::2] = sym
hps = np.array([0.02, -0.08, 0.3, 0.9, 0.35, -0.1, 0.03])
shaped = np.stack([np.convolve(up[:, i], hps, mode="same") for i in range(2)], axis=1)
th = 0.4
rot = np.array([[np.cos(th), -np.sin(th) * np.exp(0.3j)], [np.sin(th) * np.exp(-0.3j), np.cos(th)]])
rx = shaped @ rot.T
rx += 0.03 * (rng.normal(size=rx.shape) + 1j * rng.normal(size=rx.shape))
rx = pnorm(rx)
G["eq_in"] = rx
G["eq_ref"] = sym


def run_eq(tag, **kw):
    p = P(nTaps=15, SpS=2, M=M, constType="qam", prgsBar=False, returnResults=True, **kw)
    y, H, err, Hit = mimoAdaptEqualizer(rx, p, sym)
    G[f"eq_{tag}_y"], G[f"eq_{tag}_H"], G[f"eq_{tag}_err"], G[f"eq_{tag}_Hiter"] = y, H, err, Hit


run_eq("cma_rde", alg=["cma", "rde"], mu=[5e-3, 2e-3], L=[1000, 2000], numIter=2)
run_eq("nlms_ddlms", alg=["nlms", "dd-lms"], mu=[5e-3, 1e-3], L=[800, 2200])
run_eq("darde_rde", alg=["da-rde", "rde"], mu=[5e-3, 2e-3], L=[600, 2400], numIter=3)
run_eq("cma_static_store", alg=["cma", "static"], mu=[5e-3, 0.0], L=[2500, 500], storeCoeff=True)
p = P(nTaps=7, SpS=2, M=4, constType="qam", prgsBar=False, alg=["cma"], mu=[2e-3])
G["eq_1d_y"] = mimoAdaptEqualizer(rx[:, 0], p)

# ---- 7. BPS / CPR ----------------------------------------------------------------------------------------
nsym = 2000
sym = const_n[rng.integers(0, M, size=(nsym, 2))]
pn = np.cumsum(rng.normal(scale=np.sqrt(2 * np.pi * 100e3 / 32e9), size=(nsym, 2)), axis=0)
r = sym * np.exp(1j * pn) + 0.05 * (rng.normal(size=sym.shape) + 1j * rng.normal(size=sym.shape))
G["bps_in"] = r
cn = grayMapping(M, "qam")
cn = cn / np.sqrt(np.mean(np.abs(cn) ** 2))
G["bps_const"] = cn
G["bps_N12_B64"] = bps(r, 12, cn, 64)
G["bps_N0_B16"] = bps(r, 0, cn, 16)
G["bps_N5_B32_psk"] = bps(r[:500], 5, grayMapping(8, "psk"), 32)
o, ph = cpr(r, P(alg="bps", M=M, constType="qam", N=25, B=64, runFOE=False, returnPhases=True))
G["cpr_nofoe_out"], G["cpr_nofoe_ph"] = o, ph
fo = 30e6
rfo = r * np.exp(1j * 2 * np.pi * fo * np.arange(nsym)[:, None] / 32e9)
o, ph = cpr(rfo, P(alg="bps", M=M, constType="qam", N=35, B=64, runFOE=True, returnPhases=True, Ts=1 / 32e9))
G["cpr_foe_in"], G["cpr_foe_out"], G["cpr_foe_ph"] = rfo, o, ph

# ---- 8. constellations -------------------------------------------------------------------------------------
for Mq in (4, 16, 64, 256):
    G[f"const_qam{Mq}"] = grayMapping(Mq, "qam")
for Mp in (4, 8, 16):
    G[f"const_psk{Mp}"] = grayMapping(Mp, "psk")
G["const_apsk16"] = grayMapping(16, "apsk")

np.savez_compressed(OUT, **G)
print(f"wrote {OUT}: {len(G)} arrays, {os.path.getsize(OUT) / 1e6:.2f} MB")
for k in sorted(G):
    if k.endswith("_counts"):
        print(k, G[k])
