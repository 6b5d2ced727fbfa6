"""Stage-by-stage trace of the WDM link of examples/test_WDM_transmission.ipynb (reduced size) through either the unmodified
reference (--impl reference: CPU, build container) or this package (--impl b200: GPU), and a comparison of two traces.

    python tools/link_trace.py --impl reference --out scratch_ref.npz     # here (CPU)
    python tools/link_trace.py --impl b200 --compare scratch_ref.npz      # on the B200
    python tools/link_trace.py --impl reference --golden tests/golden/ref_link.npz   # compact fixture of the default size
                                                                                     # (tests/test_gpu_link.py)
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np


def load_api(impl):
    if impl == "reference":
        from unittest.mock import MagicMock
        os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/nbcache")
        for m in ["matplotlib", "matplotlib.pyplot", "matplotlib.mlab", "matplotlib.cm", "matplotlib.colors", "matplotlib.animation",
                  "mpl_scatter_density", "simple_pid", "prettytable", "tqdm", "tqdm.notebook"]:
            sys.modules[m] = MagicMock()
        sys.modules["tqdm.notebook"].tqdm = lambda it, **kw: it
        for cand in (os.path.join(ROOT, "baseline", "_ref"), "/root/reference"):
            if os.path.isdir(os.path.join(cand, "optic")):
                sys.path.insert(0, cand)
                break
        from optic.comm.metrics import fastBERcalc
        from optic.dsp.carrierRecovery import cpr
        from optic.dsp.core import decimate, firFilter, pnorm, pulseShape, symbolSync
        from optic.dsp.equalization import edc, mimoAdaptEqualizer
        from optic.models.channels import manakovSSF
        from optic.models.devices import basicLaserModel, pdmCoherentReceiver
        from optic.models.tx import simpleWDMTx
        from optic.utils import parameters
    else:
        from opticommpy_b200.carrierRecovery import cpr
        from opticommpy_b200.channels import manakovSSF
        from opticommpy_b200.core import decimate, firFilter, pnorm, symbolSync
        from opticommpy_b200.devices import basicLaserModel, pdmCoherentReceiver
        from opticommpy_b200.equalization import edc, mimoAdaptEqualizer
        from opticommpy_b200.metrics import fastBERcalc
        from opticommpy_b200.tx import pulseShape, simpleWDMTx
        from opticommpy_b200.utils import parameters
    return locals()


def run(api, nsym, spans, nch):
    P = api["parameters"]
    out = {}
    tx = P()
    tx.M, tx.Rs, tx.SpS, tx.pulseType, tx.nFilterTaps, tx.pulseRollOff = 16, 32e9, 16, "rrc", 1024, 0.01
    tx.powerPerChannel, tx.nChannels, tx.Fc, tx.laserLinewidth, tx.wdmGridSpacing, tx.nPolModes = -2, nch, 193.1e12, 100e3, 37.5e9, 2
    tx.nBits, tx.seed, tx.prgsBar = 4 * nsym, 123, False
    sigTx, symbTx_, tx = api["simpleWDMTx"](tx)
    out["tx"] = sigTx
    ch = P()
    ch.Ltotal, ch.Lspan, ch.alpha, ch.D, ch.gamma, ch.Fc, ch.hz, ch.maxIter, ch.tol = 50 * spans, 50, 0.2, 16, 1.3, tx.Fc, 0.5, 5, 1e-5
    ch.nlprMethod, ch.maxNlinPhaseRot, ch.prgsBar, ch.Fs, ch.seed = True, 2e-2, False, tx.Rs * tx.SpS, 456
    sig = api["manakovSSF"](sigTx, ch)
    out["fiber"] = sig
    Fs = ch.Fs
    k = nch // 2
    symbTx = symbTx_[:, :, k]
    lo = P()
    lo.P, lo.lw, lo.RIN_var, lo.Ns, lo.Fs, lo.seed, lo.freqShift = 10, 100e3, 0, len(sig), Fs, 789, tx.wdmFreqGrid[k] - 128e6
    sigLO = api["basicLaserModel"](lo)
    fe = P()
    fe.Fs, fe.polRotation, fe.pdl, fe.polDelay = Fs, np.pi / 3, 0, 3 / tx.Rs
    pd = P()
    pd.B, pd.Fs, pd.ideal, pd.seed = tx.Rs, Fs, True, 1011
    rx = api["pdmCoherentReceiver"](sig, sigLO, fe, pd)
    out["frontend"] = rx
    ps = P()
    ps.SpS, ps.nFilterTaps, ps.rollOff, ps.pulseType = 16, 1024, 0.01, "rrc"
    rx = api["firFilter"](api["pulseShape"](ps), rx)
    dec = P()
    dec.SpSin, dec.SpSout = 16, 2
    rx = api["decimate"](rx, dec)
    out["decimated"] = rx
    e = P()
    e.L, e.D, e.Fc, e.Rs, e.Fs = ch.Ltotal, 16, tx.Fc, tx.Rs, 2 * tx.Rs
    rx = api["edc"](rx, e)
    out["edc"] = rx
    symbRx = api["symbolSync"](rx, symbTx, 2)
    x, d = api["pnorm"](rx), api["pnorm"](symbRx)
    out["ref_symbols"] = d
    q = P()
    q.nTaps, q.SpS, q.numIter, q.storeCoeff, q.M, q.shapingFactor = 35, 2, 2, False, 16, 0
    q.L, q.prgsBar, q.alg, q.mu = [int(0.2 * d.shape[0]), int(0.8 * d.shape[0])], False, ["da-rde", "rde"], [5e-3, 5e-4]
    y = api["mimoAdaptEqualizer"](x, q, d)
    out["equalized"] = y
    c = P()
    c.alg, c.M, c.constType, c.shapingFactor, c.N, c.B, c.returnPhases, c.Ts = "bps", 16, "qam", 0, 25, 64, False, 1 / tx.Rs
    y = api["cpr"](y, c)
    out["cpr"] = y
    disc = min(5000, d.shape[0] // 10)
    ind = np.arange(disc, d.shape[0] - disc)
    ber, ser, snr = api["fastBERcalc"](y[ind, :], d[ind, :], 16, "qam")
    out["ber"], out["ser"], out["snr"] = np.asarray(ber), np.asarray(ser), np.asarray(snr)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--symbols", type=int, default=8192)
    ap.add_argument("--spans", type=int, default=4)
    ap.add_argument("--channels", type=int, default=11)
    ap.add_argument("--out")
    ap.add_argument("--compare")
    ap.add_argument("--golden", help="write the compact test fixture: 2-SpS stages, symbols, metrics and every 16th fiber sample")
    a = ap.parse_args()
    out = run(load_api(a.impl), a.symbols, a.spans, a.channels)
    print({k: (v.tolist() if v.size <= 2 else v.shape) for k, v in out.items()})
    if a.out:
        np.savez_compressed(a.out, **{k: (v.astype(np.complex64) if np.iscomplexobj(v) else v) for k, v in out.items()})
    if a.golden:
        keep = {k: out[k] for k in ("decimated", "edc", "ref_symbols", "equalized", "cpr", "ber", "ser", "snr")}
        keep["fiber_every16"] = out["fiber"][::16]
        keep["tx_every16"] = out["tx"][::16]
        keep["geometry"] = np.array([a.symbols, a.spans, a.channels])
        np.savez_compressed(a.golden, **{k: (v.astype(np.complex64) if np.iscomplexobj(v) else v) for k, v in keep.items()})
    if a.compare:
        with np.load(a.compare) as z:
            for k in z.files:
                r, g = z[k], out[k]
                if r.size <= 2:
                    print(f"{k:12s} reference {r} this {g}")
                else:
                    n = min(len(r), len(g))
                    print(f"{k:12s} rel L2 {np.linalg.norm(g[:n] - r[:n]) / np.linalg.norm(r[:n]):.3e}")


if __name__ == "__main__":
    main()
