"""The reference's flagship example, examples/test_WDM_transmission.ipynb (cells 10-33), with the imports swapped to this
package: 11-channel DP-16QAM transmitter -> 14 x 50 km Manakov SSFM (adaptive step) -> coherent receiver with a noisy LO,
polarisation rotation and delay -> matched filter -> decimation -> EDC -> symbol synchronisation -> 2x2 adaptive equalizer
(DA-RDE -> RDE, 35 taps, two passes) -> BPS carrier recovery -> BER / SER / SNR.  Same parameter objects, same calls,
numpy in / numpy out; every stage runs on the GPU.  `--symbols` scales the run (the notebook uses 1e5 per polarisation).

    python tools/example_wdm_link.py [--symbols 100000] [--spans 14]
"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from opticommpy_b200.carrierRecovery import cpr
from opticommpy_b200.channels import manakovSSF
from opticommpy_b200.core import decimate, firFilter, pnorm, symbolSync
from opticommpy_b200.devices import basicLaserModel, pdmCoherentReceiver
from opticommpy_b200.equalization import edc, mimoAdaptEqualizer
from opticommpy_b200.metrics import fastBERcalc
from opticommpy_b200.tx import pulseShape, simpleWDMTx
from opticommpy_b200.utils import parameters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--symbols", type=float, default=1e5)
    ap.add_argument("--spans", type=int, default=14)
    a = ap.parse_args()
    T = {}

    def timed(name, fn, *args):
        t0 = time.time()
        out = fn(*args)
        T[name] = time.time() - t0
        return out

    paramTx = parameters()
    paramTx.M, paramTx.Rs, paramTx.SpS = 16, 32e9, 16
    paramTx.pulseType, paramTx.nFilterTaps, paramTx.pulseRollOff = "rrc", 1024, 0.01
    paramTx.powerPerChannel, paramTx.nChannels, paramTx.Fc = -2, 11, 193.1e12
    paramTx.laserLinewidth, paramTx.wdmGridSpacing, paramTx.nPolModes = 100e3, 37.5e9, 2
    paramTx.nBits, paramTx.seed, paramTx.prgsBar = int(np.log2(paramTx.M) * a.symbols), 123, False
    sigWDM_Tx, symbTx_, paramTx = timed("transmitter", simpleWDMTx, paramTx)

    paramCh = parameters()
    paramCh.Ltotal, paramCh.Lspan, paramCh.alpha, paramCh.D, paramCh.gamma = 50 * a.spans, 50, 0.2, 16, 1.3
    paramCh.Fc, paramCh.hz, paramCh.maxIter, paramCh.tol = paramTx.Fc, 0.5, 5, 1e-5
    paramCh.nlprMethod, paramCh.maxNlinPhaseRot, paramCh.prgsBar = True, 2e-2, False
    paramCh.Fs, paramCh.seed = paramTx.Rs * paramTx.SpS, 456
    sigWDM = timed("manakovSSF", manakovSSF, sigWDM_Tx, paramCh)
    Fs = paramCh.Fs

    chIndex = int(np.floor(paramTx.nChannels / 2))
    freqGrid = paramTx.wdmFreqGrid
    symbTx = symbTx_[:, :, chIndex]
    paramLO = parameters()
    paramLO.P, paramLO.lw, paramLO.RIN_var, paramLO.Ns, paramLO.Fs, paramLO.seed = 10, 100e3, 0, len(sigWDM), Fs, 789
    paramLO.freqShift = freqGrid[chIndex] - 128e6
    sigLO = timed("laser", basicLaserModel, paramLO)
    paramFE = parameters()
    paramFE.Fs, paramFE.polRotation, paramFE.pdl, paramFE.polDelay = Fs, np.pi / 3, 0, 3 * 1 / paramTx.Rs
    paramPD = parameters()
    paramPD.B, paramPD.Fs, paramPD.ideal, paramPD.seed = paramTx.Rs, Fs, True, 1011
    sigRx = timed("front end", pdmCoherentReceiver, sigWDM, sigLO, paramFE, paramPD)

    paramPS = parameters()
    paramPS.SpS, paramPS.nFilterTaps, paramPS.rollOff, paramPS.pulseType = paramTx.SpS, paramTx.nFilterTaps, paramTx.pulseRollOff, paramTx.pulseType
    sigRx = timed("matched filter", firFilter, pulseShape(paramPS), sigRx)
    paramDec = parameters()
    paramDec.SpSin, paramDec.SpSout = paramTx.SpS, 2
    sigRx = timed("decimation", decimate, sigRx, paramDec)
    paramEDC = parameters()
    paramEDC.L, paramEDC.D, paramEDC.Fc, paramEDC.Rs, paramEDC.Fs = paramCh.Ltotal, paramCh.D, paramCh.Fc, paramTx.Rs, 2 * paramTx.Rs
    sigRx = timed("CD compensation", edc, sigRx, paramEDC)
    symbRx = timed("symbol sync", symbolSync, sigRx, symbTx, 2)
    x, d = pnorm(sigRx), pnorm(symbRx)

    paramEq = parameters()
    paramEq.nTaps, paramEq.SpS, paramEq.numIter, paramEq.storeCoeff, paramEq.M = 35, 2, 2, False, paramTx.M
    paramEq.shapingFactor, paramEq.L, paramEq.prgsBar = paramTx.shapingFactor, [int(0.2 * d.shape[0]), int(0.8 * d.shape[0])], False
    paramEq.alg, paramEq.mu = ["da-rde", "rde"], [5e-3, 5e-4]
    y_EQ = timed("adaptive equalization", mimoAdaptEqualizer, x, paramEq, d)

    paramCPR = parameters()
    paramCPR.alg, paramCPR.M, paramCPR.constType, paramCPR.shapingFactor = "bps", paramTx.M, paramTx.constType, paramTx.shapingFactor
    paramCPR.N, paramCPR.B, paramCPR.returnPhases, paramCPR.Ts = 25, 64, False, 1 / paramTx.Rs
    y_CPR = timed("carrier phase recovery", cpr, y_EQ, paramCPR)

    discard = min(5000, d.shape[0] // 10)
    ind = np.arange(discard, d.shape[0] - discard)
    BER, SER, SNR = fastBERcalc(y_CPR[ind, :], d[ind, :], paramTx.M, "qam", px=paramTx.pmf)
    st = getattr(paramCh, "_b200_stats", {})
    print(f"samples per polarisation: {len(sigWDM)}, SSFM steps {st.get('steps')}, iterations {st.get('iterations')}")
    print("      pol.X      pol.Y")
    print(" SER: %.2e,  %.2e" % (SER[0], SER[1]))
    print(" BER: %.2e,  %.2e" % (BER[0], BER[1]))
    print(" SNR: %.2f dB,  %.2f dB" % (SNR[0], SNR[1]))
    print("-" * 44)
    for k, v in T.items():
        print(f"| {k:<30} | {v:6.3f} s |")
    print("-" * 44)
    return BER, SER, SNR


if __name__ == "__main__":
    main()
