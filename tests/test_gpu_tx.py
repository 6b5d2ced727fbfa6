"""Device-side WDM transmitter (opticommpy_b200.tx, csrc/tx.cu) against golden vectors of the unmodified reference
(tests/golden/ref_tx.npz) and against the CPU oracle at a larger size.  Tolerance: the reference computes in complex128,
the device path shapes the pulses in complex64 (stated: relative L2 <= 2e-6); the transmitted symbols are exact."""
import os

import numpy as np
import pytest

from conftest import ROOT, Bag, rel_l2

pytestmark = pytest.mark.gpu

CASES = {
    "dp5": dict(M=16, Rs=32e9, SpS=8, nBits=4 * 2048, pulseType="rrc", nFilterTaps=1024, pulseRollOff=0.01, powerPerChannel=-2.0,
                nChannels=5, wdmGridSpacing=37.5e9, nPolModes=2, seed=321),
    "sp3": dict(M=4, Rs=10e9, SpS=16, nBits=2 * 1024, pulseType="rrc", nFilterTaps=256, pulseRollOff=0.1,
                powerPerChannel=[-1.0, 0.0, 1.5], nChannels=3, wdmGridSpacing=25e9, nPolModes=1, seed=17, laserLinewidth=100e3),
    "dp4": dict(M=16, Rs=32e9, SpS=4, nBits=4 * 1024, pulseType="rrc", nFilterTaps=128, pulseRollOff=0.2, powerPerChannel=0.0,
                nChannels=4, wdmGridSpacing=40e9, nPolModes=2, seed=5, mzmScale=0.25),
}


@pytest.fixture(scope="module")
def gtx():
    with np.load(os.path.join(ROOT, "tests", "golden", "ref_tx.npz")) as z:
        return {k: z[k] for k in z.files}


@pytest.mark.parametrize("tag", sorted(CASES))
def test_simple_wdm_tx_vs_reference_golden(gtx, tag):
    from opticommpy_b200.tx import simpleWDMTx
    p = Bag(prgsBar=False, **CASES[tag])
    sig, symb, p2 = simpleWDMTx(p)
    ref = gtx[f"tx_{tag}_sig"]
    assert sig.shape == ref.shape and sig.dtype == np.complex128
    assert np.allclose(symb, gtx[f"tx_{tag}_symb"], rtol=0, atol=1e-15)          # the reference's symbols, exactly
    assert np.allclose(p2.wdmFreqGrid, gtx[f"tx_{tag}_grid"])
    assert rel_l2(sig, ref) < 2e-6
    # per-channel launch power: total = sum of the channel powers (pnorm + sqrt(Pch / nPol) scaling, tx.py:206)
    pw = CASES[tag]["powerPerChannel"]
    tot = np.sum(10 ** (np.asarray(pw if isinstance(pw, list) else [pw] * CASES[tag]["nChannels"], dtype=float) / 10) * 1e-3)
    assert abs(np.sum(np.mean(np.abs(sig) ** 2, axis=0)) / tot - 1) < 0.05      # channels are nearly orthogonal
    assert p2.pmf.shape == (CASES[tag]["M"],) and p is p2                       # defaults / results written back into param


def test_tx_rows_device_at_scale_vs_oracle_and_into_the_fiber():
    """11-channel DP-16QAM at N = 2^17 on the device vs the float64 oracle, then straight into manakov_rows_device."""
    import torch
    from opticommpy_b200.channels import manakov_rows_device
    from opticommpy_b200.tx import wdm_tx_rows_device
    from oracle import tx_oracle as to
    kw = dict(M=16, Rs=32e9, SpS=16, nBits=4 * 8192, nFilterTaps=1024, pulseRollOff=0.01, powerPerChannel=-2.0, nChannels=11,
              wdmGridSpacing=37.5e9, nPolModes=2, seed=123)
    rows, symb, p = wdm_tx_rows_device(Bag(prgsBar=False, pulseType="rrc", **kw))
    assert rows.shape == (2, 8192 * 16) and rows.dtype == torch.complex64 and rows.is_cuda
    sig_o, symb_o, grid_o = to.simple_wdm_tx(**kw)
    assert np.allclose(symb, symb_o, rtol=0, atol=1e-15)
    assert rel_l2(rows.cpu().numpy().T, sig_o) < 2e-6
    prm = Bag(Fs=32e9 * 16, Ltotal=80, Lspan=80, hz=4.0, alpha=0.2, D=16, gamma=1.3, Fc=193.1e12, amp="ideal", NF=4.5, maxIter=10,
              tol=1e-5, nlprMethod=False, maxNlinPhaseRot=2e-2, seed=None)
    p_in = float(torch.sum(torch.mean(torch.abs(rows) ** 2, dim=1)))
    st = manakov_rows_device(rows, prm, +1)
    p_out = float(torch.sum(torch.mean(torch.abs(rows) ** 2, dim=1)))
    assert st["steps"] == 20 and abs(p_out / p_in - 1) < 1e-3                    # ideal amplification restores the launch power


def test_tx_argument_errors():
    from opticommpy_b200.tx import simpleWDMTx
    with pytest.raises(ValueError):
        simpleWDMTx(Bag(prgsBar=False, probDist="gaussian", nBits=4096, seed=1))
    with pytest.raises(AssertionError):
        simpleWDMTx(Bag(prgsBar=False, powerPerChannel=[0.0, 1.0], nChannels=3, nBits=8192, seed=1))
    with pytest.raises(ValueError):
        simpleWDMTx(Bag(prgsBar=False, nBits=4 * 16, SpS=4, nFilterTaps=1024, seed=1))   # filter longer than the signal
