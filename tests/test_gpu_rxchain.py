"""The device-resident receiver chain (rxChain) against the stand-alone mirrors fed with each other's outputs."""
import numpy as np
import pytest

from conftest import Bag, rel_l2

pytestmark = pytest.mark.gpu


def test_rxchain_equals_staged_calls(golden):
    import os
    import sys
    from opticommpy_b200 import _cabi
    _cabi.require_cuda()
    from opticommpy_b200.carrierRecovery import cpr
    from opticommpy_b200.equalization import edc, mimoAdaptEqualizer
    from opticommpy_b200.rxchain import rxChain
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    from cfg3_signal import make_signal
    c = golden["const_qam16"] / np.sqrt(np.mean(np.abs(golden["const_qam16"]) ** 2))
    nsym = 1 << 14
    x, _ = make_signal(nsym, c, seed=3)
    x = x.astype(np.complex64)   # complex64 in: the staged calls then hand complex64 from stage to stage as well
    pe = Bag(L=800, D=16, Fc=193.1e12, Fs=64e9, Rs=32e9)
    pq = lambda: Bag(nTaps=31, SpS=2, M=16, constType="qam", alg=["cma", "rde"], mu=[1e-3, 2e-4],
                     L=[int(0.2 * nsym), int(0.8 * nsym)], prgsBar=False, returnResults=True)
    pc = Bag(alg="bps", M=16, constType="qam", N=25, B=64, runFOE=False, returnPhases=True)
    y1 = edc(x, pe)
    y2, H, err, _ = mimoAdaptEqualizer(y1, pq())
    y3, ph = cpr(y2, pc)
    timing = {}
    out, Hc, errc, phc = rxChain(x, pe, pq(), pc, returnAll=True, timing=timing)
    assert np.array_equal(Hc, H)
    assert np.array_equal(out, y3) and np.array_equal(phc, ph)
    assert set(timing) >= {"edc", "equalizer", "cpr_bps"} and all(timing[k] > 0 for k in ("edc", "equalizer", "cpr_bps"))
    # complex128 input: the chain keeps complex64 between the stages like the stand-alone calls
    out2 = rxChain(x.astype(np.complex128), pe, pq(), pc)
    assert rel_l2(out2, y3) < 1e-6
