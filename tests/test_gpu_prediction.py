"""The predicted-last-iteration launch protocol of the fused step loop (TM_ITERF chaining the next step,
TM_ROT recovery after a wrong prediction) must not change anything observable: same step / iteration counts
and bit-identical fields as the plain protocol (OCB_PREDICT=0), forwards and backwards.  The backward (DBP)
case has a growing power profile inside each span, so the iteration count rises along the span and the
wrong-prediction path is exercised."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _run(direction, predict, n, hz, ltot, pdbm, maxIter=10):
    import torch

    from opticommpy_b200.channels import manakov_rows_device
    from opticommpy_b200.utils import parameters

    rng = np.random.default_rng(5)
    p_lin = 1e-3 * 10 ** (pdbm / 10)
    x = (rng.normal(size=(2, n)) + 1j * rng.normal(size=(2, n))) * np.sqrt(p_lin / 4)
    rows = torch.from_numpy(x.astype(np.complex64)).cuda()
    p = parameters()
    p.Fs, p.Ltotal, p.Lspan, p.hz = 64e9, ltot, ltot / 2, hz
    p.alpha, p.D, p.gamma, p.Fc = 0.2, 16, 1.3, 193.1e12
    p.amp, p.NF, p.maxIter, p.tol, p.nlprMethod, p.maxNlinPhaseRot = "ideal", 4.5, maxIter, 1e-5, False, 2e-2
    os.environ["OCB_PREDICT"] = "1" if predict else "0"
    try:
        st = manakov_rows_device(rows, p, direction)
    finally:
        os.environ.pop("OCB_PREDICT", None)
    torch.cuda.synchronize()
    return torch.view_as_real(rows).cpu().numpy(), st


@pytest.mark.parametrize("direction", [+1, -1])
@pytest.mark.parametrize("n,pdbm", [(1 << 16, 12.0), (1 << 16, -3.0), (1 << 18, -3.0)])
def test_prediction_is_invisible(direction, n, pdbm):
    ref, st0 = _run(direction, False, n, 1.0, 100.0, pdbm)
    out, st1 = _run(direction, True, n, 1.0, 100.0, pdbm)
    assert st0["steps"] == st1["steps"] == 100
    assert st0["iterations"] == st1["iterations"]
    assert st0["nonconverged"] == st1["nonconverged"]
    if pdbm < 0:
        # the iteration count changes inside each span (3 -> 2 forwards: early convergence; 2 -> 3 backwards:
        # wrong prediction and TM_ROT recovery), so both off-nominal paths of the protocol are exercised
        assert 2 * st0["steps"] < st0["iterations"] < 3 * st0["steps"]
    else:
        assert st0["iterations"] == 3 * st0["steps"]
    assert np.array_equal(out, ref)


@pytest.mark.parametrize("maxIter", [1, 2, 3])
def test_prediction_with_iteration_limit(maxIter):
    """Steps that hit maxIter without converging (channels.py:431-434): the predicted-last iteration is then the
    last allowed one and does not converge; counts and fields must still equal the plain protocol's."""
    ref, st0 = _run(+1, False, 1 << 16, 1.0, 100.0, 12.0, maxIter)
    out, st1 = _run(+1, True, 1 << 16, 1.0, 100.0, 12.0, maxIter)
    assert st0 == st1
    if maxIter < 3:
        assert st0["nonconverged"] == st0["steps"] == 100 and st0["iterations"] == maxIter * 100
    assert np.array_equal(out, ref)
