"""The reference's flagship WDM notebook chain (examples/test_WDM_transmission.ipynb, cells 10-33, reduced to 8192 symbols and
4 x 50 km) through this package's drop-in mirrors against the SAME chain run through the unmodified reference on the CPU
(tests/golden/ref_link.npz, made by `tools/link_trace.py --impl reference --golden ...`): transmitter, adaptive-step
Manakov fiber with seeded ASE, noisy LO, coherent front end with polarisation rotation and delay, matched filter,
decimation, EDC, symbolSync, DA-RDE -> RDE equalizer (35 taps, two passes), BPS, error counting.  The fiber stage is compared
at the adaptive-step tolerance (the step controller amplifies complex64 rounding), the receiver
metrics must agree: SNR within 0.05 dB, no symbol errors on either side."""
import os
import sys

import numpy as np
import pytest

from conftest import ROOT, rel_l2

pytestmark = pytest.mark.gpu


def test_notebook_chain_vs_reference_golden():
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import link_trace
    with np.load(os.path.join(ROOT, "tests", "golden", "ref_link.npz")) as z:
        g = {k: z[k] for k in z.files}
    nsym, spans, nch = (int(v) for v in g["geometry"])
    out = link_trace.run(link_trace.load_api("b200"), nsym, spans, nch)
    assert rel_l2(out["tx"][::16], g["tx_every16"]) < 2e-6                      # same symbols, complex64 pulse shaping
    assert rel_l2(out["fiber"][::16], g["fiber_every16"]) < 3e-3 * spans        # adaptive step: 3e-3 per span (measured 7.7e-3)
    assert rel_l2(out["ref_symbols"], g["ref_symbols"]) < 1e-6                  # symbolSync takes the same decisions
    assert out["decimated"].shape == g["decimated"].shape and out["cpr"].shape == g["cpr"].shape
    assert rel_l2(out["edc"], g["edc"]) < 3e-3 * spans
    assert np.all(np.abs(out["snr"] - g["snr"]) < 0.05), (out["snr"], g["snr"])
    assert np.array_equal(out["ber"], g["ber"]) and np.all(g["ber"] == 0) and np.all(out["ser"] == 0)
