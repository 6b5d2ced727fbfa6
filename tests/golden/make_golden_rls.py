"""Generate tests/golden/ref_rls.npz ('rls' stages of mimoAdaptEqualizer) by running the
UNMODIFIED reference (/root/reference) on seeded inputs.  Build container only:

    NUMBA_CACHE_DIR=/tmp/nbcache PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_rls.py

'dd-rls' is deliberately absent: the reference initialises the inverse correlation matrix only for alg == 'rls'
(equalization.py:447-451), so its 'dd-rls' stage runs on an undefined matrix and cannot serve as a golden vector.
"""
import os
import sys
from unittest.mock import MagicMock

os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/nbcache")
sys.dont_write_bytecode = True
for _m in ["matplotlib", "matplotlib.pyplot", "matplotlib.mlab", "matplotlib.cm", "matplotlib.colors",
           "matplotlib.animation", "mpl_scatter_density", "simple_pid", "prettytable"]:
    sys.modules[_m] = MagicMock()
sys.path.insert(0, os.environ.get("OPTICOMMPY_REF", "/root/reference"))

import numpy as np  # noqa: E402

from optic.comm.modulation import grayMapping  # noqa: E402
from optic.dsp.equalization import mimoAdaptEqualizer  # noqa: E402
from optic.utils import parameters  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_rls.npz")
G = {}
rng = np.random.default_rng(31)
c = grayMapping(16, "qam")
c = c / np.sqrt(np.mean(np.abs(c) ** 2))
n = 1500
sym = c[rng.integers(0, 16, size=(n, 2))].astype(np.complex128)
up = np.zeros((2 * n, 2), dtype=complex)
up[0::2] = sym
X = np.fft.fft(up, axis=0)
X *= (np.abs(np.fft.fftfreq(2 * n)) < 0.3)[:, None]
x = np.fft.ifft(X, axis=0) * 2
th = 0.4
x = x @ np.array([[np.cos(th), -np.sin(th)], [np.sin(th), np.cos(th)]]).T
x += 0.03 * (rng.normal(size=x.shape) + 1j * rng.normal(size=x.shape))
G["in"], G["ref"] = x, sym


def run(tag, **kw):
    p = parameters()
    p.nTaps, p.SpS, p.M, p.constType, p.prgsBar, p.returnResults = 11, 2, 16, "qam", False, True
    for k, v in kw.items():
        setattr(p, k, v)
    y, H, err, Hiter = mimoAdaptEqualizer(x, p, sym)
    G[f"{tag}_y"], G[f"{tag}_H"], G[f"{tag}_err"], G[f"{tag}_Hiter"] = y, H, err, Hiter


run("rls", alg=["rls"], mu=[1e-3], L=[n], lambdaRLS=0.99)
run("nlms_rls", alg=["nlms", "rls"], mu=[5e-3, 1e-3], L=[500, 1000], lambdaRLS=0.995)
run("rls_store", alg=["rls"], mu=[1e-3], L=[400], lambdaRLS=0.98, storeCoeff=True)
run("rls35", alg=["nlms", "rls"], mu=[5e-3, 1e-3], L=[300, 1200], lambdaRLS=0.995, nTaps=35)  # the notebook's tap count (> 32)
np.savez_compressed(OUT, **G)
print({k: (v.shape, v.dtype) for k, v in G.items()})
