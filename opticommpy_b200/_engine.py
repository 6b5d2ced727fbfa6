"""Plan/workspace cache shared by the host-side mirrors.

A plan (cuFFT handles + workspace partition + convergence mailbox) is created once per
(device, stream, N, rows) and kept in a small LRU; its device memory is a torch uint8 tensor so that
all HBM use goes through one allocator.  The stream is part of the key because a plan owns ONE
workspace: host threads that propagate independent waveforms concurrently, each on its own CUDA
stream (``sharding.run_concurrent``), get one plan each.  Nothing here computes: every numeric
operation is a C-ABI call.
"""
from __future__ import annotations

import collections
import ctypes as C
import os
import threading

import numpy as np

from . import _cabi

_MAX_PLANS = 24  # far above the number of plans in use at once (<= 8 worker streams x a few geometries): only idle plans reach the LRU tail
_lock = threading.Lock()
ENGINES = {"auto": 0, "cufft": 1, "fused": 2}
_default_engine = os.environ.get("OCB_ENGINE", "auto")


def set_default_engine(name: str) -> None:
    """'auto' (fused four-step kernels when N = 2^16..2^20 and K = 1, else cuFFT-driven),
    'cufft' or 'fused'.  Both engines implement the same reference loop; this is a tuning /
    cross-check switch, not a backend dispatch."""
    global _default_engine
    if name not in ENGINES:
        raise ValueError(f"unknown engine {name!r}")
    _default_engine = name
    clear_plans()


class SsfmPlan:
    def __init__(self, N: int, rows: int, device: int):
        torch = _cabi.require_cuda()
        self.N, self.rows, self.device = int(N), int(rows), int(device)
        self._lib = _cabi.lib()
        h = C.c_void_p()
        with torch.cuda.device(self.device):
            _cabi.check(self._lib.ocb_ssfm_plan_create(self.N, self.rows, C.byref(h)), "ocb_ssfm_plan_create")
            self.handle = h
            nbytes = int(self._lib.ocb_ssfm_plan_workspace_bytes(h))
            self.workspace = torch.empty(nbytes + 256, dtype=torch.uint8, device=f"cuda:{self.device}")
            base = self.workspace.data_ptr()
            aligned = (base + 255) // 256 * 256
            _cabi.check(self._lib.ocb_ssfm_plan_bind_workspace(h, C.c_void_p(aligned), nbytes),
                        "ocb_ssfm_plan_bind_workspace")
            _cabi.check(self._lib.ocb_ssfm_plan_set_engine(h, ENGINES[_default_engine]), "ocb_ssfm_plan_set_engine")
            self.engine = {1: "cufft", 2: "fused"}[int(self._lib.ocb_ssfm_plan_engine(h))]

    def close(self):
        if getattr(self, "handle", None):
            self._lib.ocb_ssfm_plan_destroy(self.handle)
            self.handle = None
            self.workspace = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_plans: "collections.OrderedDict[tuple, SsfmPlan]" = collections.OrderedDict()


def get_plan(N: int, rows: int, device: int | None = None) -> SsfmPlan:
    """The plan of (device, current stream of that device, N, rows); created on first use."""
    torch = _cabi.require_cuda()
    if device is None:
        device = torch.cuda.current_device()
    stream = int(torch.cuda.current_stream(device).cuda_stream)
    key = (int(device), stream, int(N), int(rows))
    with _lock:
        plan = _plans.get(key)
        if plan is not None:
            _plans.move_to_end(key)
            return plan
        evicted = []
        while len(_plans) >= _MAX_PLANS:
            evicted.append(_plans.popitem(last=False)[1])
    for old in evicted:  # an evicted plan is idle: its stream's owner is the thread asking for a new one, or it is stale
        old.close()
    plan = SsfmPlan(N, rows, device)
    with _lock:
        _plans[key] = plan
    return plan


def clear_plans() -> None:
    with _lock:
        old = list(_plans.values())
        _plans.clear()
    for p in old:
        p.close()


def dtype_tag(dt) -> int:
    dt = np.dtype(dt)
    if dt == np.complex64:
        return _cabi.OCB_C64
    if dt == np.complex128:
        return _cabi.OCB_C128
    raise TypeError(f"unsupported sample dtype {dt}")


def as_host_complex(a: np.ndarray):
    """C-contiguous complex64/complex128 view or copy of ``a`` (other dtypes -> complex128)."""
    a = np.asarray(a)
    if a.dtype not in (np.complex64, np.complex128):
        a = a.astype(np.complex128)
    return np.ascontiguousarray(a)


def legacy_complex_noise(shape, var: float, seed: int) -> np.ndarray:
    """The exact noise realisation of optic.dsp.core.gaussianComplexNoise (core.py:758-763):
    ``np.random.seed(seed)`` then the real parts, then the imaginary parts, from the legacy
    MT19937 + polar-method Gaussian stream (numba's implementation is bit-identical to
    ``numpy.random.RandomState``)."""
    rs = np.random.RandomState(seed)
    s = np.sqrt(var / 2)
    re = rs.normal(0, s, shape)
    im = rs.normal(0, s, shape)
    return re + 1j * im
