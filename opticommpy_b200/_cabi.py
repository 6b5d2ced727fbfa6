"""ctypes binding of ``libopticomm_b200.so`` (the C-ABI declared in ``include/opticomm_b200.h``).

The product path has no CPU fallback: if the shared library is missing or a call fails, an
exception is raised.  PyTorch is used only to own device memory and streams.
"""
from __future__ import annotations

import ctypes as C
import os
import re

_HERE = os.path.dirname(os.path.abspath(__file__))
# OCB_LIB: another build of the same library (A/B experiments, e.g. tools/drift_experiment.py)
LIB_PATH = os.environ.get("OCB_LIB") or os.path.join(_HERE, "libopticomm_b200.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "opticomm_b200.h")

OCB_C64, OCB_C128 = 0, 1
AMP_NONE, AMP_IDEAL, AMP_EDFA = 0, 1, 2
NOISE_INJECTED, NOISE_PHILOX = 0, 1
ALG_IDS = {"cma": 0, "rde": 1, "nlms": 2, "dd-lms": 3, "da-rde": 4, "static": 5}


class OcbError(RuntimeError):
    """A C-ABI call returned a non-zero status."""


class ManakovParams(C.Structure):
    _fields_ = [
        ("alpha_lin", C.c_double), ("beta2", C.c_double), ("gamma", C.c_double), ("Fs", C.c_double),
        ("Lspan", C.c_double), ("hz", C.c_double), ("maxNlinPhaseRot", C.c_double), ("tol", C.c_double),
        ("n_spans", C.c_int32), ("maxIter", C.c_int32), ("nlprMethod", C.c_int32), ("direction", C.c_int32),
        ("amp_mode", C.c_int32), ("noise_mode", C.c_int32),
        ("edfa_gain_lin", C.c_double), ("edfa_noise_var", C.c_double), ("seed", C.c_uint64),
        ("n_save", C.c_int32), ("reserved", C.c_int32),
    ]


class ManakovStats(C.Structure):
    _fields_ = [
        ("steps", C.c_int64), ("iterations", C.c_int64), ("nonconverged", C.c_int64),
        ("last_lim", C.c_double), ("z_last_step", C.c_double),
    ]


class NlseParams(C.Structure):
    _fields_ = [
        ("alpha_lin", C.c_double), ("beta2", C.c_double), ("gamma", C.c_double), ("Fs", C.c_double),
        ("hz", C.c_double),
        ("n_spans", C.c_int32), ("n_steps", C.c_int32), ("amp_mode", C.c_int32), ("noise_mode", C.c_int32),
        ("edfa_gain_lin", C.c_double), ("edfa_noise_var", C.c_double), ("seed", C.c_uint64),
    ]


class WdmTxParams(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("mzmScale", "Vpi", "VbI", "VbQ", "Vphi", "ERI", "ERQ")]


_vp, _i, _i64, _d, _f, _u64 = C.c_void_p, C.c_int, C.c_int64, C.c_double, C.c_float, C.c_uint64

# name -> (restype, argtypes); must list every symbol declared in include/opticomm_b200.h
SIGNATURES = {
    "ocb_abi_version": (_i, []),
    "ocb_last_error": (C.c_char_p, []),
    "ocb_launch_count": (_i64, []),
    "ocb_launch_count_reset": (None, []),
    "ocb_ssfm_plan_create": (_i, [_i64, _i, C.POINTER(_vp)]),
    "ocb_ssfm_plan_workspace_bytes": (_i64, [_vp]),
    "ocb_ssfm_plan_bind_workspace": (_i, [_vp, _vp, _i64]),
    "ocb_ssfm_plan_destroy": (_i, [_vp]),
    "ocb_ssfm_plan_set_engine": (_i, [_vp, _i]),
    "ocb_ssfm_plan_engine": (_i, [_vp]),
    "ocb_ssfm_plan_profile": (_i, [_vp, _i]),
    "ocb_ssfm_plan_profile_read": (_i, [_vp, C.POINTER(C.c_double)]),
    "ocb_mimo_eq_rls_workspace_bytes": (_i64, [_i, _i, _i64, _i]),
    "ocb_mimo_eq_rls_run": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i64, _i64, _i64, _i64, _i64, _i64, _i64, _i, _i, _i, _i,
                                C.c_float, _vp, _i, _vp, _i64, _vp]),
    "ocb_pnorm_run": (_i, [_vp, _i64, _vp, _i64, _vp]),
    "ocb_decimate_workspace_bytes": (_i64, [_i, _i]),
    "ocb_decimate_run": (_i, [_vp, _vp, _i64, _i, _i, _i, _vp, _vp, _i64, _vp]),
    "ocb_ssfm_plan_pass_time": (_i, [_vp, _i, _i, C.POINTER(C.c_double), _vp]),
    "ocb_pack_fields": (_i, [_vp, _i, _i64, _i, _i, _vp, _vp]),
    "ocb_unpack_fields": (_i, [_vp, _i64, _i, _i, _vp, _i, _vp]),
    "ocb_cast_complex": (_i, [_vp, _i, _vp, _i, _i64, _vp]),
    "ocb_manakov_run": (_i, [_vp, _vp, C.POINTER(ManakovParams), _vp, C.POINTER(C.c_int32), _vp,
                             C.POINTER(ManakovStats), _vp]),
    "ocb_manakov_run_host": (_i, [_vp, _vp, _i, _vp, _i, C.POINTER(ManakovParams), _vp,
                                  C.POINTER(C.c_int32), C.POINTER(ManakovStats), _vp]),
    "ocb_manakov_nl_pass": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _i, _d, _d, _i, _vp]),
    "ocb_nlse_run": (_i, [_vp, _vp, C.POINTER(NlseParams), _vp, _vp]),
    "ocb_nlse_run_host": (_i, [_vp, _vp, _i, _vp, _i, C.POINTER(NlseParams), _vp, _vp]),
    "ocb_edfa_apply": (_i, [_vp, _i, _i64, _d, _d, _i, _vp, _i, _u64, _u64, _vp]),
    "ocb_edc_workspace_bytes": (_i64, [_i64, _i, _i]),
    "ocb_edc_run": (_i, [_vp, _vp, _i64, _i, _vp, _i, _vp, _i64, _vp]),
    "ocb_mimo_eq_run": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i64, _i64, _i64, _i64, _i64, _i64, _i64,
                             _i, _i, _i, _i, _f, _vp, _i, _vp, _i, _f, _i, _vp]),
    "ocb_bps_run": (_i, [_vp, _i64, _i, _vp, _i, _i, _i, _vp, _vp, _vp]),
    "ocb_cpr_workspace_bytes": (_i64, [_i64, _i]),
    "ocb_cpr_bps_run": (_i, [_vp, _i, _i64, _i, _vp, _i, _i, _i, _i, _d, _i, _vp, _vp, C.POINTER(C.c_double), _vp,
                             _i64, _vp]),
    "ocb_pdm_frontend_run": (_i, [_vp, _vp, _vp, _i64, _d, _d, _d, _d, _d, _d, C.POINTER(C.c_double), _vp]),
    "ocb_freq_shift_run": (_i, [_vp, _i, _i64, _d, _d, _vp]),
    "ocb_sync_sequence_run": (_i, [_vp, _i, _i64, _i, _vp, _vp]),
    "ocb_xcorr_workspace_bytes": (_i64, [_i, _i64, _i, _i64]),
    "ocb_xcorr_peak_run": (_i, [_vp, _i, _i64, _vp, _i, _i64, C.POINTER(C.c_int64), C.POINTER(C.c_double), _vp, _i64, _vp]),
    "ocb_sync_apply_run": (_i, [_vp, _vp, _i64, _i, _vp, _vp, _vp, _vp, _vp]),
    "ocb_upsample_run": (_i, [_vp, _i, _i64, _i, _vp, _vp]),
    "ocb_wdm_tx_workspace_bytes": (_i64, [_i, _i]),
    "ocb_wdm_tx_combine_run": (_i, [_vp, _vp, _i, _i, _i64, C.POINTER(WdmTxParams), C.POINTER(C.c_double), C.POINTER(C.c_double),
                                   _d, _vp, _vp, _i64, _vp]),
    "ocb_min_euclid": (_i, [_vp, _i, _i64, _vp, _i, _vp, _vp, _vp]),
    "ocb_ber_workspace_bytes": (_i64, [_i]),
    "ocb_ber_count": (_i, [_vp, _vp, _i, _i64, _i, _vp, _i, _i, _d, C.POINTER(C.c_double), C.POINTER(C.c_double),
                           C.POINTER(C.c_double), C.POINTER(C.c_int64), _vp, _i64, _vp]),
}


def declared_symbols(header_path: str = HEADER_PATH):
    """Names of all functions declared in the public header (used by the symbol-export test)."""
    text = open(header_path).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ocb_[a-z0-9_]+)\s*\(", text)))


_lib = None


def lib() -> C.CDLL:
    """Load the shared library (once).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise OcbError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a).  There is no CPU fallback for the product path."
        )
    # cuFFT: prefer whatever is already mapped (torch ships one); else the toolkit copy via rpath.
    handle = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(handle, name)  # AttributeError if the .so does not export it
        fn.restype = res
        fn.argtypes = args
    if handle.ocb_abi_version() != 1:
        raise OcbError("libopticomm_b200.so ABI version mismatch")
    _lib = handle
    return _lib


def check(status: int, what: str = "") -> None:
    if status != 0:
        msg = lib().ocb_last_error()
        raise OcbError(f"{what}: {msg.decode() if msg else 'unknown error'}")


def require_cuda():
    """Return torch after verifying a CUDA device is usable; fail loudly otherwise."""
    import torch

    if not torch.cuda.is_available():
        raise OcbError("opticommpy_b200 needs a CUDA device (B200, sm_100a); no CPU fallback exists.")
    return torch


def stream_ptr(torch_mod) -> int:
    return int(torch_mod.cuda.current_stream().cuda_stream)


def launch_count() -> int:
    return int(lib().ocb_launch_count())


def launch_count_reset() -> None:
    lib().ocb_launch_count_reset()
