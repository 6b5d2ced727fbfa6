"""Accumulated complex64 error of the fused engine against the CPU oracle as a function of the number of
SSFM steps (N = 2^16, cfg2 fiber: Fs = 512 GSa/s, hz = 0.08 km, 11 x -2 dBm), plus the per-pass kernel times
at N = 2^20.  `OCB_LIB=<path>` selects another build of the library (e.g. an OCB_TW_MODE=0/1 build made by
`make -C opticommpy_b200/csrc twmodes`) for an A/B of the double-single twiddles.  Prints one JSON line."""
import ctypes as C
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


class Bag:
    def __init__(self, **kw):
        self.__dict__.update(kw)


def main():
    import bench
    import torch
    from opticommpy_b200 import _cabi, _engine
    from opticommpy_b200.channels import manakovSSF
    from oracle import fiber_oracle as fo
    steps_list = [int(s) for s in (sys.argv[1] if len(sys.argv) > 1 else "10,100,1000").split(",")]
    n = 1 << 16
    x = bench.synth_waveform(6, n)
    _engine.set_default_engine("fused")
    res = {}
    for steps in steps_list:
        L = 0.08 * steps
        st = {}
        ref = fo.manakov(x, fo.FiberConfig(Fs=512e9, Ltotal=L, Lspan=L, hz=0.08, amp=None, nlprMethod=False), stats=st)
        p = Bag(Fs=512e9, Ltotal=L, Lspan=L, hz=0.08, alpha=0.2, D=16, gamma=1.3, Fc=193.1e12, amp=None,
                nlprMethod=False, maxIter=10, tol=1e-5, saveSpanN=[], prgsBar=False)
        out = manakovSSF(x, p)
        res[str(st["steps"])] = {"rel_l2": float(np.linalg.norm(out - ref) / np.linalg.norm(ref)),
                                 "iters_equal": p._b200_stats["iterations"] == st["iterations"]}
    plan = _engine.get_plan(1 << 20, 2)
    lib = _cabi.lib()
    st_ = C.c_void_p(_cabi.stream_ptr(torch))
    pass_us = {}
    for which, name in enumerate(["k_freq", "k_time_FIRST", "k_time_ITER", "k_time_ITERF", "k_time_FWD"]):
        us = C.c_double()
        _cabi.check(lib.ocb_ssfm_plan_pass_time(plan.handle, which, 200, C.byref(us), st_), "pass_time")
        pass_us[name] = round(us.value, 2)
    print(json.dumps({"lib": os.path.basename(_cabi.LIB_PATH), "drift": res, "pass_us": pass_us}))


if __name__ == "__main__":
    main()
