"""In-situ duration of each pass kernel of the fused SSFM engine (back-to-back launches on L2-resident plan
buffers, programmatic dependent launch as in the step loop).  Tuning aid, never a bench number."""
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from opticommpy_b200 import _cabi, _engine

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
lib = _cabi.lib()
plan = _engine.get_plan(n, 2)
plan.workspace.zero_()
out = {}
for which, name in enumerate(["k_freq", "k_time FIRST", "k_time ITER", "k_time ITERF", "k_time FWD"]):
    us = C.c_double()
    _cabi.check(lib.ocb_ssfm_plan_pass_time(plan.handle, which, 200, C.byref(us), C.c_void_p(_cabi.stream_ptr(torch))), "pass_time")
    out[name] = round(us.value, 2)
print(json.dumps(out))
