"""Tuning sweep of the fused engine's kernel variants (never a bench number): for each environment
configuration run the same short cfg2-shaped propagation, check the output bit-for-bit against the
one-wave kernels and print whole-run time per SSFM step plus the in-situ per-kernel averages."""
import argparse
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import bench
from opticommpy_b200 import _cabi, _engine
from opticommpy_b200.channels import manakov_rows_device

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=300)
ap.add_argument("--n", type=int, default=1 << 20)
ap.add_argument("--configs", type=str, default="")
a = ap.parse_args()

KEYS = ["OCB_TPIPE", "OCB_TPIPE_SH", "OCB_TPIPE_GRID_ITER", "OCB_TPIPE_GRID_FIRST", "OCB_TPIPE_GRID_FWD",
        "OCB_FPIPE", "OCB_FPIPE_GRID", "OCB_PREDICT"]
default_configs = [
    {"OCB_TPIPE": "0", "OCB_FPIPE": "0", "OCB_PREDICT": "0"},
    {"OCB_TPIPE": "1", "OCB_FPIPE": "0", "OCB_PREDICT": "0"},
]
configs = json.loads(a.configs) if a.configs else default_configs

lib = _cabi.lib()
x = bench.synth_waveform(1, a.n)
rows0 = torch.from_numpy(np.ascontiguousarray(x.T.astype(np.complex64))).cuda()
prm = bench.channel_param(1, Ltotal=0.08 * a.steps, Lspan=0.08 * a.steps, amp=None)
plan = _engine.get_plan(a.n, 2)
ref = None
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for cfg in configs:
    for k in KEYS:
        os.environ.pop(k, None)
    os.environ.update({k: str(v) for k, v in cfg.items()})
    try:
        for rep in range(2):  # second run is the timed one
            r = rows0.clone()
            torch.cuda.synchronize()
            ev0.record()
            st = manakov_rows_device(r, prm)
            ev1.record()
            torch.cuda.synchronize()
        ms = ev0.elapsed_time(ev1)
        prof = (C.c_double * 6)()
        _cabi.check(lib.ocb_ssfm_plan_profile(plan.handle, 1), "profile on")
        r2 = rows0.clone()
        manakov_rows_device(r2, bench.channel_param(1, Ltotal=0.08 * 60, Lspan=0.08 * 60, amp=None))
        _cabi.check(lib.ocb_ssfm_plan_profile_read(plan.handle, prof), "profile read")
        _cabi.check(lib.ocb_ssfm_plan_profile(plan.handle, 0), "profile off")
        out = torch.view_as_real(r).cpu().numpy()
        if ref is None:
            ref, ref_st = out, st
        same = bool(np.array_equal(out, ref))
        rel = float(np.linalg.norm(out - ref) / np.linalg.norm(ref))
        print(json.dumps({"cfg": cfg, "us_per_step": 1e3 * ms / st["steps"], "Msamples_s": a.n * st["steps"] / (ms * 1e-3) / 1e6,
                          "steps": st["steps"], "iters": st["iterations"], "iters_equal": st["iterations"] == ref_st["iterations"],
                          "bit_equal": same, "rel_l2_vs_first": rel,
                          "iter_us": 1e3 * prof[0] / max(prof[1], 1), "first_us": 1e3 * prof[2] / max(prof[3], 1),
                          "lin_us": 1e3 * prof[4] / max(prof[5], 1)}), flush=True)
    except Exception as e:  # keep sweeping
        print(json.dumps({"cfg": cfg, "error": repr(e)}), flush=True)
