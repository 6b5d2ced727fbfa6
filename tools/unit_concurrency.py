"""Sweep of the units kept in flight per GPU (sharding.run_concurrent) on the two sharded workloads of bench_extras:
cfg5 Monte-Carlo seeds (N = 2^18) and cfg4 per-channel DBP (N = 2^17).  Usage: python tools/unit_concurrency.py [seeds]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench_extras as bx  # noqa: E402

seeds = int(sys.argv[1]) if len(sys.argv) > 1 else 16
res = {}
for nw in (1, 2, 4, 6, 8):
    os.environ["OCB_UNIT_WORKERS"] = str(nw)
    a = bx.extra_cfg5_mc(torch, None, 1, 0, n_seeds=seeds)
    b = bx.extra_cfg4_dbp(torch, None, 1, 0)
    res[nw] = {"cfg5_seconds_per_seed": a["seconds"] / seeds, "cfg5_Msamples_per_s": a["value"], "cfg5_ber": a["ber_mean"],
               "cfg5_snr_db": a["snr_db_mean"], "cfg4_seconds": b["seconds"], "cfg4_Msamples_per_s": b["value"]}
    print(nw, json.dumps(res[nw]), flush=True)
print(json.dumps(res))
