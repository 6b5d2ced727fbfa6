"""Compact per-launch table from `ncu -i <rep> --page raw --csv` output (stdin or file): duration, issue/pipe
utilisation, top stall reasons, DRAM bytes.  Usage: ncu -i x.ncu-rep --page raw --csv | python tools/ncu_summary.py"""
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1]) if len(sys.argv) > 1 else sys.stdin))
hdr = rows[0]
col = {h: i for i, h in enumerate(hdr)}
pick = [
    ("dur_us", "gpu__time_duration.sum"),
    ("regs", "launch__registers_per_thread"),
    ("grid", "launch__grid_size"),
    ("warps%", "sm__warps_active.avg.pct_of_peak_sustained_active"),
    ("issue%", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
    ("fma%", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"),
    ("lsu%", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"),
    ("inst", "smsp__inst_executed.sum"),
    ("dramR_MB", "dram__bytes_read.sum"),
    ("dramW_MB", "dram__bytes_write.sum"),
    ("l2hit%", "lts__t_sector_hit_rate.pct"),
    ("bankconf", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"),
]
stall = {h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]: i for h, i in col.items()
         if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")}
units = rows[1]
print(" | ".join(["kernel".ljust(34)] + [p[0] for p in pick] + ["top stalls (per issue)"]))
for r in rows[2:]:
    name = re.sub(r"^void ", "", r[col["Kernel Name"]])
    name = re.sub(r"\(.*", "", name)[:34].ljust(34)
    vals = []
    for label, h in pick:
        if h not in col:
            vals.append("-")
            continue
        v = r[col[h]]
        try:
            f = float(v.replace(",", ""))
            u = units[col[h]]
            if label.endswith("_MB") and u.lower().startswith("kbyte"):
                f /= 1e3
            if label.endswith("_MB") and u.lower() == "byte":
                f /= 1e6
            vals.append(f"{f:.4g}")
        except ValueError:
            vals.append(v[:8])
    st = sorted(((float(r[i]) if r[i] not in ("", "no data") else 0.0, k) for k, i in stall.items()), reverse=True)[:5]
    print(" | ".join([name] + vals + [", ".join(f"{k}={v:.2f}" for v, k in st if k != "selected")]))
