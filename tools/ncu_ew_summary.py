"""Summarise an `ncu --set full` capture of the bandwidth kernels (ncu -i X.ncu-rep --page raw --csv):
    python tools/ncu_ew_summary.py gpurun_out/prof_ew.csv"""
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[0]
col = {h: i for i, h in enumerate(hdr)}
want = [
    ("dur us", r"^gpu__time_duration.sum$", 1.0),
    ("DRAM %", r"^gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed$", 1.0),
    ("rd MB", r"^dram__bytes_read.sum$", 1.0),
    ("wr MB", r"^dram__bytes_write.sum$", 1.0),
    ("issue %", r"^smsp__issue_active.avg.pct_of_peak_sustained_active$", 1.0),
    ("occ %", r"^sm__warps_active.avg.pct_of_peak_sustained_active$", 1.0),
    ("regs", r"^launch__registers_per_thread$", 1.0),
    ("grid", r"^launch__grid_size$", 1.0),
    ("L2 hit %", r"^lts__t_sector_hit_rate.pct$", 1.0),
    ("long_sb", r"^smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio$", 1.0),
    ("short_sb", r"^smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio$", 1.0),
    ("barrier", r"^smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio$", 1.0),
    ("lg_thr", r"^smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio$", 1.0),
    ("mio_thr", r"^smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio$", 1.0),
    ("math_thr", r"^smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio$", 1.0),
    ("wait", r"^smsp__average_warps_issue_stalled_wait_per_issue_active.ratio$", 1.0),
]
cols = []
for name, pat, sc in want:
    c = next((col[h] for h in hdr if re.search(pat, h)), None)
    cols.append((name, c, sc))
print("| kernel | " + " | ".join(n for n, _, _ in cols) + " |")
print("|---|" + "---|" * len(cols))
units = rows[1]
for r in rows[2:]:
    name = re.sub(r"^void ", "", r[col["Kernel Name"]])
    name = re.sub(r"\(.*", "", name).replace("pg::", "")[:48]
    vals = []
    for n, c, sc in cols:
        if c is None or r[c] in ("", "n/a"):
            vals.append("-")
            continue
        try:
            v = float(r[c].replace(",", ""))
            u = units[c]
            if n == "dur us" and u in ("ns", "nsecond"):
                v /= 1000.0
            if n in ("rd MB", "wr MB"):
                v = v / 1e6 if u in ("byte", "B") else (v / 1e3 if u.lower().startswith("k") else (v if u.lower().startswith("m") else v * 1e3))
            vals.append(f"{v:.1f}" if abs(v) < 1e4 else f"{v:.0f}")
        except ValueError:
            vals.append(r[c])
    print(f"| `{name}` | " + " | ".join(vals) + " |")
