"""Summarise an `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv` launch list of
bench.py (tools/gpu_evidence.sh): per kernel launches, time share and DRAM bytes per launch -> a markdown table on
stdout and profiles/r2_ncu_traffic.json (bench.py's `roofline.traffic` reads it).
    python tools/ncu_traffic.py gpurun_out/ncu_launches_r2.csv c4_256x256_b16 > profiles/r2_ncu_launches_c4.md"""
import collections
import csv
import json
import os
import re
import sys

path, workload = sys.argv[1], sys.argv[2]
rows = collections.OrderedDict()
with open(path) as fh:
    lines = [ln for ln in fh if not ln.startswith("==")]
for r in csv.DictReader(lines):
    d = rows.setdefault(r["ID"], {"name": r["Kernel Name"]})
    d[r["Metric Name"]] = float(r["Metric Value"].replace(",", ""))


def short(name):
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"\(.*", "", name).replace("pg::", "")
    name = re.sub(r"at::native::.*", "at::native (eager ATen)", name)
    return name[:70]


def klass(name):
    for k in ("conv_gemm_swapped", "conv_gemm", "conv_rows", "wgrad_gemm", "wgrad_rows"):
        if name.startswith(k + "_kernel"):
            return k
    return None


# one training step = the launches between the 2nd and the 4th Adam launch of the capture (a step ends with the G
# optimizer's Adam kernel; D's comes mid-step)
order = list(rows.values())
adam = [i for i, d in enumerate(order) if "adam_kernel" in d["name"]]
if len(adam) >= 4:
    order = order[adam[1] + 1: adam[3] + 1]
    rows = collections.OrderedDict((i, d) for i, d in enumerate(order))
agg = collections.OrderedDict()
tot = 0.0
for d in rows.values():
    n = short(d["name"])
    a = agg.setdefault(n, {"launches": 0, "ns": 0.0, "rd": 0.0, "wr": 0.0})
    a["launches"] += 1
    a["ns"] += d.get("gpu__time_duration.sum", 0.0)
    a["rd"] += d.get("dram__bytes_read.sum", 0.0)
    a["wr"] += d.get("dram__bytes_write.sum", 0.0)
    tot += d.get("gpu__time_duration.sum", 0.0)
print(f"# r2 — ncu launch list of the {workload} step (eager launches, serialised by the profiler, cold caches)\n")
print("Command: `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none "
      "--launch-skip S --launch-count N --csv python bench.py --workload W --steps 1 --warmup 1 --no-graph --no-per-config "
      "--no-cpu-baseline --no-timeline` (tools/gpu_evidence2.sh); the table covers ONE step (between two G-optimizer Adam launches).\n"
      "Per-launch times under ncu are cold-cache and serialised:\n"
      "the kernel's SHARE of the step is what must agree with the live CUPTI timeline (profiles/r2_step_breakdown.md), not the absolute.\n")
tot_bytes = sum(a["rd"] + a["wr"] for a in agg.values())
print(f"{len(rows)} launches, {tot/1e6:.2f} ms of kernel time, {tot_bytes/1e9:.2f} GB of DRAM traffic in the step "
      f"(= {tot_bytes/6549.1e9*1e3:.2f} ms at the measured 6549 GB/s)\n")
print("| kernel | launches | share of kernel time | avg us | DRAM read MB / launch | DRAM write MB / launch | achieved DRAM GB/s |")
print("|---|---|---|---|---|---|---|")
classes = collections.defaultdict(lambda: {"launches": 0, "bytes": 0.0, "ns": 0.0})
for n, a in sorted(agg.items(), key=lambda kv: -kv[1]["ns"]):
    L = a["launches"]
    print(f"| `{n}` | {L} | {100*a['ns']/tot:.1f} % | {a['ns']/L/1e3:.1f} | {a['rd']/L/1e6:.2f} | {a['wr']/L/1e6:.2f} | "
          f"{(a['rd']+a['wr'])/a['ns']:.0f} |")
    k = klass(n)
    if k:
        classes[k]["launches"] += L
        classes[k]["bytes"] += a["rd"] + a["wr"]
        classes[k]["ns"] += a["ns"]
out_path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "r2_ncu_traffic.json")
data = {}
if os.path.exists(out_path):
    with open(out_path) as fh:
        data = json.load(fh)
data[workload] = {k: round(v["bytes"] / v["launches"]) for k, v in classes.items()}
data.setdefault("_how", "dram__bytes_read.sum + dram__bytes_write.sum per launch, averaged over the launches of each kernel "
                "class in one ncu pass over an eager step (tools/ncu_traffic.py)")
with open(out_path, "w") as fh:
    json.dump(data, fh, indent=1)
print("\nDRAM bytes per launch by kernel class (-> profiles/r2_ncu_traffic.json):", data[workload])
