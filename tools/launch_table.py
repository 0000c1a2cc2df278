"""ncu launch list (gpu__time_duration.sum per launch, --csv) -> per-kernel share table of ONE training step.
    python tools/launch_table.py gpurun_out/launches_final.csv > profiles/<name>.md
One step = the launches between two consecutive launches of the G optimizer's Adam kernel (the last kernel of a step)."""
import collections
import csv
import re
import sys

path = sys.argv[1]
rows = []
with open(path) as fh:
    lines = [l for l in fh if not l.startswith("==")]
rd = csv.reader(lines)
hdr = next(rd)
ik, iv, im = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
for r in rd:
    if len(r) <= iv or r[im] != "gpu__time_duration.sum":
        continue
    rows.append((r[ik], float(r[iv].replace(",", "")) / 1e3))          # ns -> us
adam = [i for i, (n, _) in enumerate(rows) if n.startswith("adam_kernel") or "adam_kernel" in n]
# a step has two Adam launches (D then G): take the span between the 2nd and the 4th Adam launch seen
if len(adam) >= 4:
    step = rows[adam[1] + 1: adam[3] + 1]
else:
    step = rows
agg = collections.defaultdict(lambda: [0, 0.0])
for n, us in step:
    n = re.sub(r"^void |pg::|\(.*", "", n)[:70]
    agg[n][0] += 1
    agg[n][1] += us
tot = sum(v[1] for v in agg.values())
print(f"Launches in the step: {len(step)}; sum of kernel durations: {tot/1e3:.3f} ms.\n")
print("| share | ms | launches | avg us | kernel |\n|---|---|---|---|---|")
for n, (c, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| {us/tot*100:.1f}% | {us/1e3:.3f} | {c} | {us/c:.1f} | `{n}` |")
tc = sum(us for n, (c, us) in agg.items() if any(k in n for k in ("conv_gemm", "wgrad_gemm", "conv_rows", "wgrad_rows")))
print(f"\nGEMM-class (tcgen05) kernels: {tc/tot*100:.1f}% of the step's kernel time.")
