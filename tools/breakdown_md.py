"""profiles/r2_step_breakdown.md from bench.py JSON lines (their `step_breakdown` = CUPTI timeline of two CUDA-graph
replays, true warm kernel durations by kernel name).
    python tools/breakdown_md.py gpurun_out/bench_c*_r2i.json > profiles/r2_step_breakdown.md"""
import json
import sys

print("# r2 — kernel-time breakdown of one training step per BASELINE config (CUPTI timeline inside bench.py,\n"
      "# CUDA-graph replay, warm; fp16 mode, W step).  `span` = first kernel start to last kernel end per step.\n"
      "# (With PGGAN_PDL=1|2 a pre-launched kernel's CUPTI duration would include its wait for the predecessor; it is off.)\n"
      "# The two replays in the CUPTI window run at burst clocks; the\n"
      "# ms/step in each heading is the CUDA-event time over 10 back-to-back replays (power-capped clocks).\n")
for path in sys.argv[1:]:
    try:
        d = json.loads([ln for ln in open(path) if ln.startswith("{")][-1])
    except Exception as e:
        print(f"## {path}: unreadable ({e})\n")
        continue
    sb = d.get("step_breakdown") or {}
    cfg = d["config"]
    print(f"## {cfg['workload']}  —  {d['value']} img/s, {d['ms_per_step']} ms/step (CUDA events over {d['steps']} replays), "
          f"whole-step fraction of sustained bf16 peak {d.get('whole_step_frac_of_peak')}, SM clock {d['clocks']['sm_mhz']} MHz "
          f"{d['clocks']['reasons']}")
    if not sb or "kernels" not in sb:
        print(f"(no timeline: {sb})\n")
        continue
    print(f"span {sb['span_ms_per_step']} ms, kernel-busy {sb['kernel_busy_ms_per_step']} ms, {sb['launches_per_step']} launches per step"
          f" (programmatic dependent launch {'ON: durations overlap' if sb.get('programmatic_dependent_launch') else 'off: exclusive kernel times'})\n")
    print("```")
    for k in sb["kernels"]:
        print("%5.1f%% %8.3f ms/step n/step=%6.1f avg %8.1f us  %s" % (100 * k["share"], k["ms_per_step"], k["launches_per_step"],
                                                                    1e3 * k["ms_per_step"] / max(k["launches_per_step"], 1e-9), k["kernel"]))
    print("```")
    r = d.get("roofline") or {}
    if r:
        print(f"\nroofline line of this run: dominant kernel class `{r['kernel']}` ({r['share_of_step_time']:.0%} of the step): "
              f"{r['achieved']} {r['unit']} = {r['frac']:.3f} of {r['peak']} ({r['peak_source']}); "
              f"DRAM traffic per launch {r.get('traffic')} B vs {r.get('algorithmic_bytes_per_launch')} B algorithmic\n")
        print("| kernel class | ms/step (eager, CUDA events) | launches | TFLOP/s | algorithmic GB/s |")
        print("|---|---|---|---|---|")
        for k, v in r["kernel_classes"].items():
            print(f"| {k} | {v['ms_per_step']} | {v['launches_per_step']} | {v['tflops']} | {v['gbs']} |")
    print()
