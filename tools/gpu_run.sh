#!/bin/bash
# one gpurun call: probe, GPU tests, smoke, the default bench line; everything lands in gpurun_out/
TAG=${1:-r2}
mkdir -p gpurun_out
nvidia-smi -L
python tools/probe_mixed.py 2>&1 | tee gpurun_out/probe_${TAG}.log
timeout 1500 python -m pytest tests -m gpu -q --maxfail=40 -rf -p no:cacheprovider > gpurun_out/tests_${TAG}.log 2>&1
echo "pytest rc $?" >> gpurun_out/tests_${TAG}.log
grep -v "^$" gpurun_out/tests_${TAG}.log | tail -60 | cut -c1-220
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke_${TAG}.log 2>&1; echo "smoke rc $?"; tail -5 gpurun_out/smoke_${TAG}.log | cut -c1-300
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; echo "bench rc $?"
tail -c 1500 gpurun_out/bench_${TAG}.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_${TAG}.json").read().strip().splitlines()[-1])
    print("HEADLINE", d["config"]["workload"], d["value"], "img/s", d["ms_per_step"], "ms", "frac", d.get("whole_step_frac_of_peak"), "e2e", d["e2e"]["value"], "cpu", d["cpu_baseline"] and d["cpu_baseline"]["value"])
    for p in d["per_config"]:
        print("  ", p.get("workload"), p.get("value"), p.get("ms_per_step"), p.get("whole_step_frac_of_peak"), p.get("error"))
    sb = d.get("step_breakdown") or {}
    print("span", sb.get("span_ms_per_step"), "busy", sb.get("kernel_busy_ms_per_step"))
    for k in (sb.get("kernels") or [])[:24]:
        print("   %6.3f ms %5.1f%% n=%5.1f %s" % (k["ms_per_step"], 100*k["share"], k["launches_per_step"], k["kernel"]))
except Exception as e:
    print("bench parse failed", e)
PY
