#!/bin/bash
TAG=${1:-r2end}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail=20 -rf -p no:cacheprovider > gpurun_out/tests_${TAG}.log 2>&1
echo "pytest rc $?"; tail -2 gpurun_out/tests_${TAG}.log | cut -c1-200
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke_${TAG}.log 2>&1; echo "smoke rc $?"; tail -2 gpurun_out/smoke_${TAG}.log | cut -c1-200
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; echo "bench rc $?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_${TAG}.json 2> gpurun_out/bench_ref_${TAG}.err; echo "ref bench rc $?"
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/bench_${TAG}.json") if l.startswith("{")][-1])
print("HEADLINE", d["config"]["workload"], d["value"], "img/s", d["ms_per_step"], "ms", "frac", d["roofline"].get("whole_step_frac_of_peak"), "e2e", d["e2e"]["value"], "cpu", d["cpu_baseline"] and d["cpu_baseline"]["value"], d["clocks"], "launches", d["gpu_launches"], "roofline", d["roofline"]["achieved"], d["roofline"]["frac"])
for p in d["per_config"]:
    print("  ", p.get("workload"), p.get("value"), p.get("ms_per_step"), p.get("whole_step_frac_of_peak"), p.get("error"))
PY
