#!/bin/bash
TAG=${1:-r2f}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail=20 -rf -p no:cacheprovider > gpurun_out/tests_${TAG}.log 2>&1
echo "pytest rc $?"; tail -8 gpurun_out/tests_${TAG}.log | cut -c1-220
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; echo "bench rc $?"; tail -3 gpurun_out/bench_${TAG}.err
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/bench_${TAG}.json") if l.startswith("{")][-1])
print("HEADLINE", d["config"]["workload"], d["value"], "img/s", d["ms_per_step"], "ms", d["clocks"], "launches", d["gpu_launches"], "frac", d["roofline"].get("whole_step_frac_of_peak"))
for p in d["per_config"]:
    print("  ", p.get("workload"), p.get("value"), p.get("ms_per_step"), p.get("error"))
PY
