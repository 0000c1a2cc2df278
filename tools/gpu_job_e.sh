#!/bin/bash
TAG=${1:-r2g}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -q -x -p no:cacheprovider -k "rgb or elementwise" 2>&1 | tail -4
timeout 300 python tools/ncu_elementwise.py --time > gpurun_out/ew_times_${TAG}.txt 2>&1; echo "times rc $?"; cat gpurun_out/ew_times_${TAG}.txt | cut -c1-120
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; echo "bench rc $?"
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/bench_${TAG}.json") if l.startswith("{")][-1])
print("HEADLINE", d["config"]["workload"], d["value"], "img/s", d["ms_per_step"], "ms", d["clocks"], "launches", d["gpu_launches"], "frac", d["roofline"].get("whole_step_frac_of_peak"))
for p in d["per_config"]:
    print("  ", p.get("workload"), p.get("value"), p.get("ms_per_step"), p.get("error"))
PY
