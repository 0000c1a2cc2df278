#!/bin/bash
# bandwidth kernels: CUDA-event timings + one ncu --set full capture
TAG=${1:-r2ew}
mkdir -p gpurun_out
timeout 300 python tools/ncu_elementwise.py --time > gpurun_out/ew_times_${TAG}.txt 2>&1; echo "times rc $?"; cat gpurun_out/ew_times_${TAG}.txt
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -f -o gpurun_out/prof_ew_${TAG} \
  python tools/ncu_elementwise.py > gpurun_out/prof_ew_${TAG}.log 2>&1; echo "ncu rc $?"
ncu -i gpurun_out/prof_ew_${TAG}.ncu-rep --page raw --csv > gpurun_out/prof_ew_${TAG}.csv 2>/dev/null
python tools/ncu_ew_summary.py gpurun_out/prof_ew_${TAG}.csv > gpurun_out/prof_ew_${TAG}.md; cat gpurun_out/prof_ew_${TAG}.md
# and the bench line of the current tree (clock sampler check)
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; echo "bench rc $?"
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/bench_${TAG}.json") if l.startswith("{")][-1])
print("HEADLINE", d["config"]["workload"], d["value"], "img/s", d["ms_per_step"], "ms", d["clocks"])
for p in d["per_config"]:
    print("  ", p.get("workload"), p.get("value"), p.get("ms_per_step"), p.get("error"))
PY
