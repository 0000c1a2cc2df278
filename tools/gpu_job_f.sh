#!/bin/bash
# programmatic dependent launch: tests with it on, micro + bench with it on and off
TAG=${1:-r2h}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail=20 -rf -p no:cacheprovider > gpurun_out/tests_${TAG}.log 2>&1
echo "pytest (PDL on) rc $?"; tail -6 gpurun_out/tests_${TAG}.log | cut -c1-220
for pdl in 1 0; do
  echo "== PGGAN_PDL=$pdl"
  PGGAN_PDL=$pdl timeout 300 python tools/splitk_micro.py 2>&1 | cut -c1-140 | head -4
  PGGAN_PDL=$pdl timeout 300 python tools/ncu_elementwise.py --time 2>&1 | grep "256x256x64" | cut -c1-100
  PGGAN_PDL=$pdl timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${TAG}_pdl$pdl.json 2> gpurun_out/bench_${TAG}_pdl$pdl.err; echo "bench rc $?"
  python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/bench_${TAG}_pdl$pdl.json") if l.startswith("{")][-1])
print("HEADLINE", d["config"]["workload"], d["value"], "img/s", d["ms_per_step"], "ms", d["clocks"], "launches", d["gpu_launches"], "frac", d["roofline"].get("whole_step_frac_of_peak"))
for p in d["per_config"]:
    print("  ", p.get("workload"), p.get("value"), p.get("ms_per_step"), p.get("error"))
PY
done
