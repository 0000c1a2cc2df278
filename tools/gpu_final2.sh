#!/bin/bash
# final single-GPU run of the round: full GPU tests, parity print-out, smoke, the bench line + the reference arm, and the
# per-config CUPTI breakdowns with PGGAN_PDL=0 (exclusive per-kernel times; see bench.py cupti_breakdown)
TAG=${1:-r2z}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail=20 -rf -p no:cacheprovider > gpurun_out/tests_${TAG}.log 2>&1
echo "pytest rc $?"; tail -3 gpurun_out/tests_${TAG}.log | cut -c1-200
timeout 900 python -m pytest tests/test_gpu_nets.py tests/test_gpu_golden.py tests/test_gpu_input.py -q -s -p no:cacheprovider 2>&1 | grep -v "^$" > gpurun_out/parity_${TAG}.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke_${TAG}.log 2>&1; echo "smoke rc $?"; tail -3 gpurun_out/smoke_${TAG}.log | cut -c1-200
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; echo "bench rc $?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_${TAG}.json 2> gpurun_out/bench_ref_${TAG}.err; echo "ref bench rc $?"
for wl in c1_4x4_b16 c2_32x32_b64_fade c3_128x128_b32 c4_256x256_b16 c5_1024x1024_b4; do
  PGGAN_PDL=0 timeout 600 python bench.py --workload $wl --steps 10 --warmup 3 --no-per-config --no-cpu-baseline > gpurun_out/bench_${wl}_${TAG}.json 2> gpurun_out/bench_${wl}_${TAG}.err
  echo "$wl (PDL off) rc $?"
done
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/bench_${TAG}.json") if l.startswith("{")][-1])
print("HEADLINE", d["config"]["workload"], d["value"], "img/s", d["ms_per_step"], "ms", "frac", d["roofline"].get("whole_step_frac_of_peak"), "e2e", d["e2e"]["value"], "cpu", d["cpu_baseline"] and d["cpu_baseline"]["value"], d["clocks"], "launches", d["gpu_launches"])
for p in d["per_config"]:
    print("  ", p.get("workload"), p.get("value"), p.get("ms_per_step"), p.get("whole_step_frac_of_peak"), p.get("error"))
for wl in ["c1_4x4_b16","c2_32x32_b64_fade","c3_128x128_b32","c4_256x256_b16","c5_1024x1024_b4"]:
    try:
        q=json.loads([l for l in open(f"gpurun_out/bench_{wl}_${TAG}.json") if l.startswith("{")][-1])
        print("  PDL off", wl, q["value"], q["ms_per_step"], q["clocks"]["sm_mhz"])
    except Exception as e:
        print("  PDL off", wl, "failed", e)
PY
