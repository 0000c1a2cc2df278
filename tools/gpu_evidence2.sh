#!/bin/bash
# round-2 evidence run (one GPU): full GPU tests + smoke + the bench line and the reference arm, CUPTI step breakdowns of every
# config, ncu launch lists with DRAM bytes for C4 and C5, ncu --set full of the tensor-core kernels.  Everything -> gpurun_out/.
TAG=${1:-r2fin}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail=20 -rf -p no:cacheprovider > gpurun_out/tests_${TAG}.log 2>&1
echo "pytest rc $?"; tail -3 gpurun_out/tests_${TAG}.log | cut -c1-200
timeout 900 python -m pytest tests/test_gpu_nets.py tests/test_gpu_golden.py tests/test_gpu_input.py -q -s -p no:cacheprovider 2>&1 | grep -v "^$" > gpurun_out/parity_${TAG}.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke_${TAG}.log 2>&1; echo "smoke rc $?"; tail -3 gpurun_out/smoke_${TAG}.log | cut -c1-200
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; echo "bench rc $?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_${TAG}.json 2> gpurun_out/bench_ref_${TAG}.err; echo "ref bench rc $?"
for wl in c1_4x4_b16 c2_32x32_b64_fade c3_128x128_b32 c4_256x256_b16 c5_1024x1024_b4; do
  timeout 600 python bench.py --workload $wl --steps 10 --warmup 3 --no-per-config --no-cpu-baseline > gpurun_out/bench_${wl}_${TAG}.json 2> gpurun_out/bench_${wl}_${TAG}.err
  echo "$wl rc $?"
done
B="python bench.py --steps 1 --warmup 1 --no-graph --no-per-config --no-cpu-baseline --no-timeline"
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
  --launch-count 6000 --csv --log-file gpurun_out/ncu_launches_c4_${TAG}.csv $B > gpurun_out/ncu_launches_c4_${TAG}.log 2>&1
echo "ncu C4 launch list rc $?"
timeout 1200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
  --launch-count 9000 --csv --log-file gpurun_out/ncu_launches_c5_${TAG}.csv $B --workload c5_1024x1024_b4 > gpurun_out/ncu_launches_c5_${TAG}.log 2>&1
echo "ncu C5 launch list rc $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'conv_gemm_kernel|wgrad_gemm_kernel|conv_rows_kernel|wgrad_rows_kernel' \
  --launch-skip 250 --launch-count 20 -f -o gpurun_out/ncu_gemm_full_${TAG} $B > gpurun_out/ncu_gemm_full_${TAG}.log 2>&1
echo "ncu full rc $?"
ncu -i gpurun_out/ncu_gemm_full_${TAG}.ncu-rep --page raw --csv > gpurun_out/ncu_gemm_full_${TAG}.csv 2>/dev/null
rm -f gpurun_out/ncu_gemm_full_${TAG}.ncu-rep
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/bench_${TAG}.json") if l.startswith("{")][-1])
print("HEADLINE", d["config"]["workload"], d["value"], "img/s", d["ms_per_step"], "ms", "frac", d["roofline"].get("whole_step_frac_of_peak"), "e2e", d["e2e"]["value"], "cpu", d["cpu_baseline"] and d["cpu_baseline"]["value"], d["clocks"], "launches", d["gpu_launches"])
for p in d["per_config"]:
    print("  ", p.get("workload"), p.get("value"), p.get("ms_per_step"), p.get("whole_step_frac_of_peak"), p.get("error"))
PY
