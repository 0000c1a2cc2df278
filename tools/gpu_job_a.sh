#!/bin/bash
# tests of the double-backward mask hand-shake / fused pool-backward column sums + N-tile sweep of the small trunk layers
TAG=${1:-r2a}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_nets.py -q -x -p no:cacheprovider -k "elementwise or double_backward or gradient_penalty or full_size or train_step or graph_replay" -s 2>&1 | grep -v "^$" > gpurun_out/tests_${TAG}.log
echo "pytest rc ${PIPESTATUS[0]}"; tail -5 gpurun_out/tests_${TAG}.log | cut -c1-220; grep -n "mask passes" gpurun_out/tests_${TAG}.log
for bn in 0 16 32 64 128 256; do echo "== BN $bn"; PGGAN_FORCE_BN=$bn timeout 200 python tools/gemm_micro.py 10 small 2>&1 | grep conv_fprop; done > gpurun_out/sweep_small_${TAG}.txt 2>&1
cat gpurun_out/sweep_small_${TAG}.txt
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-per-config > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; echo "bench rc $?"
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/bench_${TAG}.json") if l.startswith("{")][-1])
print("HEADLINE", d["config"]["workload"], d["value"], "img/s", d["ms_per_step"], "ms", d["clocks"], "launches", d["gpu_launches"])
PY
