#!/bin/bash
TAG=${1:-r2}
for wl in c1_4x4_b16 c2_32x32_b64_fade c3_128x128_b32 c5_1024x1024_b4 c4_256x256_b16; do
  timeout 600 python bench.py --workload $wl --steps 10 --warmup 3 --no-per-config --no-cpu-baseline > gpurun_out/bench_${wl}_${TAG}.json 2> gpurun_out/bench_${wl}_${TAG}.err
  echo "$wl rc $?"
done
