#!/bin/bash
# ncu launch list (duration + DRAM bytes) of exactly ONE warm eager step of a workload (bench.py PGGAN_PROFILE_ONE_STEP hook)
TAG=${1:-r2fin}; shift
mkdir -p gpurun_out
for wl in "$@"; do
  PGGAN_PROFILE_ONE_STEP=1 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
    --clock-control none --csv --log-file gpurun_out/ncu_step_${wl}_${TAG}.csv \
    python bench.py --workload $wl --steps 1 --warmup 1 --no-graph --no-per-config --no-cpu-baseline --no-timeline > gpurun_out/ncu_step_${wl}_${TAG}.log 2>&1
  echo "$wl ncu rc $? lines $(wc -l < gpurun_out/ncu_step_${wl}_${TAG}.csv)"
done
