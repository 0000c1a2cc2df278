#!/bin/bash
# programmatic dependent launch: 0 off / 1 all kernels / 2 tensor-core kernels only, on C1, C4, C5 (one process each)
mkdir -p gpurun_out
for mode in 2 0 1 2 0; do
  for wl in c5_1024x1024_b4 c4_256x256_b16 c1_4x4_b16; do
    PGGAN_PDL=$mode timeout 300 python bench.py --workload $wl --steps 10 --warmup 3 --no-per-config --no-cpu-baseline --no-timeline 2>/dev/null | \
      python -c "import sys,json; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('PDL=$mode', d['config']['workload'], d['ms_per_step'], 'ms', d['clocks']['sm_mhz'], 'MHz')"
  done
done
