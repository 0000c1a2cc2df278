#!/bin/bash
# evidence run (one GPU): ncu launch list with DRAM bytes, ncu --set full of the dominant GEMM kernels, compute-sanitizer,
# the stock-PyTorch GPU baseline of the reference algorithm.  Everything lands in gpurun_out/.
TAG=${1:-r2}
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 1 --no-graph --no-per-config --no-cpu-baseline --no-timeline"
$B > gpurun_out/ev_bench_plain_${TAG}.json 2> gpurun_out/ev_bench_plain_${TAG}.err; echo "plain bench rc $?"
# (1) launch list: duration + DRAM traffic of every launch of ~1.3 eager steps after the warm-up (serialised, cold caches)
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
  --launch-skip 2400 --launch-count 1100 --csv --log-file gpurun_out/ncu_launches_${TAG}.csv $B > gpurun_out/ncu_launches_${TAG}.log 2>&1
echo "ncu launch list rc $?"
# (2) full counters of the two dominant tensor-core kernels (a few launches each)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'conv_gemm_kernel|wgrad_gemm_kernel|conv_rows_kernel' \
  --launch-skip 250 --launch-count 16 -f -o gpurun_out/ncu_gemm_full_${TAG} $B > gpurun_out/ncu_gemm_full_${TAG}.log 2>&1
echo "ncu full rc $?"
ncu -i gpurun_out/ncu_gemm_full_${TAG}.ncu-rep --page raw --csv > gpurun_out/ncu_gemm_full_${TAG}.csv 2>/dev/null
# (3) compute-sanitizer over kernel-level tests (memcheck on a representative subset, racecheck on a smaller one)
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_kernels.py -q -x -p no:cacheprovider \
  -k "(test_conv_fprop and dtype2 and not 1-dtype2) or test_elementwise or test_rgb_ops or test_conv_fused_upsample or test_mbstd or test_conv_mask_and_raw_out or test_fused_adam or test_conv_fused_pixelnorm" \
  > gpurun_out/sanitizer_memcheck_${TAG}.log 2>&1
echo "memcheck rc $?"; tail -5 gpurun_out/sanitizer_memcheck_${TAG}.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_kernels.py -q -x -p no:cacheprovider \
  -k "test_elementwise or test_rgb_ops or test_mbstd" > gpurun_out/sanitizer_racecheck_${TAG}.log 2>&1
echo "racecheck rc $?"; tail -5 gpurun_out/sanitizer_racecheck_${TAG}.log
# (4) the reference algorithm on stock PyTorch/cuDNN kernels on the same GPU
for wl in c2_32x32_b64_fade c3_128x128_b32 c4_256x256_b16 c5_1024x1024_b4; do
  timeout 600 python tests/torch_gpu_baseline.py $wl W >> gpurun_out/torch_gpu_baseline_${TAG}.jsonl 2>> gpurun_out/torch_gpu_baseline_${TAG}.err
done
cat gpurun_out/torch_gpu_baseline_${TAG}.jsonl
