#!/bin/bash
# multi-GPU run: data-parallel correctness check + the bench line at N ranks
N=${1:-2}; TAG=${2:-r2}
mkdir -p gpurun_out
nvidia-smi -L | head -8
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 tools/dp_check.py > gpurun_out/dp_check_n${N}_${TAG}.log 2>&1
echo "dp_check rc $?"; grep "dp_check" gpurun_out/dp_check_n${N}_${TAG}.log | tail -8
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n${N}_${TAG}.json 2> gpurun_out/bench_n${N}_${TAG}.err
echo "bench rc $?"; tail -c 600 gpurun_out/bench_n${N}_${TAG}.err
python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/bench_n${N}_${TAG}.json") if l.startswith("{")][-1])
    print("N=${N}", d["config"]["workload"], d["value"], "img/s", d["ms_per_step"], "ms", "e2e", d["e2e"]["value"], d["clocks"])
    for p in d["per_config"]:
        print("  ", p.get("workload"), p.get("value"), p.get("ms_per_step"), p.get("error"))
except Exception as e:
    print("bench parse failed", e)
PY
