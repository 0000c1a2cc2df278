#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_nets.py tests/test_gpu_golden.py -q -x -p no:cacheprovider 2>&1 | tail -2
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -x -p no:cacheprovider -k "rgb or elementwise or split_k" 2>&1 | tail -1
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1 | cut -c1-200
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-per-config 2> /dev/null | python -c "import sys,json; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('C4', d['value'], d['ms_per_step'], d['clocks'], d['gpu_launches'])"
