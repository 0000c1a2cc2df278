#!/bin/bash
# which of the round's switches makes the graph replay differ from the eager step in a fresh process?
run() { echo "== $*"; env "$@" timeout 300 python -m pytest tests/test_gpu_nets.py -q -s -p no:cacheprovider -k "three_steps" 2>&1 | grep "3 steps\|passed\|failed" | cut -c1-150; }
run A=1
run PGGAN_PDL=0
run PGGAN_NO_SPLITK=1
run PGGAN_NO_MASK_FUSE=1
run PGGAN_NO_POOL_COLSUM=1
run PGGAN_PDL=0 PGGAN_NO_SPLITK=1 PGGAN_NO_MASK_FUSE=1 PGGAN_NO_POOL_COLSUM=1
echo "== after the kernel tests in the same process"
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_nets.py -q -s -p no:cacheprovider -k "split_k or three_steps" 2>&1 | grep "3 steps\|passed\|failed" | cut -c1-150
