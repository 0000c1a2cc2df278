#!/bin/bash
TAG=${1:-r2d}
mkdir -p gpurun_out
timeout 600 python tools/splitk_micro.py > gpurun_out/splitk_${TAG}.txt 2>&1; echo "rc $?"; cat gpurun_out/splitk_${TAG}.txt | cut -c1-400
