#!/bin/bash
# one full ncu capture of a named kernel of the learner step: bash scripts/gpu_ncu1.sh <regex> <tag>
K=$1; TAG=$2
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:$K -s 4 -c 1 \
    -o gpurun_out/prof_${TAG} -f python bench.py --steps 3 --warmup 3 --cpu-budget 0 --learner-steps 6 --fp32-steps 0 --sustained-s 0 > gpurun_out/ncu_${TAG}.log 2>&1
tail -3 gpurun_out/ncu_${TAG}.log | cut -c1-200
