#!/bin/bash
# last evidence pass of the round: full -m gpu suite, the native bench line, the sanitizer, one full capture of the backward
TAG=${1:-r02z}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/pytest_${TAG}.log; cat gpurun_out/pytest_${TAG}.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; tail -c 400 gpurun_out/bench_${TAG}.json; tail -2 gpurun_out/bench_${TAG}.err
bash scripts/gpu_sanitize.sh ${TAG}
LIGHT="--steps 3 --warmup 3 --cpu-budget 0 --learner-steps 6 --fp32-steps 0 --sustained-s 0"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:learner_bwd_f16 -s 4 -c 1 -o gpurun_out/prof_bwd_${TAG} -f python bench.py $LIGHT > gpurun_out/ncu_bwd_${TAG}.log 2>&1
tail -2 gpurun_out/ncu_bwd_${TAG}.log | cut -c1-160
