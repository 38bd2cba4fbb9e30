#!/bin/bash
# ncu launch list + full captures of the rollout kernel and the learner-targets kernel (GPU box).
TAG=${1:-r01}
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
    --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 2 --warmup 3 --cpu-budget 0.5 > gpurun_out/ncu_bench_${TAG}.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rollout_tc -s 3 -c 1 \
    -o gpurun_out/prof_rollout_${TAG} -f python bench.py --steps 2 --warmup 3 --cpu-budget 0.5 > gpurun_out/ncu_full_${TAG}.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:learner_targets -s 1 -c 1 \
    -o gpurun_out/prof_learner_${TAG} -f python bench.py --steps 2 --warmup 3 --cpu-budget 0.5 >> gpurun_out/ncu_full_${TAG}.log 2>&1
ls -la gpurun_out | tail -8
