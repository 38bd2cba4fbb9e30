#!/bin/bash
# Run on the GPU box (via gpurun): bench line, ncu launch list, full ncu capture of the rollout kernel.
set -x
mkdir -p gpurun_out
TAG=${1:-r01}
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
tail -c 3000 gpurun_out/bench_${TAG}.json
tail -5 gpurun_out/bench_${TAG}.err
timeout 600 python bench.py --steps 20 --warmup 5 --precision fp32 --cpu-budget 1 > gpurun_out/bench_${TAG}_fp32.json 2>> gpurun_out/bench_${TAG}.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 2 --warmup 3 --cpu-budget 0.5 > gpurun_out/ncu_bench_${TAG}.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rollout_tc -s 3 -c 2 \
    -o gpurun_out/prof_rollout_${TAG} -f python bench.py --steps 2 --warmup 3 --cpu-budget 0.5 > gpurun_out/ncu_full_${TAG}.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:learner_targets -s 1 -c 1 \
    -o gpurun_out/prof_learner_${TAG} -f python bench.py --steps 2 --warmup 3 --cpu-budget 0.5 >> gpurun_out/ncu_full_${TAG}.log 2>&1
ls -la gpurun_out
