#!/bin/bash
# Full GPU pass: the whole -m gpu suite, the bench line, the ncu launch list and full captures of the two hot kernels.
TAG=${1:-r01}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_${TAG}.log
cat gpurun_out/pytest_${TAG}.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
tail -c 2500 gpurun_out/bench_${TAG}.json; tail -3 gpurun_out/bench_${TAG}.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_${TAG}_reference.json 2>> gpurun_out/bench_${TAG}.err
tail -c 600 gpurun_out/bench_${TAG}_reference.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
    --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 2 --warmup 3 --cpu-budget 0.5 > gpurun_out/ncu_bench_${TAG}.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rollout_tc2 -s 3 -c 1 \
    -o gpurun_out/prof_rollout_${TAG} -f python bench.py --steps 2 --warmup 3 --cpu-budget 0.5 > gpurun_out/ncu_full_${TAG}.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:learner_targets -s 1 -c 1 \
    -o gpurun_out/prof_learner_${TAG} -f python bench.py --steps 2 --warmup 3 --cpu-budget 0.5 >> gpurun_out/ncu_full_${TAG}.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:learner_fwd_tc2 -s 1 -c 1 \
    -o gpurun_out/prof_fwd_${TAG} -f python bench.py --steps 2 --warmup 3 --cpu-budget 0.5 >> gpurun_out/ncu_full_${TAG}.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:learner_bwd_tc -s 1 -c 1 \
    -o gpurun_out/prof_bwd_${TAG} -f python bench.py --steps 2 --warmup 3 --cpu-budget 0.5 >> gpurun_out/ncu_full_${TAG}.log 2>&1
ls -la gpurun_out | tail -8
