#!/bin/bash
# tests, bench, ncu launch list of one bench run, full captures of the K3 and tail kernels
TAG=${1:-r02d}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --durations=5 2>&1 | tail -40 > gpurun_out/pytest_${TAG}.log
cat gpurun_out/pytest_${TAG}.log
timeout 600 python bench.py --steps 20 --warmup 5 --cpu-budget 2 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
tail -c 4500 gpurun_out/bench_${TAG}.json; tail -3 gpurun_out/bench_${TAG}.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv \
    --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 3 --warmup 3 --cpu-budget 0 --learner-steps 6 --fp32-steps 0 --sustained-s 0 > gpurun_out/ncu_bench_${TAG}.log 2>&1
tail -2 gpurun_out/ncu_bench_${TAG}.log | cut -c1-300
timeout 600 ncu --set full --clock-control none --import-source on -k regex:learner_targets_tb -s 4 -c 1 \
    -o gpurun_out/prof_targets_${TAG} -f python bench.py --steps 3 --warmup 3 --cpu-budget 0 --learner-steps 6 --fp32-steps 0 --sustained-s 0 > gpurun_out/ncu_full_${TAG}.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:learner_tail -s 4 -c 1 \
    -o gpurun_out/prof_tail_${TAG} -f python bench.py --steps 3 --warmup 3 --cpu-budget 0 --learner-steps 6 --fp32-steps 0 --sustained-s 0 >> gpurun_out/ncu_full_${TAG}.log 2>&1
ls -la gpurun_out | tail -6
