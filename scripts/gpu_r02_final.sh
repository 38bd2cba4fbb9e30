#!/bin/bash
# Full evidence pass of a commit (one GPU): the -m gpu suite, both bench arms, the ncu launch list and one full capture
# of every hot kernel of the learner step.     bash scripts/gpu_r02_final.sh <tag> [skip-tests]
TAG=${1:-r02}
mkdir -p gpurun_out
if [ -z "$2" ]; then
  timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/pytest_${TAG}.log
  cat gpurun_out/pytest_${TAG}.log
fi
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
tail -c 1500 gpurun_out/bench_${TAG}.json; tail -3 gpurun_out/bench_${TAG}.err
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_${TAG}_reference.json 2>> gpurun_out/bench_${TAG}.err
tail -c 600 gpurun_out/bench_${TAG}_reference.json
LIGHT="--steps 3 --warmup 3 --cpu-budget 0 --learner-steps 6 --fp32-steps 0 --sustained-s 0"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv \
    --log-file gpurun_out/launches_${TAG}.csv python bench.py $LIGHT > gpurun_out/ncu_bench_${TAG}.log 2>&1
for spec in rollout_tc2:rollout learner_fwd_tc2:fwd learner_bwd_f16:bwd learner_targets:targets learner_tail:tail; do
  K=${spec%%:*}; N=${spec##*:}
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:$K -s 4 -c 1 \
      -o gpurun_out/prof_${N}_${TAG} -f python bench.py $LIGHT > gpurun_out/ncu_${N}_${TAG}.log 2>&1
  tail -2 gpurun_out/ncu_${N}_${TAG}.log | cut -c1-160
done
ls -la gpurun_out | tail -12
