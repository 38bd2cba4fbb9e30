#!/bin/bash
# Round 2, second GPU session: the suite after the ABI v2 changes (K3 per-slot kernel, rollout stats, LearnerStep),
# K1 standalone with and without the staged table top, a first bench line of the new learner step.
TAG=${1:-r02b}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --durations=8 2>&1 | tail -40 > gpurun_out/pytest_${TAG}.log
cat gpurun_out/pytest_${TAG}.log
timeout 600 python bench.py --steps 20 --warmup 5 --cpu-budget 2 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
tail -c 2500 gpurun_out/bench_${TAG}.json; tail -3 gpurun_out/bench_${TAG}.err
