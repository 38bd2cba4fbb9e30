#!/bin/bash
# Round 2, first GPU session: the -m gpu suite (with the benchmark-config parity tests), both bench arms, the
# reference's main.py unchanged, K1 standalone numbers + ncu, and the compute-sanitizer pass.
TAG=${1:-r02a}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv,noheader
nproc
timeout 1500 python -m pytest tests -m gpu -x -q --durations=8 2>&1 | tail -25 > gpurun_out/pytest_${TAG}.log
cat gpurun_out/pytest_${TAG}.log
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_${TAG}_reference.json 2> gpurun_out/bench_${TAG}_reference.err
tail -c 1500 gpurun_out/bench_${TAG}_reference.json; tail -2 gpurun_out/bench_${TAG}_reference.err
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
tail -c 2500 gpurun_out/bench_${TAG}.json; tail -3 gpurun_out/bench_${TAG}.err
( cd /tmp && timeout 900 python $GRAFT_REPO_ROOT/scripts/run_reference_main.py > $GRAFT_REPO_ROOT/gpurun_out/reference_main_${TAG}.log 2>&1; echo "reference main.py exit $?" )
grep -c "NashConv at" gpurun_out/reference_main_${TAG}.log; grep "nash_conv\|finished\|Error\|error" gpurun_out/reference_main_${TAG}.log | tail -5
tail -3 gpurun_out/reference_main_${TAG}.log
timeout 600 python scripts/bench_k1.py > gpurun_out/k1_${TAG}.json 2> gpurun_out/k1_${TAG}.err; tail -c 1500 gpurun_out/k1_${TAG}.json; tail -2 gpurun_out/k1_${TAG}.err
timeout 600 ncu --set full --clock-control none -k regex:"observe_kernel|step_kernel" -s 20 -c 2 \
    -o gpurun_out/prof_k1_${TAG} -f python scripts/bench_k1.py --reps 2 > gpurun_out/ncu_k1_${TAG}.log 2>&1
bash scripts/gpu_sanitize.sh ${TAG}
ls -la gpurun_out | tail -12
