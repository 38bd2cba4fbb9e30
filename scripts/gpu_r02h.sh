#!/bin/bash
# cfg3 and cfg4 (one GPU's share) through bench.py, K1 standalone after the kernel rewrite (plain and staged, two batch sizes)
TAG=${1:-r02h}
mkdir -p gpurun_out
for cfg in cfg3 cfg4; do
  timeout 900 python bench.py --config $cfg --steps 10 --warmup 3 --learner-steps 50 --cpu-budget 0 --fp32-steps 0 --sustained-s 0.5 > gpurun_out/bench_${TAG}_${cfg}.json 2> gpurun_out/bench_${TAG}_${cfg}.err
  tail -c 3500 gpurun_out/bench_${TAG}_${cfg}.json; tail -3 gpurun_out/bench_${TAG}_${cfg}.err
done
timeout 300 python scripts/bench_k1.py > gpurun_out/k1_${TAG}.json 2> gpurun_out/k1_${TAG}.err; tail -c 1500 gpurun_out/k1_${TAG}.json; tail -2 gpurun_out/k1_${TAG}.err
RNAD_K1_STAGE=1 timeout 300 python scripts/bench_k1.py > gpurun_out/k1_${TAG}_staged.json 2> gpurun_out/k1_${TAG}_staged.err; tail -c 1500 gpurun_out/k1_${TAG}_staged.json
timeout 300 python scripts/bench_k1.py --batch 2097152 > gpurun_out/k1_${TAG}_2m.json 2> gpurun_out/k1_${TAG}_2m.err; tail -c 1500 gpurun_out/k1_${TAG}_2m.json; tail -2 gpurun_out/k1_${TAG}_2m.err
timeout 600 ncu --set full --clock-control none -k regex:"observe_kernel|step_kernel" -s 20 -c 4 \
    -o gpurun_out/prof_k1_${TAG} -f python scripts/bench_k1.py --reps 2 > gpurun_out/ncu_k1_${TAG}.log 2>&1
