#!/bin/bash
# the fp16-kind rollout engine: parity tests, error statistics, timing against the tf32 engine
TAG=${1:-f16}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_env_rollout.py -m gpu -q --tb=short -k "f16x2" -x 2>&1 | tail -25 > gpurun_out/pytest_${TAG}.log
cat gpurun_out/pytest_${TAG}.log
timeout 300 python scripts/precision_compare.py 2>&1 | tail -12 | tee gpurun_out/precision_${TAG}.log
for prec in tf32x2 f16x2; do
  timeout 300 python bench.py --steps 20 --warmup 5 --cpu-budget 0 --fp32-steps 0 --sustained-s 1 --learner-steps 20 --precision $prec > gpurun_out/bench_${TAG}_${prec}.json 2> gpurun_out/bench_${TAG}_${prec}.err
  python - <<P
import json
d=json.loads(open('gpurun_out/bench_${TAG}_${prec}.json').read().strip().splitlines()[-1])
print('${prec}', 'ms', d['ms_per_step'], 'sustained', d['sustained']['ms_per_step'], 'e2e', d['e2e']['ms_per_step'])
P
  tail -2 gpurun_out/bench_${TAG}_${prec}.err
done
