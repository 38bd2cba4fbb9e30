#!/bin/bash
TAG=${1:-r02o}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_learner.py tests/test_gpu_learner_step.py tests/test_gpu_env_rollout.py -m gpu -q --tb=short 2>&1 | grep -v "^  *$" | tail -40 > gpurun_out/pytest_${TAG}.log
cat gpurun_out/pytest_${TAG}.log
for v in reuse full; do
  if [ $v = full ]; then export RNAD_STEP_REUSE=0; fi
  timeout 600 python bench.py --steps 10 --warmup 3 --cpu-budget 0 --fp32-steps 0 --sustained-s 0 > gpurun_out/bench_${TAG}_${v}.json 2> gpurun_out/bench_${TAG}_${v}.err
  python - <<P
import json
d=json.loads(open('gpurun_out/bench_${TAG}_${v}.json').read().strip().splitlines()[-1])
l=d['learner']; print('${v}', l['ms_per_update'], l['free_running']['ms_per_update'], {k:round(x['ms'],4) for k,x in l['roofline']['kernels'].items()})
P
  tail -2 gpurun_out/bench_${TAG}_${v}.err
done
unset RNAD_STEP_REUSE
bash scripts/gpu_cfg5.sh
