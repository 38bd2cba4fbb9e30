#!/bin/bash
# e2e A/B: returns written by the kernel into pinned host memory (default) vs a copy node behind the kernel
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_env_rollout.py -m gpu -q -x -k "selfplay or SelfPlay or self_play" 2>&1 | tail -2
for v in direct copy copyw; do
  unset RNAD_SELFPLAY_COPY_RETURNS RNAD_SELFPLAY_COPY_WEIGHTS; if [ $v = copy ]; then export RNAD_SELFPLAY_COPY_RETURNS=1 RNAD_SELFPLAY_COPY_WEIGHTS=1; fi; if [ $v = copyw ]; then export RNAD_SELFPLAY_COPY_WEIGHTS=1; fi
  timeout 300 python bench.py --steps 20 --warmup 5 --cpu-budget 0 --fp32-steps 0 --sustained-s 0 --learner-steps 20 > gpurun_out/bench_e2e_$v.json 2> gpurun_out/bench_e2e_$v.err
  python - <<P
import json
d=json.loads(open('gpurun_out/bench_e2e_$v.json').read().strip().splitlines()[-1])
print('$v', 'kernel ms', round(d['ms_per_step'],4), 'e2e ms', round(d['e2e']['ms_per_step'],4))
P
done
