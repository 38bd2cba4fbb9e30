#!/bin/bash
# GPU iteration on the rollout engines: parity tests of the fused rollout, then kernel timings per precision.
TAG=${1:-riter}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_env_rollout.py tests/test_gpu_learner.py -m gpu -x -q -k "${2:-width256 or fused_learner}" 2>&1 | tail -40 > gpurun_out/pytest_${TAG}.log
cat gpurun_out/pytest_${TAG}.log
for prec in tf32 tf32x2; do
  timeout 300 python bench.py --steps 20 --warmup 5 --cpu-budget 0.5 --precision $prec > gpurun_out/bench_${TAG}_${prec}.json 2> gpurun_out/bench_${TAG}_${prec}.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_${TAG}_${prec}.json"))
    print("$prec: value %.4g env_steps/s  ms/step %.4f  e2e %.4g (%.3f ms)  roofline %.3f  learner %.1f upd/s (%.3f ms)" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["roofline"]["frac"], d["learner"]["updates_per_sec"], d["learner"]["ms_per_update"]))
except Exception as e:
    print("$prec: bench failed", e)
PY
  tail -3 gpurun_out/bench_${TAG}_${prec}.err
done
