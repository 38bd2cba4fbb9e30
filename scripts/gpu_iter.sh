#!/bin/bash
# Quick GPU iteration: selected parity tests, then the bench line.
TAG=${1:-iter}
TESTS=${2:-tests}
mkdir -p gpurun_out
timeout 1200 python -m pytest $TESTS -m gpu -x -q -s 2>&1 | tail -40 > gpurun_out/pytest_${TAG}.log
cat gpurun_out/pytest_${TAG}.log
timeout 600 python bench.py --steps 20 --warmup 5 --cpu-budget 2 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_${TAG}.json"))
print("value %.4g env_steps/s  ms/step %.4f  e2e %.4g  roofline %.3f  learner %.1f upd/s (%.3f ms)  cpu %.3g" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"], d["learner"]["updates_per_sec"], d["learner"]["ms_per_update"], d["cpu_baseline"]["value"]))
PY
tail -3 gpurun_out/bench_${TAG}.err
