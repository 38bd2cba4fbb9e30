#!/bin/bash
TAG=${1:-r02e}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --durations=5 2>&1 | tail -60 > gpurun_out/pytest_${TAG}.log
cat gpurun_out/pytest_${TAG}.log
timeout 600 python bench.py --steps 20 --warmup 5 --cpu-budget 2 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
python - <<'P'
import json,sys
d=json.loads(open('gpurun_out/bench_'+sys.argv[1]+'.json' if len(sys.argv)>1 else 'x').read().strip().splitlines()[-1]) if False else None
P
tail -c 3000 gpurun_out/bench_${TAG}.json; tail -3 gpurun_out/bench_${TAG}.err
