#!/bin/bash
mkdir -p gpurun_out
timeout 3000 python scripts/cfg5_sweep.py --out gpurun_out/cfg5_sweep.json > gpurun_out/cfg5_sweep.log 2>&1
tail -40 gpurun_out/cfg5_sweep.log
