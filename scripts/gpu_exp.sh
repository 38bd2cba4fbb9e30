#!/bin/bash
# timing experiments over the variant libraries in r-nad_b200/lib (scripts/build_variant.sh)
mkdir -p gpurun_out
timeout 120 python scripts/time_rollout.py | tee gpurun_out/exp.log
for lib in r-nad_b200/lib/librnad_b200_*.so; do
  RNAD_B200_LIB=$PWD/$lib timeout 120 python scripts/time_rollout.py 2>&1 | tail -1 | tee -a gpurun_out/exp.log
done
