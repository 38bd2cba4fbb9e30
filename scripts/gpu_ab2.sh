#!/bin/bash
# tests of the learner kernels, then the per-kernel times of the step for the default library and every variant library
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_learner.py tests/test_gpu_learner_step.py -m gpu -q --tb=short -x 2>&1 | tail -6
echo default; timeout 200 python scripts/time_k3.py | tail -1
for lib in r-nad_b200/lib/librnad_b200_*.so; do
  [ -f $lib ] || continue
  RNAD_B200_LIB=$PWD/$lib timeout 200 python scripts/time_k3.py | tail -1
done
