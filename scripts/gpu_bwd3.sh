#!/bin/bash
# A/B of the backward kernels: tests, then timings of the mask formulation (default) and the S^T one (RNAD_LEARNER_BWD_V2)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_learner.py tests/test_gpu_learner_step.py -m gpu -q --tb=short -x 2>&1 | tail -8 | tee gpurun_out/pytest_bwd3.log
for a in 3 2; do
  echo "A=$a default";  A=$a timeout 120 python scripts/time_learner_kernels.py 2>&1 | tail -2
  echo "A=$a V2";       A=$a RNAD_LEARNER_BWD_V2=1 timeout 120 python scripts/time_learner_kernels.py 2>&1 | tail -1
done | tee gpurun_out/bwd3_ab.log
