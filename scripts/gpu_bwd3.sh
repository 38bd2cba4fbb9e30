#!/bin/bash
# backward kernels: tests, then timings of the default library and of every variant library in r-nad_b200/lib
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_learner.py tests/test_gpu_learner_step.py -m gpu -q --tb=short -x 2>&1 | tail -12 | tee gpurun_out/pytest_bwd3.log
{
for a in 3 2; do
  echo "A=$a default";  A=$a timeout 120 python scripts/time_learner_kernels.py 2>&1 | tail -2
  echo "A=$a tf32 split";  A=$a RNAD_LEARNER_BWD_TF32=1 timeout 120 python scripts/time_learner_kernels.py 2>&1 | tail -1
  for lib in r-nad_b200/lib/librnad_b200_*.so; do
    [ -f $lib ] || continue
    echo "A=$a $lib"; A=$a RNAD_B200_LIB=$PWD/$lib timeout 120 python scripts/time_learner_kernels.py 2>&1 | tail -2
  done
done
} | tee gpurun_out/bwd3_ab.log
