#!/bin/bash
# timing experiments over the variant libraries in r-nad_b200/lib: the learner kernels at cfg2
mkdir -p gpurun_out
echo "default"; timeout 120 python scripts/time_learner_kernels.py 2>&1 | tail -1
for lib in r-nad_b200/lib/librnad_b200_*.so; do
  echo $lib; RNAD_B200_LIB=$PWD/$lib timeout 120 python scripts/time_learner_kernels.py 2>&1 | tail -1
done
