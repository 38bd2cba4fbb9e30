#!/bin/bash
TAG=${1:-r02f}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_learner_step.py "tests/test_gpu_learner.py::test_fused_learner_forward_vs_oracle" tests/test_gpu_learner.py::test_learner_targets_random_vs_oracle -m gpu -q --tb=short 2>&1 | grep -v "^  *$" | tail -150 > gpurun_out/pytest_${TAG}.log
cat gpurun_out/pytest_${TAG}.log
