#!/bin/bash
# GPU pass for the mbarrier-ordered kernels under compute-sanitizer (not run in round 1: no GPU minutes were left).
#   bash scripts/gpu_sanitize.sh [tag]      under gpurun, one GPU; small shapes - the tools slow kernels down 10-100x
# synccheck: invalid barrier / mbarrier usage; racecheck: shared-memory hazards (weight image, staging, candidates);
# memcheck: out-of-bounds global / shared accesses.  Tensor-memory hazards are NOT covered by these tools - they are what
# tests/test_gpu_env_rollout.py::test_fused_rollout_repeated_launches_have_no_ordering_race and
# scripts/stress_rollout.py look for.
TAG=${1:-san}
mkdir -p gpurun_out
for tool in memcheck synccheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 1 \
      python -m pytest tests/test_gpu_env_rollout.py tests/test_gpu_learner.py -m gpu -x -q \
      -k "(width256 and 128) or (fused_learner and (300 or 130))" > gpurun_out/sanitize_${TAG}_${tool}.log 2>&1
  echo "$tool: exit $?"; tail -5 gpurun_out/sanitize_${TAG}_${tool}.log
done
