#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_learner.py tests/test_gpu_learner_step.py -m gpu -q -x 2>&1 | tail -2
python scripts/time_k3.py | tail -1
timeout 600 compute-sanitizer --tool synccheck --error-exitcode 1 python -m pytest tests/test_gpu_learner.py -m gpu -x -q -k "fused_learner and (300 or 130)" > gpurun_out/sanitize_chk_synccheck.log 2>&1; echo "synccheck exit $?"; grep -v "^=========     " gpurun_out/sanitize_chk_synccheck.log | tail -6
