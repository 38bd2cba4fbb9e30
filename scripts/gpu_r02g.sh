#!/bin/bash
# two GPUs: the two-rank LearnerStep test (gradient exchange inside the tail kernel over CUDA-IPC peer memory) and the
# N = 2 bench line
TAG=${1:-r02g}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader
timeout 900 python -m pytest tests/test_gpu_learner_step.py "tests/test_gpu_learner.py::test_fused_learner_forward_vs_oracle" -m gpu -q --tb=short 2>&1 | grep -v "^  *$" | tail -80 > gpurun_out/pytest_${TAG}.log
cat gpurun_out/pytest_${TAG}.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_${TAG}_n2.json 2> gpurun_out/bench_${TAG}_n2.err
tail -c 2500 gpurun_out/bench_${TAG}_n2.json; tail -5 gpurun_out/bench_${TAG}_n2.err
