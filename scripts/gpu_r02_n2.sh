#!/bin/bash
# two GPUs: the two-rank learner-step test (exchange inside the tail kernel == one rank on the concatenated batch) and the
# default cfg2 bench line under torchrun
TAG=${1:-r02n2}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_learner_step.py -m gpu -q -x -k two_rank 2>&1 | tail -3 | tee gpurun_out/two_rank_${TAG}.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 20 --warmup 5 --fp32-steps 0 --sustained-s 0.5 > gpurun_out/bench_${TAG}_cfg2.json 2> gpurun_out/bench_${TAG}_cfg2.err
tail -c 2500 gpurun_out/bench_${TAG}_cfg2.json; grep -v "OMP_NUM_THREADS\|^\*\*\*\|^$" gpurun_out/bench_${TAG}_cfg2.err | tail -5
