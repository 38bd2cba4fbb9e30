#!/bin/bash
TAG=${1:-r02n8b}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 20 --warmup 5 --learner-steps 100 --fp32-steps 0 --sustained-s 0.5 > gpurun_out/bench_${TAG}_cfg2.json 2> gpurun_out/bench_${TAG}_cfg2.err
tail -c 1500 gpurun_out/bench_${TAG}_cfg2.json | head -c 600; grep -v "OMP_NUM_THREADS\|^\*\*\*\|^$" gpurun_out/bench_${TAG}_cfg2.err | tail -3
