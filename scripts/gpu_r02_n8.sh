#!/bin/bash
# eight GPUs: BASELINE config 4 (depth 8, A = 4, 1,048,576 games = 131,072 per GPU, gradient exchange inside the tail
# kernel) and the default cfg2 line, both through bench.py under torchrun
TAG=${1:-r02n8}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
nvidia-smi topo -m 2>/dev/null | head -12
for cfg in cfg4 cfg2; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --config $cfg --steps 20 --warmup 5 --learner-steps 100 --fp32-steps 0 --sustained-s 0.5 > gpurun_out/bench_${TAG}_${cfg}.json 2> gpurun_out/bench_${TAG}_${cfg}.err
  tail -c 3000 gpurun_out/bench_${TAG}_${cfg}.json; grep -v "OMP_NUM_THREADS\|^\*\*\*\|^$" gpurun_out/bench_${TAG}_${cfg}.err | tail -5
done
