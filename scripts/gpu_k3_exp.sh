#!/bin/bash
mkdir -p gpurun_out
for nb in 0 1184 592 296; do
  echo "blocks=$nb default"; RNAD_K3_BLOCKS=$nb timeout 200 python scripts/time_k3.py | tail -1
  echo "blocks=$nb prefetch"; RNAD_K3_BLOCKS=$nb RNAD_B200_LIB=$PWD/r-nad_b200/lib/librnad_b200_k3pf.so timeout 200 python scripts/time_k3.py | tail -1
done
