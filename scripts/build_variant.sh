#!/bin/bash
# development: a library with rollout_tc2.cu compiled under extra defines, linked with the default objects
#   bash scripts/build_variant.sh <name> <defines...>   ->  r-nad_b200/lib/librnad_b200_<name>.so  (RNAD_B200_LIB selects it)
set -e
cd "$(dirname "$0")/../r-nad_b200/csrc"
NAME=$1; shift
mkdir -p ../build/$NAME
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -fvisibility=hidden -cudart shared "$@" -c rollout_tc2.cu -o ../build/$NAME/rollout_tc2.o
OBJS=$(ls ../build/*.o | grep -v rollout_tc2.o)
nvcc -gencode arch=compute_100a,code=sm_100a -shared -cudart shared -Xlinker -rpath=/usr/local/cuda/lib64 -o ../lib/librnad_b200_$NAME.so $OBJS ../build/$NAME/rollout_tc2.o
