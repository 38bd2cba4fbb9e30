#!/bin/bash
# development: a library with ONE source compiled under extra defines, linked with the default objects
#   bash scripts/build_variant_src.sh <name> <source.cu> <defines...>  ->  r-nad_b200/lib/librnad_b200_<name>.so
set -e
cd "$(dirname "$0")/../r-nad_b200/csrc"
NAME=$1; SRC=$2; shift; shift
mkdir -p ../build/$NAME
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -fvisibility=hidden -cudart shared "$@" -c $SRC -o ../build/$NAME/${SRC%.cu}.o
OBJS=$(ls ../build/*.o | grep -v "/${SRC%.cu}.o")
nvcc -gencode arch=compute_100a,code=sm_100a -shared -cudart shared -Xlinker -rpath=/usr/local/cuda/lib64 -o ../lib/librnad_b200_$NAME.so $OBJS ../build/$NAME/${SRC%.cu}.o
