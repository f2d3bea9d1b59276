#!/bin/bash
# Experiment builds of the library with different k_net_tc knobs: tools/build_variant.sh <tag> <nvcc -D flags...>
set -e
cd "$(dirname "$0")/.."
TAG=$1; shift
P=othello_reinforcement_learning_test_b200
mkdir -p build_tmp
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC,-O3 --expt-relaxed-constexpr -I include "$@" -c $P/csrc/net_tc.cu -o build_tmp/net_tc_$TAG.o
OBJS=$(ls $P/build/*.o | grep -v "/net_tc.o")
nvcc -gencode arch=compute_100a,code=sm_100a --shared -o build_tmp/libothello_b200_$TAG.so $OBJS build_tmp/net_tc_$TAG.o -lcuda
echo build_tmp/libothello_b200_$TAG.so
