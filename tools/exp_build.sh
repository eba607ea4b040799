#!/bin/bash
# Tuning aid: builds tools/exp/libsrwn_<name>.so with fused_bf16.cu compiled under -DSRWN_EXP=<bits>
# (see the SRWN_EXP switches in csrc/fused_bf16.cu).  Use with SRWN_LIB=tools/exp/libsrwn_<name>.so.
set -e
cd "$(dirname "$0")/.."
mkdir -p tools/exp
C=sr-wavenet_b200/csrc
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -DSRWN_TUNING -DSRWN_EXP=$1 -c $C/fused_bf16.cu -o tools/exp/fused_$2.o
objs=""
for f in api stack_f32 mol ops_generic ar_generate ar_mma train_f32 train_tc stft_loss encoder random; do objs="$objs $C/$f.o"; done
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o tools/exp/libsrwn_$2.so $objs tools/exp/fused_$2.o -lcudart_static -ldl -lrt -lpthread
