#!/bin/bash
# Tuning aid: builds tools/exp/libsrwn_<name>.so with fused_bf16.cu compiled under -DSRWN_EXP=<bits>
# (see the SRWN_EXP switches in csrc/fused_bf16.cu).  Use with SRWN_LIB=tools/exp/libsrwn_<name>.so.
set -e
cd "$(dirname "$0")/.."
mkdir -p tools/exp
C=sr-wavenet_b200/csrc
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -DSRWN_EXP=$1 -c $C/fused_bf16.cu -o tools/exp/fused_$2.o
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o tools/exp/libsrwn_$2.so $C/api.o $C/stack_f32.o $C/mol.o $C/ops_generic.o $C/ar_generate.o $C/ar_mma.o $C/train_f32.o tools/exp/fused_$2.o -lcudart_static -ldl -lrt -lpthread
