#!/bin/bash
# Tuning aid: builds tools/exp/libsrwn_var<bits>.so with fused_bf16.cu compiled under -DSRWN_VAR=<bits> (timing-only
# variants of the ring hand-off: 1 = no row counting, 2 = loader does not wait for flags, 4 = publisher idle).
# Use with SRWN_LIB=tools/exp/libsrwn_var<bits>.so python tools/prof_fused.py ...   Results of these builds are WRONG.
set -e
cd "$(dirname "$0")/.."
mkdir -p tools/exp
C=sr-wavenet_b200/csrc
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -DSRWN_VAR=$1 -c $C/fused_bf16.cu -o tools/exp/fused_var$1.o
objs=""
for f in api stack_f32 mol ops_generic ar_generate ar_mma train_f32 train_tc stft_loss encoder random; do objs="$objs $C/$f.o"; done
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o tools/exp/libsrwn_var$1.so $objs tools/exp/fused_var$1.o -lcudart_static -ldl -lrt -lpthread
