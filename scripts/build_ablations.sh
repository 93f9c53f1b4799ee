#!/bin/bash
# Builds instrumented copies of libb2cnn.so (B2CNN_ABLATE=1,2,3; results are garbage, timing only).
# Used once to find which side bounds the fused kernel; select with B2CNN_LIB=<path> python bench.py ...
set -e
cd "$(dirname "$0")/.."
CS=time-series-kafka-demo_b200/csrc
for a in ${ABLATIONS:-1 2 3}; do
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared -cudart shared \
    -DB2CNN_ABLATE=$a -o time-series-kafka-demo_b200/lib/libb2cnn_ab$a.so \
    $CS/b2cnn_api.cu $CS/b2cnn_generic.cu $CS/b2cnn_head.cu $CS/b2cnn_small.cu $CS/b2cnn_prep.cu $CS/b2cnn_tc.cu &
done
wait
ls -la time-series-kafka-demo_b200/lib/
