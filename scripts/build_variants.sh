#!/bin/bash
# Builds experimental copies of libb2cnn.so, one per "name:flags" argument, into lib/libb2cnn_<name>.so, for A/B timing
# inside ONE gpurun call (box-to-box variation is ~2 %):  scripts/build_variants.sh "nomont:-DB2CNN_MONTGOMERY=0" ...
set -e
cd "$(dirname "$0")/.."
CS=time-series-kafka-demo_b200/csrc
for spec in "$@"; do
  name=${spec%%:*}; flags=${spec#*:}
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared -cudart shared \
    $flags -o time-series-kafka-demo_b200/lib/libb2cnn_$name.so \
    $CS/b2cnn_api.cu $CS/b2cnn_generic.cu $CS/b2cnn_head.cu $CS/b2cnn_small.cu $CS/b2cnn_batch.cu $CS/b2cnn_prep.cu $CS/b2cnn_wire.cu $CS/b2cnn_train.cu $CS/b2cnn_tc.cu &
done
wait
ls time-series-kafka-demo_b200/lib/
