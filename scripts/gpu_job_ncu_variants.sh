#!/bin/bash
# full ncu capture (with source) of the fused kernel for each variant library in $VARIANTS (base = the shipped library)
mkdir -p gpurun_out
L=$PWD/time-series-kafka-demo_b200/lib
for name in ${VARIANTS}; do
  if [ $name = base ]; then lib=$L/libb2cnn.so; else lib=$L/libb2cnn_$name.so; fi
  B2CNN_LIB=$lib timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:tc_fused -s 2 -c 1 -f -o gpurun_out/ncu_$name \
    python bench.py --steps 1 --warmup 3 --e2e-steps 1 --no-cpu-baseline --sustained-seconds 0 --parity-windows 0 > gpurun_out/ncu_$name.log 2>&1
  tail -1 gpurun_out/ncu_$name.log | cut -c1-200
done
