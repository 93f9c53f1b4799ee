#!/bin/bash
# A/B of variant libraries on one box (no tests): base, variants, base; prints step, kernel and head times
mkdir -p gpurun_out
L=$PWD/time-series-kafka-demo_b200/lib
for name in base ${VARIANTS} base; do
  if [ $name = base ]; then lib=$L/libb2cnn.so; else lib=$L/libb2cnn_$name.so; fi
  B2CNN_LIB=$lib timeout -k 10 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --e2e-steps 1 --sustained-seconds 0 --parity-windows 64 --extra-steps 0 > gpurun_out/ab_$name.json 2>> gpurun_out/ab.err
  echo "$name: $(grep -o '"ms_per_step": [0-9.]*\|"kernel_ms": [0-9.]*\|"head_ms": [0-9.]*\|"max_rel": [0-9.e-]*' gpurun_out/ab_$name.json | tr '\n' ' ')"
done
tail -2 gpurun_out/ab.err
