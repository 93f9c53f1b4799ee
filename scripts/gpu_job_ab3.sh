#!/bin/bash
# A/B of variant libraries with longer timed regions and interleaved repeats: STEPS (default 100), ROUNDS (default 3)
mkdir -p gpurun_out
L=$PWD/time-series-kafka-demo_b200/lib
for r in $(seq 1 ${ROUNDS:-3}); do
for name in base ${VARIANTS}; do
  if [ $name = base ]; then lib=$L/libb2cnn.so; else lib=$L/libb2cnn_$name.so; fi
  B2CNN_LIB=$lib timeout -k 10 200 python bench.py --steps ${STEPS:-100} --warmup 5 --no-cpu-baseline --e2e-steps 1 --sustained-seconds 0 --parity-windows 64 --extra-steps 0 > gpurun_out/ab3_${name}_$r.json 2>> gpurun_out/ab3.err
  echo "$r $name: $(grep -o '"ms_per_step": [0-9.]*\|"kernel_ms": [0-9.]*\|"head_ms": [0-9.]*\|"sm_mhz": [0-9.]*' gpurun_out/ab3_${name}_$r.json | tr '\n' ' ')"
done
done
tail -2 gpurun_out/ab3.err
