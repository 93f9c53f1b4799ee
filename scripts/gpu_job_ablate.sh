#!/bin/bash
# ablation timing of the fused kernel (instrumented libraries from scripts/build_ablations.sh; garbage results, kernel_ms only)
mkdir -p gpurun_out
L=$PWD/time-series-kafka-demo_b200/lib
for a in ${ABL:-0 1 2 3}; do
  if [ $a = 0 ]; then lib=$L/libb2cnn.so; else lib=$L/libb2cnn_ab$a.so; fi
  B2CNN_LIB=$lib timeout -k 10 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/j_ab${a}.json 2>> gpurun_out/j.err
  echo "ablate $a: $(grep -o '"ms_per_step": [0-9.]*\|"kernel_ms": [0-9.]*' gpurun_out/j_ab${a}.json | tr '\n' ' ')"
done
tail -3 gpurun_out/j.err
