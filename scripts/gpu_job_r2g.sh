#!/bin/bash
# round 2, job G: parity tests of the current build, A/B against variants, wait-time counters of the timing build
mkdir -p gpurun_out
timeout -k 10 300 python -m pytest tests/test_gpu_tc.py tests/test_gpu_parity.py -q -x --timeout 120 > gpurun_out/g_tc.log 2>&1; rc=$?; echo "pytest rc=$rc" >> gpurun_out/g_tc.log
tail -5 gpurun_out/g_tc.log
L=$PWD/time-series-kafka-demo_b200/lib
for name in base ${VARIANTS} base; do
  if [ $name = base ]; then lib=$L/libb2cnn.so; else lib=$L/libb2cnn_$name.so; fi
  B2CNN_LIB=$lib timeout -k 10 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --e2e-steps 1 --sustained-seconds 0 --parity-windows 64 --extra-steps 0 > gpurun_out/ab_$name.json 2>> gpurun_out/ab.err
  echo "$name: $(grep -o '"ms_per_step": [0-9.]*\|"kernel_ms": [0-9.]*\|"max_rel": [0-9.e-]*' gpurun_out/ab_$name.json | tr '\n' ' ')"
done
tail -2 gpurun_out/ab.err
if [ -f $L/libb2cnn_timing.so ]; then CTAS=${CTAS:-148} B2CNN_LIB=$L/libb2cnn_timing.so timeout 200 python scripts/fused_timing.py 2>&1 | tail -7; fi
