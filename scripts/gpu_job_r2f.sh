#!/bin/bash
# round 2, job F: tensor-core parity tests of the current build, A/B against variant libraries, full ncu capture (with source) of the fused kernel
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests/test_gpu_tc.py tests/test_gpu_parity.py -q -x > gpurun_out/f_tc.log 2>&1; rc=$?; echo "pytest rc=$rc" >> gpurun_out/f_tc.log
tail -5 gpurun_out/f_tc.log
L=$PWD/time-series-kafka-demo_b200/lib
for name in base ${VARIANTS} base; do
  if [ $name = base ]; then lib=$L/libb2cnn.so; else lib=$L/libb2cnn_$name.so; fi
  B2CNN_LIB=$lib timeout -k 10 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --e2e-steps 1 --sustained-seconds 0 --parity-windows 64 --extra-steps 0 > gpurun_out/ab_$name.json 2>> gpurun_out/ab.err
  echo "$name: $(grep -o '"ms_per_step": [0-9.]*\|"kernel_ms": [0-9.]*\|"max_rel": [0-9.e-]*' gpurun_out/ab_$name.json | tr '\n' ' ')"
done
tail -2 gpurun_out/ab.err
if [ "$NCU" != "" ] && [ $rc -eq 0 ]; then
timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:tc_fused -s 2 -c 1 -f -o gpurun_out/f_fused \
    python bench.py --steps 1 --warmup 3 --e2e-steps 1 --no-cpu-baseline --sustained-seconds 0 --parity-windows 0 > gpurun_out/f_ncu_full.log 2>&1
tail -2 gpurun_out/f_ncu_full.log | cut -c1-200
fi
