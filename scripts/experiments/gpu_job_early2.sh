#!/bin/bash
# early-head chain: tensor-core tests, timelines of the early and the serial chain, then A/B on this box
mkdir -p gpurun_out
L=$PWD/time-series-kafka-demo_b200/lib
timeout -k 10 400 python -m pytest tests/test_gpu_tc.py -q -x > gpurun_out/e_tc.log 2>&1; echo "tc pytest rc=$?" >> gpurun_out/e_tc.log
tail -4 gpurun_out/e_tc.log
grep -q "rc=0" gpurun_out/e_tc.log || exit 1
rm -f gpurun_out/early_timeline2.txt
for cfg in "etiming 1"; do
  set -- $cfg
  echo "== lib $1 early_head=$2" | tee -a gpurun_out/early_timeline2.txt
  B2CNN_LIB=$L/libb2cnn_$1.so B2CNN_EARLY_HEAD=$2 timeout -k 10 400 python -X faulthandler scripts/early_timing.py 2>&1 | tail -8 | tee -a gpurun_out/early_timeline2.txt
done
for r in 1 2 3; do
for v in "base 1" "poll4 1" "base 0"; do
  set -- $v
  if [ $1 = base ]; then lib=$L/libb2cnn.so; else lib=$L/libb2cnn_$1.so; fi
  B2CNN_LIB=$lib B2CNN_EARLY_HEAD=$2 timeout -k 10 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --e2e-steps 1 --sustained-seconds 0 --parity-windows 64 --extra-steps 0 > gpurun_out/early2_$1_$2_$r.json 2>> gpurun_out/early2.err
  echo "$r $1 early_head=$2: $(grep -o '"ms_per_step": [0-9.]*\|"kernel_ms": [0-9.]*' gpurun_out/early2_$1_$2_$r.json | tr '\n' ' ')"
done
done
tail -3 gpurun_out/early2.err
