#!/bin/bash
# sustained clocks / power of the fused kernel and of its ablations (which part of the kernel draws the power cap)
mkdir -p gpurun_out
L=$PWD/time-series-kafka-demo_b200/lib
for name in base ${VARIANTS}; do
  if [ $name = base ]; then lib=$L/libb2cnn.so; else lib=$L/libb2cnn_$name.so; fi
  B2CNN_LIB=$lib timeout -k 10 120 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --e2e-steps 1 --sustained-seconds 2.5 --parity-windows 0 --extra-steps 0 ${EXTRA} > gpurun_out/pw_$name.json 2>> gpurun_out/pw.err
  python - <<PY
import json
d = json.loads(open("gpurun_out/pw_$name.json").read().strip().splitlines()[-1])
s = d.get("sustained", {}); c = s.get("clocks", {})
print("$name: burst ms/step %.4f kernel %.4f | sustained ms/step %.4f  sm_mhz %s  power median %s W max %s W  %s" % (d["ms_per_step"], d["roofline"]["kernel_ms"], s.get("ms_per_step", -1), c.get("sm_mhz"), c.get("power_w_median"), c.get("power_w_max"), c.get("reasons")))
PY
done
tail -2 gpurun_out/pw.err
