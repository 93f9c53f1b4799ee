#!/bin/bash
# soak: the fused-kernel tests repeated (rare-interleaving deadlocks / races in the persistent pipeline would show up as a timeout or a mismatch)
mkdir -p gpurun_out
for i in 1 2 3 4 5 6; do
  timeout -k 10 240 python -m pytest tests/test_gpu_tc.py -q -x --timeout 100 -k "persistent or prefix or logits or nan or geometries or pitch or padded" > gpurun_out/soak_$i.log 2>&1; echo "run $i rc=$? $(tail -1 gpurun_out/soak_$i.log)"
done
timeout -k 10 300 python - <<'PY'
import torch, tskd_b200
m = tskd_b200.B200MyCNN(tskd_b200.ARCH_PRESETS["mycnn5"].with_shape(3, 75000)).to("cuda:0")
x = tskd_b200.synth.make_windows(4096, 3, 75000, "normal", seed=5, dtype=torch.bfloat16, device="cuda:0")
a = tskd_b200.synth.make_ages(4096, seed=5, device="cuda:0")
y0 = m.predict(x, a).clone()
bad = 0
for i in range(400):
    y = m.predict(x, a)
    if i % 50 == 49:
        torch.cuda.synchronize()
        bad += int(not torch.equal(y, y0))
torch.cuda.synchronize()
print("400 headline steps, mismatching checkpoints:", bad)
PY
