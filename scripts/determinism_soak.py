#!/usr/bin/env python
"""Every one of N headline steps must reproduce the first step bit-for-bit (a race in the streaming kernel's pipelines
would show up as a rare mismatch).  Prints the number of mismatching steps and the largest deviation seen."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import tskd_b200
N = int(os.environ.get("STEPS", 3000))
B = int(os.environ.get("B", 4096))
m = tskd_b200.B200MyCNN(tskd_b200.ARCH_PRESETS["mycnn5"].with_shape(3, 75000)).to("cuda:0")
x = tskd_b200.synth.make_windows(B, 3, 75000, "normal", seed=1234, dtype=torch.bfloat16, device="cuda:0")
a = tskd_b200.synth.make_ages(B, seed=1234, device="cuda:0")
y0 = m.predict(x, a).clone()
bad = torch.zeros((), device="cuda:0", dtype=torch.int64)
worst = torch.zeros((), device="cuda:0")
nwin = torch.zeros((), device="cuda:0", dtype=torch.int64)
for i in range(N):
    y = m.predict(x, a)
    d = (y - y0).abs()
    bad += (d > 0).any()
    nwin += (d > 0).sum()
    worst = torch.maximum(worst, d.max())
torch.cuda.synchronize()
print(f"{N} steps of [{B},3,75000]: {int(bad)} steps differ from the first ({int(nwin)} window results in total), largest |difference| {float(worst):.3e}, path {m.last_path}")
