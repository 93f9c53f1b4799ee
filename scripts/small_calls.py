#!/usr/bin/env python
"""A handful of production-shape calls for an ncu launch list: model(x[1,10,120]) x 20, predict(x[256,10,120]) x 5,
predict(x[4096,10,120]) x 5, and 20 replays of a CUDA graph of the B = 1 call."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import tskd_b200
dev = torch.device("cuda", 0)
m = tskd_b200.B200MyCNN(tskd_b200.ARCH_PRESETS["mycnn5"]).to(dev)
a1 = torch.tensor([65.0], device=dev)
for P, n in ((1, 20), (256, 5), (4096, 5)):
    x = torch.randn(P, 10, 120, device=dev); a = torch.full((P,), 65.0, device=dev)
    for _ in range(n):
        y = m.predict(x, a)
    torch.cuda.synchronize()
x = torch.randn(1, 10, 120, device=dev)
plan = m.call_plan(x, a1)
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    plan(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=s):
        plan()
torch.cuda.synchronize()
for _ in range(20):
    g.replay()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5000):
    g.replay()
torch.cuda.synchronize()
print("graph replay us/call", (time.perf_counter() - t0) / 5000 * 1e6)
t0 = time.perf_counter()
for _ in range(5000):
    plan()
torch.cuda.synchronize()
print("plan() us/call", (time.perf_counter() - t0) / 5000 * 1e6)
