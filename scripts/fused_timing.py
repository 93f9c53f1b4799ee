#!/usr/bin/env python
"""Where the fused kernel's warps wait: runs the headline shape through a library built with -DB2CNN_TIMING
(B2CNN_LIB=.../libb2cnn_timing.so) and prints the wait counters as shares of each role's loop time."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import tskd_b200
from tskd_b200 import capi
dev = torch.device("cuda", 0)
B, C, W = int(os.environ.get("B", 4096)), 3, int(os.environ.get("W", 75000))
arch = tskd_b200.ARCH_PRESETS["mycnn5"].with_shape(C, W)
m = tskd_b200.B200MyCNN(arch).to(dev)
x = torch.randn(B, C, W, device=dev, dtype=torch.bfloat16)
age = torch.full((B,), 65.0, device=dev)
lib = capi.load_library()
fn = lib.b2cnn_debug_timing
fn.argtypes = [ctypes.POINTER(ctypes.c_ulonglong)]; fn.restype = ctypes.c_int
buf = (ctypes.c_ulonglong * 16)()
for _ in range(3):
    m.predict(x, age)
torch.cuda.synchronize(); fn(buf)
N = 5
for _ in range(N):
    m.predict(x, age)
torch.cuda.synchronize()
assert fn(buf) == 0
v = list(buf)
names = ["epi loop", "epi TFull spin", "epi smem-stage wait", "epi PEmpty wait", "mma loop", "mma Full (TMA) wait", "mma TEmpty wait",
         "mma proj waits", "prod loop", "prod Empty wait", "epi gate drain", "launches"]
for i, n in enumerate(names):
    print(f"{n:22s} {v[i]:16d}")
e, mm, pr = max(v[0], 1), max(v[4], 1), max(v[8], 1)
print(f"epilogue: TFull {v[1] / e:.3f}  stage {v[2] / e:.3f}  PEmpty {v[3] / e:.3f}  (of loop); gate drain / loop {v[10] / e:.3f}")
print(f"mma:      Full(TMA) {v[5] / mm:.3f}  TEmpty {v[6] / mm:.3f}  proj {v[7] / mm:.3f}")
print(f"producer: Empty {v[9] / pr:.3f}")
nw = max(v[11], 1) * 8 * int(os.environ.get("CTAS", 592))
print(f"per epilogue warp, cycles: entry->loop {v[12] / nw:.0f}  loop {v[0] / nw:.0f}  gate drain {v[10] / nw:.0f}  lifetime {v[13] / nw:.0f}")
if v[14]:
    print(f"CTA 0: {v[14] / max(v[11], 1) / 1e3:.1f} us per launch, SM clock while it ran: {v[15] / v[14] * 1e3:.0f} MHz")
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ev0.record()
for _ in range(N):
    m.predict(x, age)
ev1.record(); torch.cuda.synchronize()
print(f"ms per forward (instrumented build): {ev0.elapsed_time(ev1) / N:.4f}")
