#!/usr/bin/env python
"""When the grids of the early-head chain start and end (library built with -DB2CNN_EARLY_TIMING,
B2CNN_LIB=.../libb2cnn_etiming.so): %globaltimer stamps, relative to the fused kernel's first instruction.
STEPS=n back-to-back steps per line: the stamps are then min (entries) / max (exits) over the n steps, i.e. the first
step's entries and the last step's exits."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import tskd_b200
from tskd_b200 import capi
dev = torch.device("cuda", 0)
B, C, W = int(os.environ.get("B", 4096)), 3, int(os.environ.get("W", 75000))
m = tskd_b200.B200MyCNN(tskd_b200.ARCH_PRESETS["mycnn5"].with_shape(C, W)).to(dev)
if os.environ.get("B2CNN_EARLY_HEAD") is not None:
    m.set_option("early_head", int(os.environ["B2CNN_EARLY_HEAD"]))
x = torch.randn(B, C, W, device=dev, dtype=torch.bfloat16)
age = torch.full((B,), 65.0, device=dev)
lib = capi.load_library()
fn = lib.b2cnn_debug_early
fn.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_ulonglong), ctypes.c_int]; fn.restype = ctypes.c_int
buf = (ctypes.c_ulonglong * 7)()
for _ in range(5):
    m.predict(x, age)
torch.cuda.synchronize()
for n in (1, 1, 1, 2, 2, 3, 3):
    assert fn(m._handle, buf, 1) == 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        y = m.predict(x, age)
    e1.record()
    torch.cuda.synchronize()
    assert fn(m._handle, buf, 0) == 0
    v = list(buf); t0 = v[4]
    r = lambda i: (v[i] - t0) / 1e3
    print(f"{n} step(s), {e0.elapsed_time(e1) * 1e3:8.1f} us by events: fused first instr 0, set-up done {r(0):6.1f}, LAST exit {r(1):8.1f} | head first entry {r(2):6.1f}, last exit {r(3):8.1f} | "
          f"finish first entry {r(5):8.1f}, last exit {r(6):8.1f}")
