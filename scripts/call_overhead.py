#!/usr/bin/env python
"""Where the ~20 us of one production call model(x[1,10,120], age[1]) go (device-resident tensors): cProfile of the
Python wrapper + timings of the bare ctypes call.  python scripts/call_overhead.py  (needs a GPU)"""
import cProfile, ctypes, json, os, pstats, sys, time, io
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import tskd_b200
from tskd_b200 import capi

dev = torch.device("cuda", 0)
m = tskd_b200.B200MyCNN(tskd_b200.ARCH_PRESETS["mycnn5"]).to(dev)
x = torch.randn(1, 10, 120, device=dev); a = torch.tensor([65.0], device=dev)
for _ in range(50):
    m(x, a)
torch.cuda.synchronize()


def timeit(fn, n=5000):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e6


res = {"model(x, age) us": timeit(lambda: m(x, a)), "predict(x, age) us": timeit(lambda: m.predict(x, a))}
lib, h = m._ensure_handle()
out = torch.empty(1, device=dev); ws = m._workspace(lib, h, 1, 0, 0, dev)
st = torch.cuda.current_stream().cuda_stream
xp, ap, op, wp, wn = x.data_ptr(), a.data_ptr(), out.data_ptr(), ws.data_ptr(), ws.numel()
res["bare ctypes b2cnn_forward us"] = timeit(lambda: lib.b2cnn_forward(h, xp, 0, 1, ap, 1, 0, 0, op, wp, wn, st))
res["torch.empty(1) us"] = timeit(lambda: torch.empty(1, dtype=torch.float32, device=dev))
res["current_stream().cuda_stream us"] = timeit(lambda: torch.cuda.current_stream().cuda_stream)
res["_weights_version us"] = timeit(m._weights_version)
res["x.data_ptr() us"] = timeit(x.data_ptr)
if hasattr(m, "call_plan"):
    plan = m.call_plan(x, a)
    res["call_plan()() us"] = timeit(plan)
print(json.dumps(res))
pr = cProfile.Profile(); pr.enable()
for _ in range(3000):
    m(x, a)
pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(14)
print(s.getvalue()[:3000])
