#!/bin/bash
set -x
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/e_build.log 2>&1
timeout -k 10 900 python -m pytest tests -m gpu -q -x > gpurun_out/e_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/e_pytest.log
tail -12 gpurun_out/e_pytest.log
timeout -k 10 300 python - > gpurun_out/e_latency.log 2>&1 <<'PY'
import time, torch, json, sys, os
sys.path.insert(0, os.getcwd())
import tskd_b200
dev = "cuda:0"
m = tskd_b200.B200MyCNN(tskd_b200.ARCH_PRESETS["mycnn5"]).to(dev)
x = torch.randn(1, 10, 120, device=dev); a = torch.tensor([65.0], device=dev)
for small in (1, 0):
    m.predict(x, a); m.set_option("small_kernel", small)
    for _ in range(50): m(x, a)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(3000): m(x, a)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 3000
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(200): m(x, a)
    e1.record(); torch.cuda.synchronize()
    print(json.dumps({"small_kernel": small, "us_per_call_wall": dt * 1e6, "launches": m.gpu_launches,
                      "us_per_call_device_span": e0.elapsed_time(e1) * 1e3 / 200}))
xh = torch.randn(1, 10, 120).double().numpy()
m.set_option("small_kernel", 1)
t0 = time.perf_counter()
for _ in range(1000):
    y = torch.sigmoid(m(torch.from_numpy(xh).float(), torch.tensor([65.0])))
dt = (time.perf_counter() - t0) / 1000
print(json.dumps({"host_tensors_in_out_us_per_call": dt * 1e6}))
PY
cat gpurun_out/e_latency.log
