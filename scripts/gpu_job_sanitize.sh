#!/bin/bash
# compute-sanitizer memcheck over one small invocation of every kernel family (smoke + fp32 stream + prep)
mkdir -p gpurun_out
cat > /tmp/san.py <<'PY'
import sys, numpy as np, torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import __graft_entry__ as g
g.smoke()
import tskd_b200
from tskd_b200 import stream as S
from oracle import mycnn_torch as O
from dataclasses import replace
for kind in ("mycnn5", "mycnn3"):
    oarch = O.stretched(O.ARCHS[kind], 3, 3008)
    ref = O.make_ref(oarch, seed=0)
    arch = replace(tskd_b200.ARCH_PRESETS[kind].with_shape(3, 3008), age_coef=oarch.age_coef)
    m = tskd_b200.B200MyCNN(arch, has_out12=oarch.has_out12).to("cuda:0"); m.load_state_dict(ref.state_dict())
    for dt in (torch.float32, torch.bfloat16):
        x = tskd_b200.synth.make_windows(300, 3, 3008, "normal", seed=3, dtype=dt).to("cuda:0")
        a = tskd_b200.synth.make_ages(300, seed=3).to("cuda:0")
        y = m.predict(x, a); torch.cuda.synchronize()
        want = O.ref_independent(ref, x.float().cpu(), a.cpu()).numpy()
        err = float(np.max(np.abs(y.cpu().numpy() - want) / np.maximum(np.abs(want), 1e-6)))
        print(kind, dt, m.last_path, "rel", err)
        ys = m(x[:5], a[:5]); torch.cuda.synchronize()
# persistent grid: more (window-tile pair, range) items than SMs, several items per CTA
oarch = O.stretched(O.ARCHS["mycnn5"], 3, 1528)
ref = O.make_ref(oarch, seed=0)
arch = replace(tskd_b200.ARCH_PRESETS["mycnn5"].with_shape(3, 1528), age_coef=oarch.age_coef)
m = tskd_b200.B200MyCNN(arch, has_out12=oarch.has_out12, path="tensorcore").to("cuda:0"); m.load_state_dict(ref.state_dict())
x = tskd_b200.synth.make_windows(2600, 3, 1528, "physio", seed=4, dtype=torch.bfloat16).to("cuda:0")
a = tskd_b200.synth.make_ages(2600, seed=4).to("cuda:0")
y = m.predict(x, a); torch.cuda.synchronize()
want = O.ref_independent(ref, x[:64].float().cpu(), a[:64].cpu()).numpy()
print("persistent", m.last_path, "rel", float(np.max(np.abs(y[:64].cpu().numpy() - want) / np.maximum(np.abs(want), 1e-6))))
# one training step (row f4)
from tskd_b200.trainer import B200Trainer
mt = tskd_b200.B200MyCNN(tskd_b200.ARCH_PRESETS["mycnn5"]).to("cuda:0")
tr = B200Trainer(mt, dropout=0.1)
xt = torch.randn(24, 10, 120, device="cuda:0"); at = torch.full((24,), 60.0, device="cuda:0"); yt = (torch.rand(24, device="cuda:0") > 0.5).float()
print("train loss", float(tr.step(xt, at, yt)), float(tr.step(xt, at, yt)))
from conftest import load_golden
gd, _ = load_golden("p000194_replay.npz")
rec = S.NumericsRecord(tuple(str(n) for n in gd["names"]), gd["gains"], gd["baselines"], float(gd["fs"]), gd["raw"])
xg, t0 = S.assemble_windows_gpu(rec, "cuda:0"); torch.cuda.synchronize()
print("prep", tuple(xg.shape))
PY
for tool in ${TOOLS:-memcheck racecheck synccheck}; do
timeout -k 10 900 compute-sanitizer --tool $tool --print-limit 20 python /tmp/san.py > gpurun_out/san_$tool.log 2>&1; echo "$tool rc=$?"
grep -c "Invalid\|out of bounds\|misaligned\|hazard\|Barrier error" gpurun_out/san_$tool.log; tail -4 gpurun_out/san_$tool.log
done
