#!/usr/bin/env python
"""GPU-box diagnostic for the tcgen05 front end: unit-tap weights decode which (tap, channel,
position, window) the tensor-core path gets wrong.  Prints compact error maps."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import tskd_b200
from oracle import mycnn_torch as O

DEV = "cuda:0"
C, W, B = 3, 1528, 128
oarch = O.stretched(O.ARCH_MYCNN5, C, W)
torch.manual_seed(0)
x = torch.randn(B, C, W).to(torch.bfloat16)


def run(sd, label):
    ref = O.RefMyCNN(oarch); ref.load_state_dict(sd); ref.eval()
    out = {}
    for path in ("generic", "tensorcore"):
        m = tskd_b200.B200MyCNN(tskd_b200.ARCH_PRESETS["mycnn5"].with_shape(C, W), path=path).to(DEV)
        m.load_state_dict(sd)
        try:
            out[path] = m.features(x.to(DEV)).cpu().numpy()
        except Exception as e:
            print(label, path, "FAILED:", e); return
    fw = O.ref_features(ref, x.float()).numpy()
    for path in out:
        err = np.abs(out[path] - fw)
        bad = err > 1e-4
        print(f"{label:28s} {path:10s} max_err={err.max():.3e} bad={bad.mean()*100:6.2f}%", end="")
        if bad.any() and path == "tensorcore":
            pos_bad = bad.mean(axis=0); win_bad = bad.mean(axis=1)
            print(f"  first_bad_pos={int(np.argmax(pos_bad>0))} bad_pos_mod14={sorted(set(np.nonzero(pos_bad>0)[0]%14))[:14]}"
                  f" bad_win_mod8={sorted(set(np.nonzero(win_bad>0)[0]%8))} nan={np.isnan(out[path]).mean():.3f}", end="")
        print()
    return out, fw


base = O.make_ref(oarch, seed=0).state_dict()
run(base, "random weights")
for c0 in range(C):
    for k0 in (0, 1, 7, 8, 9):
        sd = {k: v.clone() for k, v in base.items()}
        sd["conv1.weight"].zero_(); sd["conv1.bias"].zero_()
        sd["conv1.weight"][0, c0, k0] = 1.0
        sd["conv2.weight"].zero_(); sd["conv2.bias"].zero_(); sd["conv2.weight"][0, 0, 0] = 1.0
        run(sd, f"unit tap c={c0} k={k0}")
