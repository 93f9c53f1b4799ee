"""Synthetic waveform windows for benchmarks and parity tests (SURVEY.md section 8 d).

Value distributions:
  "normal"  (A) N(0,1): does not saturate tanh -- the strictest case for parity;
  "physio"  (B) vital-sign-like: per-channel baseline in {80, 16, 97, ...} + N(0, 2), 10 % of
            the windows with one all-zero channel (a missing signal, bin/predictStream.py:131),
            clipped to [0, 200] -- saturates conv1's tanh the way the shipped X.TESTINPUT does;
  "edge"    (C) "normal" with NaN / +-inf / denormal samples injected in a few windows.
Ages ~ U(15, 80): the notebook clamps ages to that range (bin/explore_torch.ipynb:1923-1928).
"""
from __future__ import annotations

import torch

_BASELINES = (80.0, 16.0, 97.0, 2.0, 97.0, 8.0, 0.5, 85.0, 60.0, 120.0)   # HR, RESP, PULSE, ...


def make_windows(B: int, C: int, W: int, dist: str = "normal", seed: int = 1234,
                 dtype: torch.dtype = torch.float32, device="cpu", chunk: int = 256) -> torch.Tensor:
    """[B, C, W] windows.  Generated chunk-wise on `device` so the 1.8 GB headline batch never
    needs a second full-size temporary; deterministic for a given (seed, device type)."""
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    out = torch.empty(B, C, W, dtype=dtype, device=device)
    base = torch.tensor([_BASELINES[c % len(_BASELINES)] for c in range(C)], device=device).view(1, C, 1)
    for b0 in range(0, B, chunk):
        nb = min(chunk, B - b0)
        x = torch.randn(nb, C, W, generator=g, device=device)
        if dist == "physio":
            x = x * 2.0 + base
            drop = torch.rand(nb, generator=g, device=device) < 0.10
            ch = torch.randint(0, C, (nb,), generator=g, device=device)
            idx = torch.nonzero(drop).flatten()
            x[idx, ch[idx]] = 0.0
            x.clamp_(0.0, 200.0)
        elif dist == "edge":
            if nb >= 4:
                x[0, 0, W // 3] = float("nan")
                x[1, C - 1, W // 2] = float("inf")
                x[2, 0, 5] = float("-inf")
                x[3, 0, 7] = 1e-41
        elif dist != "normal":
            raise ValueError(f"unknown distribution {dist!r}")
        out[b0:b0 + nb] = x.to(dtype)
    return out


def make_ages(B: int, seed: int = 1234, device="cpu") -> torch.Tensor:
    g = torch.Generator(device=device)
    g.manual_seed(seed + 7)
    return torch.rand(B, generator=g, device=device) * 65.0 + 15.0
