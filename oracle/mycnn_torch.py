"""TEST INFRASTRUCTURE ONLY -- PyTorch-CPU restatement of the reference forward pass.

The reference's arithmetic for the hot path lives in a third-party dependency, PyTorch
(ATen / oneDNN CPU kernels): the reference neither vendors nor pins it (no torch line in
``requirements.txt``; the install is commented out at ``Dockerfile:14``), so the oracle
version is this image's ``torch 2.11.0``.  The reference's own code for the path is the 35
lines of ``bin/models.py``; this module restates them with the hard-coded sizes turned into
parameters so the same layer stack can be re-instantiated for the synthetic shapes in
``BASELINE.json`` and for the older checkpoint revisions.

Parity pin: ``tests/golden/mycnn5_xtestinput.npz`` holds the outputs of the UNMODIFIED
reference (``bin/models.py`` + ``model/MyCNN5.pth`` + ``explore_output/X.TESTINPUT``) and
``tests/test_oracle.py`` checks this restatement against them bit-for-bit, including the
repo's one known-answer value ``0.5668570399284363`` (``bin/explore_torch.ipynb:4271``).
For MyCNN2/3/4 the age coefficient at save time is unknown -> "parity unpinned" for those
(see DESIGN.md).

All file:line citations are relative to /root/reference.
"""
from __future__ import annotations

from dataclasses import dataclass, replace

import torch
import torch.nn as nn


# --------------------------------------------------------------------------------------
# Architecture description
# --------------------------------------------------------------------------------------
@dataclass(frozen=True)
class RefArch:
    """Hyper-parameters that ``bin/models.py:6-20`` hard-codes."""

    in_channels: int = 10      # models.py:10
    k1: int = 10               # models.py:10  conv1 kernel_size
    c_mid: int = 4             # models.py:10  conv1 out_channels
    k2: int = 5                # models.py:11  conv2 kernel_size
    pool_k: int = 3            # models.py:12
    pool_s: int = 2            # models.py:12
    hidden: int = 16           # models.py:16
    layers: int = 2            # models.py:16
    window: int = 120          # config.cfg:23 WINDOWSIZE
    age_coef: float = 1e-8     # models.py:32
    dropout: float = 0.1       # models.py:15
    has_out12: bool = True     # models.py:13,18 (out1/out2: constructed, never used)

    @property
    def l1(self) -> int:       # conv1 output length (valid, stride 1)
        return self.window - self.k1 + 1

    @property
    def p1(self) -> int:       # MaxPool1d floor mode, no padding
        return (self.l1 - self.pool_k) // self.pool_s + 1

    @property
    def l2(self) -> int:
        return self.p1 - self.k2 + 1

    @property
    def l_out(self) -> int:    # == MAGICNUM (models.py:8) when the view is row-per-window
        return (self.l2 - self.pool_k) // self.pool_s + 1


# MyCNN5 == bin/models.py as shipped.  MyCNN2/3/4 == the older revision in
# bin/explore_torch copy.ipynb:189-277 (k1=5, pool(2,2), dropout .5, view(-1,27),
# age coefficient 1e-4 at :259).
ARCH_MYCNN5 = RefArch()
ARCH_MYCNN4 = RefArch(in_channels=10, k1=5, pool_k=2, pool_s=2, age_coef=1e-4,
                      dropout=0.5, has_out12=False)
ARCH_MYCNN3 = replace(ARCH_MYCNN4, in_channels=7)
ARCH_MYCNN2 = ARCH_MYCNN3
ARCHS = {"mycnn5": ARCH_MYCNN5, "mycnn4": ARCH_MYCNN4, "mycnn3": ARCH_MYCNN3,
         "mycnn2": ARCH_MYCNN2}


def stretched(arch: RefArch, in_channels: int, window: int) -> RefArch:
    """The BASELINE.json synthetic shapes: same layer stack, other C / W."""
    return replace(arch, in_channels=in_channels, window=window)


# --------------------------------------------------------------------------------------
# The module (bin/models.py:5-36, parameterised)
# --------------------------------------------------------------------------------------
class RefMyCNN(nn.Module):
    def __init__(self, arch: RefArch = ARCH_MYCNN5):
        super().__init__()
        self.arch = arch
        self.MAGICNUM = arch.l_out                                    # models.py:8
        self.conv1 = nn.Conv1d(arch.in_channels, arch.c_mid, arch.k1)  # models.py:10
        self.conv2 = nn.Conv1d(arch.c_mid, 1, arch.k2)                 # models.py:11
        self.pool = nn.MaxPool1d(arch.pool_k, arch.pool_s)             # models.py:12
        if arch.has_out12:
            self.out1 = nn.Linear(567, 1)                              # models.py:13 (unused)
        self.dropout = nn.Dropout(arch.dropout)                        # models.py:15
        self.lstm = nn.LSTM(self.MAGICNUM, arch.hidden, arch.layers)   # models.py:16
        self.out = nn.Linear(arch.hidden, 1)                           # models.py:17
        if arch.has_out12:
            self.out2 = nn.Linear(arch.hidden, 1)                      # models.py:18 (unused)
        self.age_fn = nn.Linear(1, 1)                                  # models.py:20 (unused)

    def features(self, x):
        x = torch.tanh(self.conv1(x))       # models.py:23
        x = self.pool(x)                    # models.py:24
        x = self.dropout(x)                 # models.py:25
        x = torch.tanh(self.conv2(x))       # models.py:26
        x = self.pool(x)                    # models.py:27
        x = self.dropout(x)                 # models.py:28
        return x.view(-1, self.MAGICNUM)    # models.py:29

    def forward(self, x, age):
        x = self.features(x)
        x, _ = self.lstm(x)                 # models.py:30  (2-D input => unbatched sequence)
        x = self.out(x)                     # models.py:31
        age_scale = torch.relu(age.unsqueeze(1) * self.arch.age_coef + 1)  # models.py:32
        x = x * age_scale                   # models.py:33
        return x.squeeze(1)                 # models.py:34


def make_ref(arch: RefArch, seed: int = 0) -> RefMyCNN:
    """Seeded default-initialised module in eval mode (predictStream.py:37)."""
    torch.manual_seed(seed)
    m = RefMyCNN(arch)
    m.eval()
    return m


# --------------------------------------------------------------------------------------
# The two batch semantics
# --------------------------------------------------------------------------------------
@torch.no_grad()
def ref_sequence(m: RefMyCNN, x: torch.Tensor, age: torch.Tensor) -> torch.Tensor:
    """``model(x_batch, age)`` exactly as utils.py:204,249,682 call it: the LSTM scans the
    batch axis (bin/models.py:29-30)."""
    return m(x.float(), age.float())


@torch.no_grad()
def ref_independent_loop(m: RefMyCNN, x: torch.Tensor, age: torch.Tensor) -> torch.Tensor:
    """predictStream.py:154-157 semantics: one ``model(x[i:i+1], age[i:i+1])`` per window."""
    x = x.float()
    age = age.float()
    return torch.cat([m(x[i:i + 1], age[i:i + 1]) for i in range(x.shape[0])])


@torch.no_grad()
def ref_independent(m: RefMyCNN, x: torch.Tensor, age: torch.Tensor) -> torch.Tensor:
    """Same result as :func:`ref_independent_loop` but batched: the LSTM is fed a
    ``[seq=1, batch=B, L]`` tensor so every window starts from the zero state."""
    x = x.float()
    age = age.float()
    f = m.features(x)
    h, _ = m.lstm(f.unsqueeze(0))
    y = m.out(h.squeeze(0))
    y = y * torch.relu(age.unsqueeze(1) * m.arch.age_coef + 1)
    return y.squeeze(1)


@torch.no_grad()
def ref_features(m: RefMyCNN, x: torch.Tensor) -> torch.Tensor:
    return m.features(x.float())


def post_process(logit: torch.Tensor):
    """predictStream.py:160-162 / utils.py:685-687."""
    y = torch.sigmoid(logit).detach().to("cpu")
    return y.round().long().numpy().tolist(), y.numpy().tolist()
