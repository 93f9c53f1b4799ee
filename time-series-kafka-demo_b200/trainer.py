"""One training step on the device.

The reference trains with ``utils.train`` (bin/utils.py:183-227): per batch

    optimizer.zero_grad(); output = model(input, age); loss = criterion(output, target)
    loss.backward(); optimizer.step()

with ``criterion = nn.BCEWithLogitsLoss()`` (bin/utils.py:663) and ``torch.optim.Adam``
(bin/explore_torch.ipynb:3204-3205).  :class:`B200Trainer` is that loop body as ONE C-ABI call
(``b2cnn_train_step``, csrc/b2cnn_train.cu: hand-written forward-with-saved-activations, BPTT over the
batch axis, pooling/conv backward, Adam) on a :class:`B200MyCNN`'s parameters::

    trainer = B200Trainer(model, lr=1e-3)
    for x, age, y in loader:                       # x [B,10,120], age [B], y [B] in {0,1}
        loss = trainer.step(x, age, y)             # == the five reference lines above
    model.predict(...)                             # scores with the updated weights

``mode="sequence"`` (default) is what ``model(input_batch, age)`` computes in the reference: the LSTM scans
the batch axis (bin/models.py:29-30).  Dropout (bin/models.py:15, p = 0.1) uses masks drawn by torch on the
device; torch's own Philox stream cannot be reproduced by another implementation, so parity tests pass the
same explicit masks to both sides.
"""
from __future__ import annotations

import ctypes
from typing import Dict, Optional, Tuple

import torch

from . import capi
from .arch import BLOB_KEYS
from .model import B200MyCNN


class B200Trainer:
    def __init__(self, model: B200MyCNN, lr: float = 1e-3, betas: Tuple[float, float] = (0.9, 0.999), eps: float = 1e-8,
                 mode: str = "sequence", dropout: float = 0.1, seed: int = 0):
        if model.arch.affine or model.arch.act_id != 0:
            raise NotImplementedError("training covers the reference's tanh stack (bin/models.py:23,26) without the affine variant")
        if mode not in ("sequence", "independent"):
            raise ValueError("mode must be 'sequence' or 'independent'")
        if not 0.0 <= float(dropout) < 1.0:
            raise ValueError("dropout must be in [0, 1)")
        dev = model._device()
        if dev.type != "cuda":
            raise RuntimeError("B200Trainer needs the model on a CUDA device (there is no CPU fallback)")
        self.model, self.mode, self.dropout = model, mode, float(dropout)
        self._lib = capi.load_library()
        self._cfg = capi.make_config(model.arch, dev.index if dev.index is not None else torch.cuda.current_device())
        self._opt = capi.Adam(lr, betas[0], betas[1], eps)
        self._params = model.packed_weights().to(dev).contiguous()          # master copy, updated in place by the library
        self._m = torch.zeros_like(self._params)
        self._v = torch.zeros_like(self._params)
        self._grads = torch.zeros_like(self._params)
        self._loss = torch.zeros(1, device=dev)
        self._ws: Optional[torch.Tensor] = None
        self._gen = torch.Generator(device=dev)
        self._gen.manual_seed(seed)
        self.steps = 0
        a = model.arch
        self._p1 = ((a.window - a.k1 + 1) - a.pool_k) // a.pool_s + 1

    # ------------------------------------------------------------------
    def _views(self, flat: torch.Tensor) -> Dict[str, torch.Tensor]:
        sd, out, at = self.model.state_dict(), {}, 0
        for k in BLOB_KEYS:
            n = sd[k].numel()
            out[k] = flat[at:at + n].view(sd[k].shape)
            at += n
        return out

    def grads(self) -> Dict[str, torch.Tensor]:
        """d loss / d parameter of the most recent step, keyed like the reference's state_dict."""
        return self._views(self._grads)

    def draw_masks(self, B: int) -> Tuple[Optional[torch.Tensor], Optional[torch.Tensor]]:
        """The two nn.Dropout masks of bin/models.py:25,28, scaled by 1/(1-p) like torch's dropout."""
        if self.dropout <= 0.0:
            return None, None
        keep = 1.0 - self.dropout
        dev = self._params.device
        m1 = torch.bernoulli(torch.full((B, self.model.arch.c_mid, self._p1), keep, device=dev), generator=self._gen) / keep
        m2 = torch.bernoulli(torch.full((B, self.model.arch.l_out), keep, device=dev), generator=self._gen) / keep
        return m1, m2

    def step(self, x: torch.Tensor, age: torch.Tensor, target: torch.Tensor,
             masks: Optional[Tuple[Optional[torch.Tensor], Optional[torch.Tensor]]] = None, update: bool = True) -> torch.Tensor:
        """zero_grad + forward + BCEWithLogitsLoss + backward + Adam step; returns the batch loss (a 0-d device tensor)."""
        dev = self._params.device
        a = self.model.arch
        if x.dim() != 3 or x.shape[1] != a.in_channels or x.shape[2] != a.window:
            raise RuntimeError(f"expected x of shape [B, {a.in_channels}, {a.window}], got {tuple(x.shape)}")
        B = x.shape[0]
        x = x.to(dev, torch.float32).contiguous()
        age = age.to(dev, torch.float32).reshape(-1).contiguous()
        target = target.to(dev, torch.float32).reshape(-1).contiguous()
        if age.numel() != B or target.numel() != B:
            raise RuntimeError("age and target must have one entry per window")
        m1, m2 = masks if masks is not None else self.draw_masks(B)
        if m1 is not None:
            m1 = m1.to(dev, torch.float32).contiguous()
            if tuple(m1.shape) != (B, a.c_mid, self._p1):
                raise RuntimeError(f"dropout mask 1 must be {(B, a.c_mid, self._p1)}, got {tuple(m1.shape)}")
        if m2 is not None:
            m2 = m2.to(dev, torch.float32).contiguous()
            if tuple(m2.shape) != (B, a.l_out):
                raise RuntimeError(f"dropout mask 2 must be {(B, a.l_out)}, got {tuple(m2.shape)}")
        need = self._lib.b2cnn_train_workspace_bytes(ctypes.byref(self._cfg), B)
        if need < 0:
            capi.check(capi.EINVAL, "b2cnn_train_workspace_bytes")
        if self._ws is None or self._ws.numel() < need:
            self._ws = torch.empty(need, dtype=torch.uint8, device=dev)
        ptr = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None else None
        st = torch.cuda.current_stream(dev).cuda_stream
        rc = self._lib.b2cnn_train_step(ctypes.byref(self._cfg), ptr(self._params), ptr(self._m), ptr(self._v), ptr(self._grads),
                                        self.steps + 1, ctypes.byref(self._opt), 1 if update else 0, ptr(x), B, ptr(age), ptr(target),
                                        capi.MODE_SEQUENCE if self.mode == "sequence" else capi.MODE_INDEPENDENT, ptr(m1), ptr(m2),
                                        ptr(self._loss), ptr(self._ws), need, ctypes.c_void_p(st))
        capi.check(rc, "b2cnn_train_step")
        if update:
            self.steps += 1
            with torch.no_grad():                        # the inference kernels read the module's parameters
                sd = self.model.state_dict()
                for k, v in self._views(self._params).items():
                    sd[k].copy_(v)
            self.model.sync_weights()
        return self._loss[0].clone()
