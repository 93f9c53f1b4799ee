"""Loading the reference's checkpoints (bin/predictStream.py:36: ``torch.load(cfg['MODELPATH'])``).

``model/MyCNN{2,3,4,5}.pth`` are legacy (non-zip) full-module pickles of a class named
``__main__.MyCNN`` (written by ``torch.save(model, path, _use_new_zipfile_serialization=False)``,
bin/explore_torch.ipynb:995,3234,3285).  Unpickling them needs *a* class of that name; the
reference satisfies it with ``from models import MyCNN`` (bin/predictStream.py:8).  Here a
restricted unpickler maps that one global to an inert ``nn.Module`` stub and lets through an
exact (module, name) allowlist of the globals the audited pickles contain -- nothing else, so loading does not depend on
``bin/models.py`` and the architecture is read from the unpickled sub-modules, not assumed.
"""
from __future__ import annotations

import pickle
import types
import warnings
from typing import Dict, Tuple

import torch
import torch.nn as nn

from .arch import ArchConfig


class PickledMyCNN(nn.Module):
    """Stand-in for ``__main__.MyCNN``; holds whatever sub-modules the pickle carries."""

    def forward(self, *a, **k):  # pragma: no cover - never executed
        raise RuntimeError("PickledMyCNN is a container for unpickled weights; "
                           "wrap it with B200MyCNN.from_reference(...)")


# Exact (module, name) allowlist: the GLOBAL opcodes of the four audited checkpoints (pickletools over
# model/MyCNN{2,3,4,5}.pth) plus the storage classes a plain state_dict file may name.  Anything else --
# builtins.eval, os.system, getattr ... -- raises UnpicklingError instead of being imported.
_ALLOWED_GLOBALS = {
    ("collections", "OrderedDict"),
    ("torch._utils", "_rebuild_tensor_v2"),
    ("torch._utils", "_rebuild_parameter"),
    ("torch", "FloatStorage"), ("torch", "DoubleStorage"), ("torch", "HalfStorage"), ("torch", "BFloat16Storage"),
    ("torch", "LongStorage"), ("torch", "IntStorage"),
    ("torch.nn.modules.conv", "Conv1d"),
    ("torch.nn.modules.pooling", "MaxPool1d"),
    ("torch.nn.modules.dropout", "Dropout"),
    ("torch.nn.modules.linear", "Linear"),
    ("torch.nn.modules.rnn", "LSTM"),
}


class _Unpickler(pickle.Unpickler):
    def find_class(self, module, name):
        if (module, name) == ("__main__", "MyCNN"):
            return PickledMyCNN
        if (module, name) in (("__builtin__", "set"), ("builtins", "set")):
            return set
        if (module, name) in _ALLOWED_GLOBALS:
            return super().find_class(module, name)
        raise pickle.UnpicklingError(f"global {module}.{name} is not allowed in a MyCNN checkpoint")


_pickle_module = types.ModuleType("b2cnn_restricted_pickle")
_pickle_module.Unpickler = _Unpickler
_pickle_module.load = lambda f, **kw: _Unpickler(f, **kw).load()
_pickle_module.__name__ = "pickle"


def load_reference_checkpoint(path: str) -> nn.Module:
    """``torch.load(path)`` for the reference's checkpoints, CPU, without ``bin/models.py``.

    Also accepts new-style ``state_dict`` files (returns the dict unchanged)."""
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")   # SourceChangeWarning for the torch.nn classes
        obj = torch.load(path, map_location="cpu", weights_only=False, pickle_module=_pickle_module)
    return obj


def arch_of_module(m: nn.Module, window: int = 120, age_coef: float | None = None) -> ArchConfig:
    """Architecture from the unpickled sub-modules (conv kernel sizes, pool geometry, LSTM
    input size), exactly what SURVEY.md section 7 step 1 prescribes."""
    def _i(v):
        return int(v[0]) if isinstance(v, (tuple, list)) else int(v)
    k1, pk, ps = _i(m.conv1.kernel_size), _i(m.pool.kernel_size), _i(m.pool.stride)
    if age_coef is None:
        age_coef = 1e-8 if (k1, pk, ps) == (10, 3, 2) else 1e-4
    a = ArchConfig(in_channels=int(m.conv1.in_channels), k1=k1, k2=_i(m.conv2.kernel_size),
                   pool_k=pk, pool_s=ps, window=window, age_coef=age_coef,
                   c_mid=int(m.conv1.out_channels), hidden=int(m.lstm.hidden_size),
                   layers=int(m.lstm.num_layers))
    if a.l_out != int(m.lstm.input_size):
        raise RuntimeError(
            f"checkpoint LSTM input size {int(m.lstm.input_size)} != L_out(window={window})={a.l_out}: "
            "x.view(-1, MAGICNUM) would straddle windows (bin/models.py:29)")
    return a


def split_state_dict(sd) -> Tuple[Dict[str, torch.Tensor], Dict[str, torch.Tensor]]:
    from .arch import BLOB_KEYS
    used = {k: sd[k] for k in BLOB_KEYS}
    inert = {k: v for k, v in sd.items() if k not in used}
    return used, inert
