"""TEST INFRASTRUCTURE ONLY -- ctypes binding of oracle/mycnn_ref.c (plain-C restatement).

``build()`` compiles ``oracle/mycnn_ref.c`` into ``oracle/_build/libmycnn_ref.so`` with gcc.
Used by tests/ as an independent fp64 ground truth next to the PyTorch-CPU oracle.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "mycnn_ref.c")
_OUT_DIR = os.path.join(_HERE, "_build")
_LIB = os.path.join(_OUT_DIR, "libmycnn_ref.so")


class _Arch(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int32) for n in
                ("in_channels", "k1", "c_mid", "k2", "pool_k", "pool_s", "hidden", "window",
                 "act", "reserved")] + [("age_coef", ctypes.c_double)]


def build(force: bool = False) -> str:
    if (not force and os.path.exists(_LIB)
            and os.path.getmtime(_LIB) >= os.path.getmtime(_SRC)):
        return _LIB
    os.makedirs(_OUT_DIR, exist_ok=True)
    subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-o", _LIB, _SRC, "-lm"])
    return _LIB


_lib = None


def _load():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        _lib.mycnn_ref_l_out.restype = ctypes.c_int
        _lib.mycnn_ref_weight_count.restype = ctypes.c_int64
        for name in ("mycnn_ref_forward_f64", "mycnn_ref_forward_f32"):
            getattr(_lib, name).restype = ctypes.c_int
    return _lib


def _arch_struct(arch, act: int = 0) -> _Arch:
    return _Arch(arch.in_channels, arch.k1, arch.c_mid, arch.k2, arch.pool_k, arch.pool_s,
                 arch.hidden, arch.window, act, 0, float(arch.age_coef))


BLOB_KEYS = ("conv1.weight", "conv1.bias", "conv2.weight", "conv2.bias",
             "lstm.weight_ih_l0", "lstm.weight_hh_l0", "lstm.bias_ih_l0", "lstm.bias_hh_l0",
             "lstm.weight_ih_l1", "lstm.weight_hh_l1", "lstm.bias_ih_l1", "lstm.bias_hh_l1",
             "out.weight", "out.bias")


def pack_blob(state_dict) -> np.ndarray:
    parts = [np.asarray(state_dict[k].detach().cpu().numpy() if hasattr(state_dict[k], "detach")
                        else state_dict[k], dtype=np.float32).ravel() for k in BLOB_KEYS]
    return np.ascontiguousarray(np.concatenate(parts))


def forward(arch, blob: np.ndarray, x: np.ndarray, age: np.ndarray, mode: str = "independent",
            precision: str = "f64", act: int = 0, want_features: bool = False):
    """x: [B, C, W] float32; age: [B] float32.  Returns logits (and features) as float64."""
    lib = _load()
    a = _arch_struct(arch, act)
    x = np.ascontiguousarray(x, dtype=np.float32)
    age = np.ascontiguousarray(age, dtype=np.float32)
    B = x.shape[0]
    assert x.shape[1:] == (arch.in_channels, arch.window)
    assert blob.size == lib.mycnn_ref_weight_count(ctypes.byref(a)), "blob size mismatch"
    L = lib.mycnn_ref_l_out(ctypes.byref(a))
    rdt = np.float64 if precision == "f64" else np.float32
    logits = np.empty(B, dtype=rdt)
    feats = np.empty((B, L), dtype=rdt) if want_features else None
    fn = lib.mycnn_ref_forward_f64 if precision == "f64" else lib.mycnn_ref_forward_f32
    rc = fn(ctypes.byref(a), blob.ctypes.data_as(ctypes.c_void_p),
            x.ctypes.data_as(ctypes.c_void_p), age.ctypes.data_as(ctypes.c_void_p),
            ctypes.c_int64(B), ctypes.c_int(0 if mode == "independent" else 1),
            logits.ctypes.data_as(ctypes.c_void_p),
            feats.ctypes.data_as(ctypes.c_void_p) if want_features else None)
    if rc != 0:
        raise RuntimeError(f"mycnn_ref_forward failed rc={rc}")
    return (logits.astype(np.float64), feats) if want_features else logits.astype(np.float64)
