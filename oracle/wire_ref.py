"""TEST INFRASTRUCTURE ONLY -- the reference's wire formats as its own code emits them.

``sample_messages``   bin/sendStream.py:39-72: for every sample row ``val`` of ``record.p_signal`` (the selected signals'
                      physical values, NaN where the record has no value) and every signal ``i``:
                      ``producer.produce(topic=signal_list[i], key=record_id[0:7], value=json.dumps([i, val[i]]))``.
``array_message``     bin/processStream.py:126-131: ``key = concat(key, "_", channel)``,
                      ``value = to_json(collect_list(average3))`` -- a JSON array of doubles printed by the JVM
                      (``Double.toString``: shortest digits, scientific notation outside [1e-3, 1e7), "E" exponent).

The decoders under test (csrc/b2cnn_wire.cu) must give, bit for bit, the doubles ``json.loads`` gives for these strings.
"""
from __future__ import annotations

import json
from typing import List, Tuple

import numpy as np


def sample_messages(p_signal: np.ndarray, signal_list, record_id: str, i0: int = 0, i1: int | None = None) -> List[Tuple[str, bytes, bytes]]:
    """(topic, key, value) triples in the order sendStream.py produces them."""
    out = []
    for val in p_signal[i0:i1]:
        for i, _ in enumerate(val):
            jresult = json.dumps([i, float(val[i])])                 # sendStream.py:62 (numpy float64 -> the same repr)
            out.append((signal_list[i], record_id[0:7].encode(), jresult.encode()))
    return out


def java_double_to_string(v: float) -> str:
    """java.lang.Double.toString: what Spark's to_json writes for a DoubleType element."""
    if v != v:
        return '"NaN"'
    if v in (float("inf"), float("-inf")):
        return '"Infinity"' if v > 0 else '"-Infinity"'
    if v == 0:
        return "-0.0" if str(v).startswith("-") else "0.0"
    r = repr(abs(float(v)))
    mant, _, exp = r.partition("e")
    ip, _, fp = mant.partition(".")
    digits = (ip + fp).lstrip("0")
    e10 = (int(exp) if exp else 0) + len(ip.lstrip("0")) if ip.strip("0") else (int(exp) if exp else 0) - (len(fp) - len(fp.lstrip("0")))
    digits = digits.rstrip("0") or "0"
    sign = "-" if v < 0 else ""
    if 1e-3 <= abs(v) < 1e7:
        if e10 <= 0:
            s = "0." + "0" * (-e10) + digits
        elif e10 >= len(digits):
            s = digits + "0" * (e10 - len(digits)) + ".0"
        else:
            s = digits[:e10] + "." + digits[e10:]
        return sign + s
    return sign + digits[0] + "." + (digits[1:] or "0") + "E" + str(e10 - 1)


def array_message(patient_id: str, channel: int, values) -> Tuple[bytes, bytes]:
    """(key, value) of one ``call-stream`` message."""
    return f"{patient_id}_{channel}".encode(), ("[" + ",".join(java_double_to_string(float(v)) for v in values) + "]").encode()
