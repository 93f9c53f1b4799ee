"""TEST INFRASTRUCTURE ONLY -- the pandas form of the preprocessing step, the pin of oracle/stream_np.py.

Two pipelines over one signal of a numerics record (samples at t_i = i / fs seconds, NaN = missing):

``grid_notebook``   the reference's OFFLINE preprocessing, verbatim pandas calls of
                    bin/explore_torch.ipynb:402,405::

                        record_df = record_df.resample('5S').first()
                        record_df = record_df.rolling('3min').mean()

                    followed by the streaming job's fill (bin/processStream.py:62-123: ffill, bfill, fillna(0)).
                    ``first()`` keeps ONE sample per 5-second bin, so this equals the streaming job's average
                    only when the record has at most one sample per bin (fs <= 0.2 Hz: every MIMIC numerics record).

``grid_spark``      the streaming job's aggregate for ANY sampling rate, written with pandas:
                    bin/processStream.py:196-208 ``avg(value)`` over ``window(timestamp, "180 s", "5 s")`` =
                    (sum of all valid samples) / (their count) over the 36 five-second bins (tau-180, tau]
                    = sample times in [tau - 175, tau + 5)  -- Spark's half-open [start, start + 180) with
                    start = tau - 175.

pandas labels a rolling window by its RIGHT edge and closes it on the right, ``(t - 180, t]``; Spark keys a window by
its START and closes it on the left, ``[s, s + 180)``.  On the 5-second bin lattice both select bins
t-175, ..., t  (s = t - 175): tests/test_stream_oracle.py checks the boundary bins explicitly.
"""
from __future__ import annotations

import numpy as np
import pandas as pd


def _series(samples: np.ndarray, fs: float) -> pd.Series:
    period_ns = int(round(1e9 / fs))
    idx = pd.to_datetime(np.arange(samples.shape[0], dtype=np.int64) * period_ns, unit="ns")
    return pd.Series(np.asarray(samples, dtype=np.float64), index=idx)


def _fill(g: pd.Series) -> np.ndarray:
    return g.ffill().bfill().fillna(0.0).to_numpy()


def grid_notebook(samples: np.ndarray, fs: float, fill: bool = True) -> np.ndarray:
    g = _series(samples, fs).resample("5s").first().rolling("3min").mean()
    return _fill(g) if fill else g.to_numpy()


def grid_spark(samples: np.ndarray, fs: float, fill: bool = True) -> np.ndarray:
    r = _series(samples, fs).resample("5s")
    num = r.sum().rolling("3min").sum()          # NaN samples are skipped by sum() / count()
    cnt = r.count().rolling("3min").sum()
    g = (num / cnt).where(cnt > 0)
    return _fill(g) if fill else g.to_numpy()
