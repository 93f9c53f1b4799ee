"""Rows f2 + f1 of SURVEY.md section 8 -- the two steps in front of the model call -- and the streaming replay
(BASELINE.json configs[4]) at the model boundary.

Chain of evidence:
  pandas (the reference notebook's own resample / rolling calls, oracle/stream_pandas.py; fixtures written by
  tests/golden/make_golden.py)  ->  pins  oracle/stream_np.py (numpy restatement with Spark's window edges)
  ->  checks  csrc/b2cnn_prep.cu: b2cnn_prep_windows (whole record) and b2cnn_ring_* (trigger by trigger),
  and the scores against the UNMODIFIED reference model run on the pandas-built windows of record p000194.
"""
import numpy as np
import pytest
import torch

import tskd_b200
from tskd_b200 import stream as S
from conftest import load_golden, rel_err
from oracle import mycnn_torch as O
from oracle import stream_np as N
from oracle import stream_pandas as P

SYNTH_RECORDS = [(1, 1625, 1 / 60, {}), (2, 4000, 1.0, {"p_missing": 0.5}), (3, 900, 1.0, {"lead_gap": 400}),
                 (4, 2500, 0.2, {"dead": 1}), (5, 700, 1.0, {"p_missing": 0.0}), (6, 50000, 1.0, {"p_missing": 0.9}),
                 (7, 640, 1.0, {"n_sig": 2}), (8, 3000, 0.1, {"p_missing": 0.4})]      # == tests/golden/make_golden.py


def _record(g):
    return S.NumericsRecord(tuple(str(n) for n in g["names"]), g["gains"], g["baselines"], float(g["fs"]), g["raw"])


def _synthetic_record(seed, n, fs, n_sig=7, p_missing=0.2, lead_gap=0, dead=None):
    rng = np.random.default_rng(seed)
    names = ["HR", "PULSE", "junk A", "RESP", "SpO2", "NBPSys", "CVP"][:n_sig]
    raw = rng.integers(-500, 3000, size=(n, n_sig)).astype(np.int16)
    raw[rng.random((n, n_sig)) < p_missing] = -32768
    if lead_gap:
        raw[:lead_gap, 0] = -32768
    if dead is not None:
        raw[:, dead] = -32768
    gains = rng.choice([1.0, 10.0, 12.5], size=n_sig)
    bases = rng.integers(-5, 5, size=n_sig).astype(np.float64)
    return S.NumericsRecord(tuple(names), gains, bases, fs, raw)


def _np_windows(rec):
    return N.assemble_windows(rec, S.selected_signals(rec))


# ------------------------------------------------------------------------------------------ CPU: reader + oracle pins
def test_record_decoding_and_channel_selection():
    g, _ = load_golden("p000194_replay.npz")
    rec = _record(g)
    assert rec.raw.shape == (1625, 7) and abs(rec.fs - 1 / 60) < 1e-9
    # 16-bit column checksums of the header (p000194-2112-05-23-14-34n.hea:2-8)
    csum = [int(np.int16(rec.raw[:, i].astype(np.int64).sum() & 0xffff)) for i in range(7)]
    assert csum == [3240, 20492, 29088, 10310, -27206, -29717, -28780]
    # only the names config.cfg:23 lists are selected; "NBPSys" != "NBP Sys" (sendStream.py:46)
    assert [rec.names[i] for i in S.selected_signals(rec)] == ["HR", "PULSE", "RESP", "SpO2"]
    p = rec.physical
    assert np.isnan(p[rec.raw == -32768]).all() and np.nanmax(p[:, 0]) < 300


def test_numpy_restatement_is_pinned_by_pandas_fixture_on_shipped_record():
    """p000194: the committed grids were produced by pandas (resample('5S').first() -> rolling('3min').mean() ->
    ffill/bfill/fillna(0), bin/explore_torch.ipynb:402,405 + bin/processStream.py:62-123) in make_golden.py."""
    g, _ = load_golden("p000194_replay.npz")
    rec = _record(g)
    sel = S.selected_signals(rec)
    grids = N.grids_of_record(rec, sel)
    assert grids.shape == g["grids"].shape == (4, 19489)
    assert np.abs(grids - g["grids"]).max() <= 1e-10 * np.abs(g["grids"]).max()
    # unfilled: the NaN pattern (which grid points have no valid sample in their window) is identical
    phys = rec.physical
    unfilled = np.stack([N.smooth_to_grid(phys[:, s], rec.fs, fill=False) for s in sel])
    assert np.array_equal(np.isnan(unfilled), np.isnan(g["grids_unfilled"]))
    x, t0 = N.windows_from_grids(grids)
    assert x.shape == (1615, 10, 120) and x.dtype == np.float64
    assert np.abs(x[0] - g["x_first"]).max() <= 1e-10 * 255 and np.abs(x[-1] - g["x_last"]).max() <= 1e-10 * 255
    assert np.array_equal(x[0].astype(np.float32), g["x_first"].astype(np.float32))     # what the model sees (.float())
    assert (x[:, 4:, :] == 0).all()           # the six signals the record lacks (predictStream.py:131)
    assert t0[1] - t0[0] == 60.0 and np.array_equal(t0, g["t0"])
    # consecutive windows overlap by 108 of 120 points (600 s window, 60 s slide)
    assert np.array_equal(x[1, 0, :108], x[0, 0, 12:])


def test_numpy_restatement_matches_live_pandas_and_fixture_on_synthetic_records():
    """Gaps, dead signals, leading gaps, 1/60 .. 1 Hz: numpy == pandas run now == the committed pandas fixture."""
    fx = np.load(__import__("os").path.join(__import__("conftest").GOLDEN, "stream_synth_grids.npz"))
    for seed, n, fs, kw in SYNTH_RECORDS:
        rec = _synthetic_record(seed, n, fs, **kw)
        sel = S.selected_signals(rec)
        phys = rec.physical
        got = N.grids_of_record(rec, sel)
        live = np.stack([P.grid_spark(phys[:, s], rec.fs) for s in sel])
        scale = max(1.0, np.abs(live).max())
        assert got.shape == live.shape == fx[f"grid{seed}"].shape, seed
        assert np.abs(got - live).max() <= 1e-10 * scale, seed
        assert np.abs(got - fx[f"grid{seed}"]).max() <= 1e-10 * scale, seed
        if fs <= 0.2:       # <= one sample per 5-second bin: the notebook's resample().first() pipeline is the same thing
            nb = np.stack([P.grid_notebook(phys[:, s], rec.fs) for s in sel])
            assert np.abs(got - nb).max() <= 1e-10 * scale, seed


def test_window_edge_convention_boundary_samples():
    """Spark windows are [start, start+180), pandas' are (t-180, t]: on the 5-second lattice both put a sample at
    exactly t-180 OUT and one at exactly t IN.  One sample per minute, distinct values so membership is visible."""
    fs = 1 / 60.0
    s = np.array([np.nan, 60.0, 90.0, np.nan, 30.0, 30.0, 30.0])      # samples at t = 0, 60, ..., 360 s
    for grid in (N.smooth_to_grid(s, fs), P.grid_notebook(s, fs), P.grid_spark(s, fs)):
        assert len(grid) == 6 * 12 + 1
        assert grid[0] == 60.0                      # nothing valid yet -> back-filled from the first valid mean
        assert grid[12] == 60.0                     # label 60:  (-120, 60]  = {nan, 60}
        assert grid[24] == 75.0                     # label 120: (-60, 120]  = {nan, 60, 90} -> 75
        assert grid[35] == 75.0 and grid[36] == 75.0    # label 180: (0, 180] = {60, 90, nan}; t=0 is exactly t-180 -> OUT
        assert grid[47] == 75.0                     # label 235: (55, 235]   = {60, 90, nan}
        assert grid[48] == 60.0                     # label 240: (60, 240]   = {90, nan, 30}: the sample at exactly t-180 = 60 left,
        #                                             the sample at exactly t = 240 entered
        assert grid[72] == 30.0
    s2 = np.arange(1, 12, dtype=np.float64)        # 11 samples, one per minute, all valid
    g = N.smooth_to_grid(s2, fs, fill=False)
    # Spark view of the same point: label 240 == windowStart 65 -> [65, 245) holds t = 120, 180, 240 (values 3, 4, 5)
    assert g[48] == 4.0 and g[47] == np.mean([2.0, 3.0, 4.0])
    # off the lattice (1 Hz): a sample at tau+5-1 is IN, at tau+5 OUT; at tau-175 IN, at tau-176 OUT
    s3 = np.zeros(400); s3[180] = 1.0
    g3 = N.smooth_to_grid(s3, 1.0, fill=False)
    assert g3[35] == 0.0 and g3[36] > 0 and g3[71] > 0 and g3[72] == 0.0      # IN for tau in [180-4 .. 180+175] on the lattice
    assert np.allclose(g3, P.grid_spark(s3, 1.0, fill=False), atol=1e-15)
    z = N.smooth_to_grid(np.full(5, np.nan), fs)
    assert (z == 0).all()                          # never observed -> zeros (processStream.py:123 fillna(0))


def test_prep_window_count_matches_host_logic():
    """Host-side arithmetic of the C ABI (no device work): window count of b2cnn_prep_window_count == the oracle's."""
    import ctypes
    from tskd_b200 import capi
    lib = capi.load_library()
    cfg = capi.PrepConfig(S.N_CHANNELS, S.WINDOW_POINTS, S.GRID_S, S.SMOOTH_S, S.STRIDE_S)
    for n, fs in [(1625, 1 / 60), (700, 1.0), (595, 1.0), (596, 1.0), (3000, 0.2), (12, 1 / 60), (2, 1.0), (3000, 0.1)]:
        n_grid = ((n - 1) * N.sample_period_ns(fs)) // (S.GRID_S * N.NS) + 1
        want = len(np.arange(0, n_grid - S.WINDOW_POINTS + 1, S.STRIDE_S // S.GRID_S))
        assert lib.b2cnn_prep_window_count(n, fs, ctypes.byref(cfg)) == want, (n, fs)
    assert lib.b2cnn_prep_window_count(0, 1.0, ctypes.byref(cfg)) < 0
    bad = capi.PrepConfig(10, 120, 5, 180, 62)                    # stride not a multiple of the grid
    assert lib.b2cnn_prep_window_count(1000, 1.0, ctypes.byref(bad)) < 0


def test_prep_c_abi_argument_errors_need_no_gpu():
    """b2cnn_prep_windows validates shapes, pointers and the workspace before any CUDA call."""
    import ctypes
    from tskd_b200 import capi
    lib = capi.load_library()
    cfg = capi.PrepConfig(S.N_CHANNELS, S.WINDOW_POINTS, S.GRID_S, S.SMOOTH_S, S.STRIDE_S)
    sel = np.array([0, 1], dtype=np.int32); g = np.ones(7); b = np.zeros(7)
    args = lambda raw, xo, ws, wsb: (raw, 1000, 7, sel.ctypes.data, 2, g.ctypes.data, b.ctypes.data, 1.0, ctypes.byref(cfg),
                                     xo, 0, None, ws, wsb, None)
    assert lib.b2cnn_prep_windows(*args(None, None, None, 0)) == capi.EINVAL            # null pointers
    assert "null" in capi.last_error()
    need = lib.b2cnn_prep_workspace_bytes(1000, 1.0, 2, ctypes.byref(cfg))
    assert need > 0
    assert lib.b2cnn_prep_windows(*args(0x1000, 0x2000, 0x3000, need - 1)) == capi.ESTATE   # workspace too small
    bad = (0x1000, 1000, 7, sel.ctypes.data, 2, g.ctypes.data, b.ctypes.data, 1.0, ctypes.byref(cfg), 0x2000, 7, None, 0x3000, need, None)
    assert lib.b2cnn_prep_windows(*bad) == capi.EINVAL                                   # dtype


# ------------------------------------------------------------------------------------------ GPU: whole-record form
@pytest.mark.gpu
def test_replay_parity_gpu_vs_reference_scores():
    """configs[4]: record p000194 -> device window assembly -> ONE batched predict(); expected scores = the UNMODIFIED
    reference model on the pandas-built windows (make_golden.py)."""
    g, _ = load_golden("p000194_replay.npz")
    g5, sd = load_golden("mycnn5_xtestinput.npz")
    rec = _record(g)
    model = tskd_b200.B200MyCNN.from_reference(sd).to("cuda:0")
    rows = S.replay(model, rec, subject_id=194, age=65.0)
    assert len(rows) == 1615 and rows[0][0] == 194 and rows[1][1] - rows[0][1] == 60.0
    probs = np.array([r[2] for r in rows])
    assert rel_err(probs, g["probs"]) <= 1e-4                            # vs the unmodified reference
    rows_mb = S.replay(model, rec, subject_id=194, age=65.0, micro_batch=16)   # BATCHSIZE = 16 (config.cfg:26)
    # micro-batches of 16 take the single-launch small-window kernel, the full batch the general path
    assert rel_err(np.array([r[2] for r in rows_mb]), probs) <= 1e-6
    # logits on identical x_arr: GPU vs golden vs oracle per-window loop
    x, _ = _np_windows(rec)
    xt = torch.from_numpy(x).float()
    logit = model.predict(xt.cuda(), 65.0).cpu().numpy()
    assert rel_err(logit, g["logits"]) <= 1e-4
    ref = O.RefMyCNN(O.ARCH_MYCNN5); ref.load_state_dict(sd); ref.eval()
    want = O.ref_independent_loop(ref, xt[:200], torch.full((200,), 65.0)).numpy()
    assert np.array_equal(want, g["logits"][:200]) and rel_err(logit[:200], want) <= 1e-4


@pytest.mark.gpu
def test_replay_accepts_a_model_that_was_never_moved_to_cuda():
    """B200MyCNN.from_reference(...).eval() without .to('cuda') must work through replay() like predict()."""
    g, _ = load_golden("p000194_replay.npz")
    _, sd = load_golden("mycnn5_xtestinput.npz")
    model = tskd_b200.B200MyCNN.from_reference(sd).eval()
    rows = S.replay(model, _record(g), subject_id=194)
    assert len(rows) == 1615
    short = _synthetic_record(5, 100, 1.0)                             # too short for one window: no rows, no error
    assert S.replay(model, short, subject_id=1) == [] and S.replay(model, short, subject_id=1, micro_batch=4) == []


@pytest.mark.gpu
def test_gpu_window_assembly_matches_pandas_fixture_on_shipped_record():
    g, _ = load_golden("p000194_replay.npz")
    rec = _record(g)
    want, t0w = N.windows_from_grids(g["grids"])                   # pandas-built grids (fixture)
    x, t0 = S.assemble_windows_gpu(rec, "cuda:0")
    assert x.shape == (1615, 10, 120) and x.dtype == torch.float32
    assert np.array_equal(t0.cpu().numpy(), t0w)
    got = x.cpu().numpy()
    w32 = want.astype(np.float32)                                  # predictStream.py:155 .float()
    # pandas' rolling mean is an online add/remove sum, the device sums every window directly: after the f32 cast
    # (predictStream.py:155) nothing is off by more than an f32 ulp
    assert np.abs(got.astype(np.float64) - want).max() <= 2e-7 * np.abs(want).max()
    assert (np.abs(got - w32) <= 1e-9).mean() > 0.999
    assert (got[:, 4:, :] == 0).all()
    assert np.array_equal(got[0], g["x_first"].astype(np.float32))
    # and bit-for-bit the numpy restatement (same direct sums in the same order)
    assert np.array_equal(got, _np_windows(rec)[0].astype(np.float32))
    xb, _ = S.assemble_windows_gpu(rec, "cuda:0", dtype=torch.bfloat16)
    assert ((xb.float().cpu() - torch.from_numpy(w32)).abs() <= 2.0 ** -8 * torch.from_numpy(w32).abs() + 1e-30).all()


@pytest.mark.gpu
@pytest.mark.parametrize("seed,n,fs,kw", SYNTH_RECORDS)
def test_gpu_window_assembly_synthetic_records(seed, n, fs, kw):
    rec = _synthetic_record(seed, n, fs, **kw)
    fx = np.load(__import__("os").path.join(__import__("conftest").GOLDEN, "stream_synth_grids.npz"))
    want, t0w = N.windows_from_grids(fx[f"grid{seed}"])             # pandas fixture
    x, t0 = S.assemble_windows_gpu(rec, "cuda:0")
    assert tuple(x.shape) == want.shape and np.array_equal(t0.cpu().numpy(), t0w)
    got = x.cpu().numpy().astype(np.float64)
    scale = max(1.0, float(np.abs(want).max()))
    assert np.abs(got - want).max() <= 2e-7 * scale                # f32 rounding of the f64 grid values
    assert (got[:, len(S.selected_signals(rec)):, :] == 0).all()   # absent signals are exact zeros (predictStream.py:131)


@pytest.mark.gpu
def test_prep_rejects_bad_arguments():
    rec = _synthetic_record(9, 700, 1.0)
    rec.gains[0] = 0.0
    with pytest.raises(RuntimeError):
        S.assemble_windows_gpu(rec, "cuda:0")
    with pytest.raises(RuntimeError):
        S.assemble_windows_gpu(_synthetic_record(9, 700, 1.0), "cpu")


# ------------------------------------------------------------------------------------------ GPU: streaming form (ring)
def _ring_windows(rec, per, dtype=torch.float32, P=1, as_f64=False):
    ring = S.PatientRing(P, rec.raw.shape[1], rec.fs, device="cuda:0", dtype=dtype)
    for p in range(P):
        if as_f64:
            ring.set_signals(p, S.selected_signals(rec))
        else:
            ring.set_record_signals(p, rec)
    src = rec.physical if as_f64 else rec.raw
    out, t0s = [], []
    for i0 in range(0, src.shape[0], per):
        chunk = np.repeat(src[None, i0:i0 + per], P, axis=0)
        r = ring.push(chunk)
        if r is not None:
            assert r[1] == len(out)
            out.append(r[0].clone()); t0s.append(r[2])
    ring.close()
    return out, np.array(t0s)


@pytest.mark.gpu
def test_ring_trigger_by_trigger_equals_whole_record_bit_for_bit():
    """p000194 pushed one trigger (one 1/60 Hz sample) at a time == b2cnn_prep_windows on the whole record."""
    g, _ = load_golden("p000194_replay.npz")
    rec = _record(g)
    whole, t0w = S.assemble_windows_gpu(rec, "cuda:0")
    out, t0 = _ring_windows(rec, per=1, P=3)
    # the stream also completes the window whose last points lie beyond the record's final sample label
    assert len(out) in (len(whole), len(whole) + 1)
    n = len(whole)
    assert np.array_equal(t0[:n], t0w.cpu().numpy())
    got = torch.stack([o[0] for o in out[:n]])
    assert torch.equal(got, whole)                                  # bit-for-bit
    assert all(torch.equal(o[0], o[1]) and torch.equal(o[0], o[2]) for o in out[:n:97])    # every patient slot
    # physical fp64 frames (what sendStream.py publishes) instead of ADC units: same bits
    out64, _ = _ring_windows(rec, per=1, as_f64=True)
    assert torch.equal(torch.stack([o[0] for o in out64[:n]]), whole)


@pytest.mark.gpu
@pytest.mark.parametrize("seed,n,fs,kw,per", [(2, 4000, 1.0, {"p_missing": 0.5}, 60), (4, 2500, 0.2, {"dead": 1}, 12),
                                              (8, 3000, 0.1, {"p_missing": 0.4}, 6), (5, 700, 1.0, {"p_missing": 0.0}, 37),
                                              (2, 4000, 1.0, {"p_missing": 0.5}, 7)])
def test_ring_matches_whole_record_on_synthetic_records(seed, n, fs, kw, per):
    """Other sampling rates, gaps, a dead signal, and pushes that do not line up with the 60 s stride."""
    rec = _synthetic_record(seed, n, fs, **kw)
    whole, t0w = S.assemble_windows_gpu(rec, "cuda:0")
    out, t0 = _ring_windows(rec, per=per)
    m = min(len(out), len(whole))
    assert m >= len(whole) - 1 and m > 0
    # a stream knows only the past: windows are equal wherever the whole-record pass did not back-fill from the future
    # and did not average a window that the record's end truncated (its very last grid points)
    got = torch.stack([o[0] for o in out[:m]])
    same = (got == whole[:m]).flatten(1).all(1).cpu().numpy()
    assert same[:-1].all(), np.nonzero(~same)
    outb, _ = _ring_windows(rec, per=per, dtype=torch.bfloat16)
    assert torch.equal(torch.stack([o[0] for o in outb[:m - 1]]).float(), whole[:m - 1].to(torch.bfloat16).float())


@pytest.mark.gpu
def test_ring_leading_gap_longer_than_a_window_is_zero_filled_until_the_signal_appears():
    rec = _synthetic_record(3, 1500, 1.0, lead_gap=900)              # signal 0 silent for the first 900 s
    out, _ = _ring_windows(rec, per=60)
    whole, _ = S.assemble_windows_gpu(rec, "cuda:0")
    assert (out[0][0, 0] == 0).all() and not (whole[0, 0] == 0).all()      # causal zeros vs back-fill from the future
    assert torch.equal(out[0][0, 1:], whole[0, 1:])                        # the other signals agree
    assert torch.equal(out[-2][0], whole[len(out) - 2])                    # once it has appeared: identical again


@pytest.mark.gpu
def test_replay_stream_one_predict_per_trigger_equals_whole_record_replay():
    """Three patients (the shipped record under three subject ids) scored trigger by trigger: one predict() over
    [3, 10, 120] per trigger; rows in the predictions schema (db/init.sql:24-28)."""
    g, _ = load_golden("p000194_replay.npz")
    _, sd = load_golden("mycnn5_xtestinput.npz")
    rec = _record(g)
    model = tskd_b200.B200MyCNN.from_reference(sd).to("cuda:0")
    rows = S.replay_stream(model, [rec, rec, rec], [194, 195, 196], ages=[65.0, 65.0, 65.0])
    n = 1615
    assert len(rows) >= 3 * n and [r[0] for r in rows[:3]] == [194, 195, 196]
    mine = np.array([r[2] for r in rows[0:3 * n:3]])
    assert rel_err(mine, g["probs"]) <= 1e-4                               # vs the unmodified reference's scores
    assert [r[1] for r in rows[0:3 * n:3]] == list(g["t0"])
    assert all(rows[3 * i][2] == rows[3 * i + 1][2] == rows[3 * i + 2][2] for i in range(0, n, 53))
    whole = S.replay(model, rec, subject_id=194, micro_batch=3)            # same kernel path (small batches)
    assert np.array_equal(np.array([r[2] for r in whole]), mine)           # bit-for-bit


@pytest.mark.gpu
def test_ring_argument_errors():
    ring = S.PatientRing(2, 7, 1.0, device="cuda:0")
    with pytest.raises(RuntimeError):
        ring.push(np.zeros((2, 200, 7), dtype=np.int16))                   # more than one stride of samples
    with pytest.raises(RuntimeError):
        ring.push(np.zeros((3, 10, 7), dtype=np.int16))                    # wrong patient count
    with pytest.raises(RuntimeError):
        ring.set_signals(5, [0, 1])
    with pytest.raises(RuntimeError):
        ring.set_signals(0, [0, 9])
    ring.close()
    with pytest.raises(RuntimeError):
        S.PatientRing(1, 7, 1.0, device="cpu")
