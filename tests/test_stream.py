"""Streaming replay (BASELINE.json configs[4]) at the model boundary: the restated window logic
(tskd_b200/stream.py) on CPU, and GPU predict() parity on identical x_arr against the scores the
unmodified reference produced for the shipped record p000194 (tests/golden/make_golden.py)."""
import numpy as np
import pytest
import torch

import tskd_b200
from tskd_b200 import stream as S
from conftest import load_golden, rel_err
from oracle import mycnn_torch as O


def _record(g):
    return S.NumericsRecord(tuple(str(n) for n in g["names"]), g["gains"], g["baselines"], float(g["fs"]), g["raw"])


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


def test_grid_smoothing_semantics():
    fs = 1 / 60.0
    s = np.array([np.nan, 60.0, 90.0, np.nan, 30.0, 30.0, 30.0])      # one sample per minute
    g = S.smooth_to_grid(s, fs)
    assert len(g) == 6 * 12 + 1
    assert g[0] == 60.0                       # nothing valid yet -> back-filled from the first valid mean
    assert g[12] == 60.0                      # window (-120, 60]: {nan, 60}
    assert g[24] == 75.0                      # (−60, 120]: {nan, 60, 90} -> 75
    assert g[35] == 75.0 and g[36] == 75.0    # (0, 180]: {60, 90, nan}
    assert g[48] == 60.0                      # (60, 240]: {90, nan, 30}
    assert g[72] == 30.0
    z = S.smooth_to_grid(np.full(5, np.nan), fs)
    assert (z == 0).all()                     # never observed -> zeros (processStream.py 0-fill)


def test_window_assembly_matches_golden_inputs():
    g, _ = load_golden("p000194_replay.npz")
    x, t0 = S.assemble_windows(_record(g))
    assert x.shape == (1615, 10, 120) and x.dtype == np.float64
    assert np.array_equal(x[0], g["x_first"]) and np.array_equal(x[-1], g["x_last"])
    assert (x[:, 4:, :] == 0).all()           # the six signals the record lacks (predictStream.py:131)
    assert t0[1] - t0[0] == 60.0 and np.array_equal(t0, g["t0"])
    # consecutive windows overlap by 108 of 120 points (600 s window, 60 s slide)
    assert np.array_equal(x[1, 0, :108], x[0, 0, 12:])


@pytest.mark.gpu
def test_replay_parity_gpu_vs_reference_scores():
    g, _ = load_golden("p000194_replay.npz")
    g5, sd = load_golden("mycnn5_xtestinput.npz")
    rec = _record(g)
    model = tskd_b200.B200MyCNN.from_reference(sd).to("cuda:0")
    rows = S.replay(model, rec, subject_id=194, age=65.0, on_gpu=False)  # host-assembled windows, ONE batched predict()
    assert len(rows) == 1615 and rows[0][0] == 194 and rows[1][1] - rows[0][1] == 60.0
    probs = np.array([r[2] for r in rows])
    assert rel_err(probs, g["probs"]) <= 1e-4                            # vs the unmodified reference
    rows_mb = S.replay(model, rec, subject_id=194, age=65.0, micro_batch=16, on_gpu=False)   # BATCHSIZE = 16 (config.cfg:26)
    # micro-batches of 16 take the single-launch small-window kernel, the full batch the general path
    assert rel_err(np.array([r[2] for r in rows_mb]), probs) <= 1e-6
    # logits on identical x_arr: GPU vs golden vs oracle per-window loop
    x, _ = S.assemble_windows(rec)
    xt = torch.from_numpy(x).float()
    logit = model.predict(xt.cuda(), 65.0).cpu().numpy()
    assert rel_err(logit, g["logits"]) <= 1e-4
    ref = O.RefMyCNN(O.ARCH_MYCNN5); ref.load_state_dict(sd); ref.eval()
    want = O.ref_independent_loop(ref, xt[:200], torch.full((200,), 65.0)).numpy()
    assert np.array_equal(want, g["logits"][:200]) and rel_err(logit[:200], want) <= 1e-4


# ---- f2 + f1 on the device (csrc/b2cnn_prep.cu); the numpy restatement above is its oracle ----

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


def test_prep_window_count_matches_host_logic():
    """Host-side arithmetic of the C ABI (no device work): window count of b2cnn_prep_window_count == numpy's."""
    import ctypes
    from tskd_b200 import capi
    lib = capi.load_library()
    cfg = capi.PrepConfig(S.N_CHANNELS, S.WINDOW_POINTS, S.GRID_S, S.SMOOTH_S, S.STRIDE_S)
    for n, fs in [(1625, 1 / 60), (700, 1.0), (595, 1.0), (596, 1.0), (3000, 0.2), (12, 1 / 60), (2, 1.0)]:
        t_last = (n - 1) * (1.0 / fs)
        n_grid = int(np.floor(t_last / S.GRID_S)) + 1
        want = len(np.arange(0, n_grid - S.WINDOW_POINTS + 1, S.STRIDE_S // S.GRID_S))
        assert lib.b2cnn_prep_window_count(n, fs, ctypes.byref(cfg)) == want, (n, fs)
    assert lib.b2cnn_prep_window_count(0, 1.0, ctypes.byref(cfg)) < 0
    bad = capi.PrepConfig(10, 120, 5, 180, 62)                    # stride not a multiple of the grid
    assert lib.b2cnn_prep_window_count(1000, 1.0, ctypes.byref(bad)) < 0


@pytest.mark.gpu
def test_gpu_window_assembly_matches_host_restatement_on_shipped_record():
    g, _ = load_golden("p000194_replay.npz")
    rec = _record(g)
    want, t0w = S.assemble_windows(rec)
    x, t0 = S.assemble_windows_gpu(rec, "cuda:0")
    assert x.shape == (1615, 10, 120) and x.dtype == torch.float32
    assert np.array_equal(t0.cpu().numpy(), t0w)
    got = x.cpu().numpy()
    w32 = want.astype(np.float32)                                  # predictStream.py:155 .float()
    # the device sums every window directly, numpy takes prefix-sum differences (1e-12 cancellation residue where a
    # window is all zeros); after the f32 cast (predictStream.py:155) nothing is off by more than an f32 ulp
    assert np.abs(got.astype(np.float64) - want).max() <= 2e-7 * np.abs(want).max()
    assert (np.abs(got - w32) <= 1e-9).mean() > 0.999
    assert (got[:, 4:, :] == 0).all()
    assert np.array_equal(got[0], g["x_first"].astype(np.float32))
    xb, _ = S.assemble_windows_gpu(rec, "cuda:0", dtype=torch.bfloat16)
    assert ((xb.float().cpu() - torch.from_numpy(w32)).abs() <= 2.0 ** -8 * torch.from_numpy(w32).abs() + 1e-30).all()


@pytest.mark.gpu
@pytest.mark.parametrize("seed,n,fs,kw", [
    (1, 1625, 1 / 60, {}), (2, 4000, 1.0, {"p_missing": 0.5}), (3, 900, 1.0, {"lead_gap": 400}),
    (4, 2500, 0.2, {"dead": 1}), (5, 700, 1.0, {"p_missing": 0.0}), (6, 50000, 1.0, {"p_missing": 0.9}),
    (7, 640, 1.0, {"n_sig": 2}),
])
def test_gpu_window_assembly_synthetic_records(seed, n, fs, kw):
    rec = _synthetic_record(seed, n, fs, **kw)
    want, t0w = S.assemble_windows(rec)
    x, t0 = S.assemble_windows_gpu(rec, "cuda:0")
    assert tuple(x.shape) == want.shape and np.array_equal(t0.cpu().numpy(), t0w)
    got = x.cpu().numpy().astype(np.float64)
    scale = max(1.0, float(np.abs(want).max()))
    assert np.abs(got - want).max() <= 2e-7 * scale                # f32 rounding of the f64 grid values
    assert (got[:, len(S.selected_signals(rec)):, :] == 0).all()   # absent signals are exact zeros (predictStream.py:131)


@pytest.mark.gpu
def test_replay_on_gpu_equals_host_assembled_replay():
    g, _ = load_golden("p000194_replay.npz")
    _, sd = load_golden("mycnn5_xtestinput.npz")
    rec = _record(g)
    model = tskd_b200.B200MyCNN.from_reference(sd).to("cuda:0")
    rows_h = S.replay(model, rec, subject_id=194, on_gpu=False)
    rows_d = S.replay(model, rec, subject_id=194, on_gpu=True)
    assert len(rows_h) == len(rows_d) == 1615
    assert [r[:2] for r in rows_h] == [r[:2] for r in rows_d]
    assert rel_err(np.array([r[2] for r in rows_d]), g["probs"]) <= 1e-4
    assert rel_err(np.array([r[2] for r in rows_d]), np.array([r[2] for r in rows_h])) <= 1e-6


@pytest.mark.gpu
def test_prep_rejects_bad_arguments():
    rec = _synthetic_record(9, 700, 1.0)
    rec.gains[0] = 0.0
    with pytest.raises(RuntimeError):
        S.assemble_windows_gpu(rec, "cuda:0")
    with pytest.raises(RuntimeError):
        S.assemble_windows_gpu(_synthetic_record(9, 700, 1.0), "cpu")


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
