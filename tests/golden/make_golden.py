"""Generate the committed golden fixtures by running the UNMODIFIED reference.

Run in the authoring container only (``/root/reference`` does not exist on the GPU box):

    python tests/golden/make_golden.py

What is executed is the reference's own code: ``bin/models.py`` (class ``MyCNN`` and its
``forward``), the shipped checkpoints ``model/MyCNN{2,3,4,5}.pth``, the fixture
``explore_output/X.TESTINPUT`` and (verbatim, with stub modules for the absent ``wfdb`` /
``pyspark`` imports) ``bin/utils.py``'s ``create_batch`` / ``get_arr`` / ``run_model``.
Nothing from this repo's product code is used to produce the expected values; the streaming fixtures
(section 4) get their 5-second grids from pandas (oracle/stream_pandas.py: the reference notebook's own
``resample`` / ``rolling`` calls), never from the numpy restatement or the CUDA kernels they pin.
"""
import os
import sys
import types
import warnings

import numpy as np
import pandas as pd
import torch
import torch.nn as nn

warnings.filterwarnings("ignore")
REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(REF, "bin"))
from models import MyCNN  # noqa: E402  (the reference class)

import __main__  # noqa: E402

__main__.MyCNN = MyCNN  # the legacy pickles name their class "__main__.MyCNN"


def sd_arrays(m):
    return {"sd." + k: v.detach().cpu().numpy() for k, v in m.state_dict().items()}


def load_ckpt(n):
    m = torch.load(f"{REF}/model/MyCNN{n}.pth", weights_only=False, map_location="cpu")
    m.eval()
    return m


@torch.no_grad()
def loop(m, x, a):
    return torch.cat([m(x[i:i + 1], a[i:i + 1]) for i in range(x.shape[0])])


@torch.no_grad()
def features(m, x):
    x = torch.tanh(m.conv1(x)); x = m.pool(x)
    x = torch.tanh(m.conv2(x)); x = m.pool(x)
    return x.view(-1, m.MAGICNUM)


# ---------------------------------------------------------------- 1. MyCNN5 + X.TESTINPUT
def golden_mycnn5():
    m = load_ckpt(5)
    df = pd.read_csv(f"{REF}/explore_output/X.TESTINPUT")
    x64 = np.swapaxes(df.values[None], 1, 2)            # utils.py:538 (n,120,10)->(n,10,120)
    x = torch.from_numpy(x64).type(torch.FloatTensor)   # predictStream.py:155
    out = {}
    with torch.no_grad():
        for age in (50.0, 65.0):
            a = torch.tensor([age])
            o = m(x, a)
            out[f"logit_age{int(age)}"] = o.numpy()
            out[f"prob_age{int(age)}"] = torch.sigmoid(o).numpy()
        out["features"] = features(m, x).numpy()
        # the notebook cell (explore_torch.ipynb:4277-4310) prints [[0.5668570399284363]]
        out["notebook_print"] = np.array([[0.5668570399284363]])
        # batched calls: random "vital-sign-like" windows, seed 0 (SURVEY section 4 probe)
        torch.manual_seed(0)
        xb = torch.randn(8, 10, 120) * 20 + 50
        ab = torch.tensor([15., 30., 45., 50., 65., 70., 80., 20.])
        out["xb"] = xb.numpy(); out["ab"] = ab.numpy()
        out["xb_seq_logits"] = m(xb, ab).numpy()          # utils.py:249 semantics
        out["xb_ind_logits"] = loop(m, xb, ab).numpy()    # predictStream.py:157 semantics
        torch.manual_seed(1)
        xn = torch.randn(8, 10, 120)
        out["xn"] = xn.numpy()
        out["xn_seq_logits"] = m(xn, ab).numpy()
        out["xn_ind_logits"] = loop(m, xn, ab).numpy()
        out["xn_features"] = features(m, xn).numpy()
    # verbatim utils.run_model (utils.py:671-692)
    for name in ("wfdb", "pyspark", "pyspark.sql"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["pyspark.sql"].SparkSession = object
    cwd = os.getcwd(); os.chdir(REF)
    try:
        import utils as ref_utils
        y_pred, y_prob = ref_utils.run_model(m, torch.device("cpu"), df, 50.0)
    finally:
        os.chdir(cwd)
    out["run_model_pred"] = np.array(y_pred); out["run_model_prob"] = np.array(y_prob, dtype=np.float64)
    np.savez_compressed(os.path.join(OUT, "mycnn5_xtestinput.npz"), x=x.numpy(), x_f64=x64,
                        **sd_arrays(m), **out)
    print("mycnn5:", out["logit_age50"], out["prob_age50"], y_pred, y_prob)


# ---------------------------------------------------------------- 2. older checkpoints
def golden_old_ckpts():
    for n in (2, 3, 4):
        m = load_ckpt(n)
        C = m.conv1.in_channels
        m.MAGICNUM = m.lstm.input_size           # models.py needs the attribute (27 here)
        torch.manual_seed(10 + n)
        xb = torch.randn(6, C, 120) * 20 + 50
        ab = torch.tensor([15., 30., 45., 50., 65., 80.])
        with torch.no_grad():
            seq = m(xb, ab).numpy()               # bin/models.py forward => age coef 1e-8
            ind = loop(m, xb, ab).numpy()
            f = features(m, xb).numpy()
        meta = np.array([C, m.conv1.kernel_size[0], m.conv2.kernel_size[0], m.pool.kernel_size,
                         m.pool.stride, m.lstm.input_size])
        np.savez_compressed(os.path.join(OUT, f"mycnn{n}_ckpt.npz"), xb=xb.numpy(), ab=ab.numpy(),
                            seq_logits_coef1e8=seq, ind_logits_coef1e8=ind, features=f, meta=meta,
                            **sd_arrays(m))
        print(f"mycnn{n}:", meta, ind[:3])


# ---------------------------------------------------------------- 3. stretched architectures
def stretched(kind, C, W, seed):
    """the reference class with conv1/lstm/MAGICNUM re-instantiated."""
    torch.manual_seed(seed)
    g = MyCNN()
    if kind == "mycnn5":
        k1, pk, ps = 10, 3, 2
    else:
        k1, pk, ps = 5, 2, 2
        g.pool = nn.MaxPool1d(kernel_size=pk, stride=ps)
    l1 = W - k1 + 1; p1 = (l1 - pk) // ps + 1; l2 = p1 - 5 + 1; L = (l2 - pk) // ps + 1
    g.MAGICNUM = L
    g.conv1 = nn.Conv1d(C, 4, k1)
    g.lstm = nn.LSTM(L, 16, 2)
    g.eval()
    return g, L


def golden_stretched():
    cases = [("mycnn5", 3, 1500, 4, 0), ("mycnn3", 3, 1500, 4, 0),
             ("mycnn3", 3, 7500, 1, 0),      # BASELINE.json configs[0]: [1,3,7500] fp32
             ("mycnn5", 3, 7500, 2, 0)]
    for kind, C, W, B, seed in cases:
        g, L = stretched(kind, C, W, seed)
        torch.manual_seed(1)
        x = torch.randn(B, C, W)
        age = torch.linspace(15, 80, B) if B > 1 else torch.tensor([65.0])
        xbf = x.to(torch.bfloat16).float()       # bf16-representable copy (config 2 dtype)
        with torch.no_grad():
            rec = dict(x=x.numpy(), age=age.numpy(), L=np.array(L),
                       seq_logits=g(x, age).numpy(), ind_logits=loop(g, x, age).numpy(),
                       features=features(g, x).numpy(),
                       bf16_ind_logits=loop(g, xbf, age).numpy())
        np.savez_compressed(os.path.join(OUT, f"stretched_{kind}_c{C}_w{W}_b{B}.npz"),
                            **rec, **sd_arrays(g))
        print(kind, C, W, B, "L_out", L, rec["ind_logits"][:2])


# ---------------------------------------------------------------- 4. streaming replay (configs[4])
def _synthetic_record(seed, n, fs, n_sig=7, p_missing=0.2, lead_gap=0, dead=None):
    """The synthetic numerics records of tests/test_stream.py (same generator, same seeds)."""
    import tskd_b200.stream as S
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


SYNTH_RECORDS = [(1, 1625, 1 / 60, {}), (2, 4000, 1.0, {"p_missing": 0.5}), (3, 900, 1.0, {"lead_gap": 400}),
                 (4, 2500, 0.2, {"dead": 1}), (5, 700, 1.0, {"p_missing": 0.0}), (6, 50000, 1.0, {"p_missing": 0.9}),
                 (7, 640, 1.0, {"n_sig": 2}), (8, 3000, 0.1, {"p_missing": 0.4})]


def golden_replay():
    """The shipped numerics record p000194 (7 signals @ 1/60 Hz, 1625 samples): the 5-second grid is built with
    PANDAS -- the reference's own offline calls ``resample('5S').first()`` / ``rolling('3min').mean()``
    (bin/explore_torch.ipynb:402,405) + the streaming job's ffill / bfill / fillna(0) (bin/processStream.py:62-123),
    oracle/stream_pandas.py -- cut into the 600 s / 60 s windows of bin/predictStream.py:245-259, and scored by the
    UNMODIFIED reference model called once per window exactly as predictStream.py:154-162 does.  Nothing of the
    product (nor of the numpy restatement) produces an expected value here; only the WFDB byte reader is shared,
    and it is pinned by the header's checksums."""
    root = os.path.dirname(os.path.dirname(OUT))
    sys.path.insert(0, root)
    import tskd_b200.stream as S
    from oracle import stream_pandas as P
    d = f"{REF}/data/waveform/physionet.org/files/mimic3wdb-matched/1.0/p00/p000194"
    rec = S.NumericsRecord.from_wfdb_files(f"{d}/p000194-2112-05-23-14-34n.hea", f"{d}/3400942n.dat")
    csum = [int(np.int16(rec.raw[:, i].astype(np.int64).sum() & 0xffff)) for i in range(rec.raw.shape[1])]
    assert csum == [3240, 20492, 29088, 10310, -27206, -29717, -28780], csum      # header checksums
    sel = S.selected_signals(rec)
    phys = rec.physical
    grids = np.stack([P.grid_notebook(phys[:, s_], rec.fs) for s_ in sel])
    grids_spark = np.stack([P.grid_spark(phys[:, s_], rec.fs) for s_ in sel])
    assert np.abs(grids - grids_spark).max() <= 1e-9                             # lattice record: both pipelines agree
    n_grid = grids.shape[1]
    starts = np.arange(0, n_grid - 120 + 1, 12)
    x = np.zeros((len(starts), 10, 120))
    for ch in range(len(sel)):
        x[:, ch, :] = np.lib.stride_tricks.sliding_window_view(grids[ch], 120)[starts]
    t0 = starts * 5.0
    m = load_ckpt(5)
    probs, logits = [], []
    with torch.no_grad():
        for i in range(x.shape[0]):
            x_arr = torch.from_numpy(x[i:i + 1]).type(torch.FloatTensor).float()          # predictStream.py:155
            a_arr = torch.from_numpy(np.array(65.0)).type(torch.FloatTensor).unsqueeze(0)  # :149,156
            output = m(x_arr, a_arr)                                                       # :157
            y = torch.sigmoid(output)                                                      # :160
            logits.append(output.numpy()[0]); probs.append(y.numpy().tolist()[0])
    np.savez_compressed(os.path.join(OUT, "p000194_replay.npz"), raw=rec.raw, names=np.array(rec.names),
                        gains=rec.gains, baselines=rec.baselines, fs=np.array(rec.fs),
                        logits=np.array(logits, dtype=np.float32), probs=np.array(probs, dtype=np.float64),
                        t0=t0, x_first=x[0], x_last=x[-1], grids=grids, grids_unfilled=np.stack(
                            [P.grid_notebook(phys[:, s_], rec.fs, fill=False) for s_ in sel]))
    print("replay:", x.shape, "selected", [rec.names[i] for i in sel], logits[:3], probs[-1])
    # synthetic records (gaps, dead signals, several sampling rates): the streaming aggregate written with pandas
    out = {}
    for seed, n, fs, kw in SYNTH_RECORDS:
        r = _synthetic_record(seed, n, fs, **kw)
        ph = r.physical
        sl = S.selected_signals(r)
        g = np.stack([P.grid_spark(ph[:, s_], r.fs) for s_ in sl])
        out[f"grid{seed}"] = g
        if fs <= 0.2:       # at most one sample per 5-second bin: the notebook's resample().first() is the same thing
            assert np.abs(g - np.stack([P.grid_notebook(ph[:, s_], r.fs) for s_ in sl])).max() <= 1e-9
    np.savez_compressed(os.path.join(OUT, "stream_synth_grids.npz"), **out)
    print("synthetic grids:", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    if "--replay-only" in sys.argv:
        golden_replay(); sys.exit(0)
    golden_mycnn5()
    golden_old_ckpts()
    golden_stretched()
    golden_replay()
