"""CPU: pin the oracle (oracle/) against the golden vectors produced by the UNMODIFIED
reference (tests/golden/make_golden.py) -- the only known-answer the reference holds is
explore_torch.ipynb:4271 (MyCNN5 + X.TESTINPUT -> 0.5668570399284363)."""
import numpy as np
import pytest
import torch

from conftest import load_golden, rel_err
from oracle import mycnn_c
from oracle import mycnn_torch as O


def _ref_from(sd, arch):
    m = O.RefMyCNN(arch)
    m.load_state_dict(sd)
    return m.eval()


def test_known_answer_bit_exact(golden5):
    g, sd = golden5
    m = _ref_from(sd, O.ARCH_MYCNN5)
    x = torch.from_numpy(g["x"])
    for age in (50, 65):
        logit = O.ref_sequence(m, x, torch.tensor([float(age)]))
        assert np.array_equal(logit.numpy(), g[f"logit_age{age}"])
    prob = torch.sigmoid(O.ref_sequence(m, x, torch.tensor([50.0]))).numpy().tolist()
    assert prob == [0.5668570399284363]                    # explore_torch.ipynb:4271
    assert g["run_model_prob"].tolist() == [0.5668570399284363]   # verbatim utils.run_model
    assert np.array_equal(O.ref_features(m, x).numpy(), g["features"])
    pred, p = O.post_process(O.ref_sequence(m, x, torch.tensor([50.0])))
    assert pred == [1] and p == [0.5668570399284363]


def test_batch_semantics_match_reference(golden5):
    g, sd = golden5
    m = _ref_from(sd, O.ARCH_MYCNN5)
    for tag in ("xb", "xn"):
        x, a = torch.from_numpy(g[tag]), torch.from_numpy(g["ab"])
        assert np.array_equal(O.ref_sequence(m, x, a).numpy(), g[f"{tag}_seq_logits"])
        assert np.array_equal(O.ref_independent_loop(m, x, a).numpy(), g[f"{tag}_ind_logits"])
        np.testing.assert_allclose(O.ref_independent(m, x, a).numpy(), g[f"{tag}_ind_logits"], rtol=0, atol=2e-7)
        # the LSTM really does carry state along the batch axis
        assert np.abs(g[f"{tag}_seq_logits"][1:] - g[f"{tag}_ind_logits"][1:]).max() > 1e-4
        assert g[f"{tag}_seq_logits"][0] == g[f"{tag}_ind_logits"][0]


@pytest.mark.parametrize("n", [2, 3, 4])
def test_older_checkpoints(n):
    g, sd = load_golden(f"mycnn{n}_ckpt.npz")
    arch = O.ARCHS[f"mycnn{n}"]
    # the goldens ran bin/models.py's forward => age coefficient 1e-8 (save-time value unknown)
    from dataclasses import replace
    m = _ref_from(sd, replace(arch, age_coef=1e-8))
    x, a = torch.from_numpy(g["xb"]), torch.from_numpy(g["ab"])
    assert list(g["meta"]) == [arch.in_channels, arch.k1, arch.k2, arch.pool_k, arch.pool_s, arch.l_out]
    assert np.array_equal(O.ref_sequence(m, x, a).numpy(), g["seq_logits_coef1e8"])
    assert np.array_equal(O.ref_independent_loop(m, x, a).numpy(), g["ind_logits_coef1e8"])
    assert np.array_equal(O.ref_features(m, x).numpy(), g["features"])


@pytest.mark.parametrize("name,kind,C,W", [
    ("stretched_mycnn5_c3_w1500_b4.npz", "mycnn5", 3, 1500),
    ("stretched_mycnn3_c3_w1500_b4.npz", "mycnn3", 3, 1500),
    ("stretched_mycnn3_c3_w7500_b1.npz", "mycnn3", 3, 7500),
    ("stretched_mycnn5_c3_w7500_b2.npz", "mycnn5", 3, 7500),
])
def test_stretched_architectures(name, kind, C, W):
    g, sd = load_golden(name)
    from dataclasses import replace
    arch = replace(O.stretched(O.ARCHS[kind], C, W), age_coef=1e-8, has_out12=True)
    assert arch.l_out == int(g["L"])
    m = _ref_from(sd, arch)
    x, a = torch.from_numpy(g["x"]), torch.from_numpy(g["age"])
    assert np.array_equal(O.ref_sequence(m, x, a).numpy(), g["seq_logits"])
    assert np.array_equal(O.ref_independent_loop(m, x, a).numpy(), g["ind_logits"])
    assert np.array_equal(O.ref_features(m, x).numpy(), g["features"])


def test_c_restatement_matches_golden(golden5):
    g, sd = golden5
    blob = mycnn_c.pack_blob(sd)
    for age in (50, 65):
        y64 = mycnn_c.forward(O.ARCH_MYCNN5, blob, g["x"], np.array([age], np.float32), precision="f64")
        assert rel_err(y64, g[f"logit_age{age}"]) < 2e-6
        y32 = mycnn_c.forward(O.ARCH_MYCNN5, blob, g["x"], np.array([age], np.float32), precision="f32")
        assert rel_err(y32, g[f"logit_age{age}"]) < 1e-5
    for mode, key in (("independent", "xn_ind_logits"), ("sequence", "xn_seq_logits")):
        y, f = mycnn_c.forward(O.ARCH_MYCNN5, blob, g["xn"], g["ab"], mode=mode, want_features=True)
        assert rel_err(y, g[key]) < 5e-6
        assert np.abs(f - g["xn_features"]).max() < 1e-6


def test_c_restatement_stretched_and_old():
    g, sd = load_golden("stretched_mycnn3_c3_w7500_b1.npz")
    from dataclasses import replace
    arch = replace(O.stretched(O.ARCH_MYCNN3, 3, 7500), age_coef=1e-8)
    y = mycnn_c.forward(arch, mycnn_c.pack_blob(sd), g["x"], g["age"])
    assert rel_err(y, g["ind_logits"]) < 1e-5
    g, sd = load_golden("mycnn4_ckpt.npz")
    arch = replace(O.ARCH_MYCNN4, age_coef=1e-8)
    y = mycnn_c.forward(arch, mycnn_c.pack_blob(sd), g["xb"], g["ab"], mode="sequence")
    assert rel_err(y, g["seq_logits_coef1e8"]) < 2e-5


def test_nan_inf_semantics():
    """MaxPool1d and tanh propagate NaN; +inf saturates."""
    p = torch.nn.MaxPool1d(3, 2)(torch.tensor([[[1.0, float("nan"), 2.0, 3.0, 4.0]]]))
    assert torch.isnan(p[0, 0, 0]) and p[0, 0, 1] == 4.0
    m = O.make_ref(O.ARCH_MYCNN5, seed=0)
    x = torch.randn(2, 10, 120)
    x[0, 3, 50] = float("nan")
    x[1, 3, 50] = float("inf")
    y = O.ref_independent_loop(m, x, torch.tensor([50.0, 50.0]))
    assert torch.isnan(y[0]) and torch.isfinite(y[1])
    blob = mycnn_c.pack_blob(m.state_dict())
    yc = mycnn_c.forward(O.ARCH_MYCNN5, blob, x.numpy(), np.array([50, 50], np.float32))
    assert np.isnan(yc[0]) and rel_err(yc[1:], y[1:].numpy()) < 1e-5


def test_view_contract():
    """L_out must equal MAGICNUM for a row to be a window (bin/models.py:29)."""
    assert O.ARCH_MYCNN5.l_out == 25 and O.ARCH_MYCNN4.l_out == 27
    assert O.stretched(O.ARCH_MYCNN5, 3, 75000).l_out == 18745
    assert O.stretched(O.ARCH_MYCNN3, 3, 75000).l_out == 18747
    assert (O.ARCH_MYCNN5.l1, O.ARCH_MYCNN5.p1, O.ARCH_MYCNN5.l2) == (111, 55, 51)
