"""The oracle (numpy restatement + torch port) pinned against fixtures generated from the
UNMODIFIED reference (scripts/gen_golden.py).  CPU only."""
import numpy as np
import pytest
import torch

from oracle import att2s_numpy, aggr_numpy, torch_port

FEATS = ("kmer", "kpass", "ipd", "pw", "kmer2", "kpass2", "ipd2", "pw2")
EDGE_CASES = ("n1", "n3", "allN", "extreme", "zeroh0", "fraccode")


def _np_forward(ckpt, g, pfx="", dtype=np.float64):
    return att2s_numpy.forward(ckpt, *[g[pfx + k] for k in FEATS], g[pfx + "h0_f"], g[pfx + "h0_r"], dtype=dtype)


def test_numpy_oracle_matches_reference_synth(ckpt_att2s, golden_synth):
    logits, probs = _np_forward(ckpt_att2s, golden_synth)
    assert np.abs(logits - golden_synth["logits"]).max() < 2e-5
    assert np.abs(probs - golden_synth["probs"]).max() < 5e-6


def test_numpy_oracle_fp32_matches_reference_synth(ckpt_att2s, golden_synth):
    _, probs = _np_forward(ckpt_att2s, golden_synth, dtype=np.float32)
    assert np.abs(probs - golden_synth["probs"]).max() < 2e-5


@pytest.mark.parametrize("case", EDGE_CASES)
def test_numpy_oracle_edge_cases(ckpt_att2s, golden_edge, case):
    logits, probs = _np_forward(ckpt_att2s, golden_edge, pfx=case + ".")
    assert logits.shape == golden_edge[case + ".logits"].shape
    assert np.abs(probs - golden_edge[case + ".probs"]).max() < 5e-6


def test_numpy_oracle_seeded_h0_stream(ckpt_att2s, golden_seeded):
    # the reference draws h0 itself: strand 1 then strand 2 from torch's CPU generator
    torch.manual_seed(int(golden_seeded["tseed"]))
    h0_f = torch.randn(6, 64, 256).numpy()
    h0_r = torch.randn(6, 64, 256).numpy()
    assert np.array_equal(h0_f, golden_seeded["h0_f"]) and np.array_equal(h0_r, golden_seeded["h0_r"])
    g = dict(golden_seeded)
    _, probs = _np_forward(ckpt_att2s, g)
    assert np.abs(probs - golden_seeded["probs"]).max() < 5e-6


def test_torch_port_matches_reference(ckpt_att2s, golden_synth):
    m = torch_port.load_numpy_state(torch_port.Att2sPort(), ckpt_att2s)
    t = {k: torch.from_numpy(golden_synth[k]) for k in FEATS + ("h0_f", "h0_r")}
    logits, probs = m(*[t[k] for k in FEATS], t["h0_f"], t["h0_r"])
    # same ATen kernels as the reference -> (near) bit-identical
    assert np.abs(probs.detach().numpy() - golden_synth["probs"]).max() < 1e-6


def test_batchloop_golden_is_reproducible_from_h0_stream(ckpt_att2s, golden_batchloop):
    """Reference _call_mods2s semantics: for each 512-slice in order draw h0_f then h0_r; output
    round(p1/(p0+p1), 6)  (reference call_modifications.py:177-224)."""
    g = golden_batchloop
    n, bs = g["kmer"].shape[0], int(g["batch_size"])
    torch.manual_seed(int(g["tseed"]))
    out = []
    for s in range(0, n, bs):
        e = min(n, s + bs)
        h0_f = torch.randn(6, e - s, 256).numpy()
        h0_r = torch.randn(6, e - s, 256).numpy()
        _, probs = att2s_numpy.forward(ckpt_att2s, *[g[k][s:e] for k in FEATS], h0_f, h0_r, dtype=np.float32)
        out.append(att2s_numpy.prob1_norm(probs))
    out = np.concatenate(out)
    assert int(g["batch_num"]) == 3 and len(out) == n
    assert np.abs(out - g["prob1"]).max() < 2e-5


def test_aggr_numpy_oracle(ckpt_aggr, golden_aggr):
    g = golden_aggr
    pm, hm = aggr_numpy.build_windows(g["pos"], list(g["histos"]))
    assert np.array_equal(pm, g["pos_mat"])
    raw = aggr_numpy.forward(ckpt_aggr, pm.astype(np.float32), hm.astype(np.float32), g["h0"])
    assert np.abs(raw - g["raw"]).max() < 1e-5


def test_aggr_loop_golden(ckpt_aggr, golden_aggr):
    g = golden_aggr
    pm, hm = aggr_numpy.build_windows(g["pos"], list(g["histos"]))
    outs = []
    for s, h0 in ((0, g["loop_h0_b0"]), (1024, g["loop_h0_b1"])):
        e = s + h0.shape[1]
        outs.append(aggr_numpy.postprocess(aggr_numpy.forward(ckpt_aggr, pm[s:e].astype(np.float32),
                                                              hm[s:e].astype(np.float32), h0))[:, 0])
    assert np.abs(np.concatenate(outs) - g["loop_probs"]).max() < 2e-6


def test_aggr_torch_port(ckpt_aggr, golden_aggr):
    g = golden_aggr
    m = torch_port.load_numpy_state(torch_port.AggrPort(), ckpt_aggr)
    pm, hm = aggr_numpy.build_windows(g["pos"], list(g["histos"]))
    raw = m(torch.tensor(pm, dtype=torch.float), torch.tensor(np.array(hm), dtype=torch.float), torch.from_numpy(g["h0"]))
    assert np.abs(raw.detach().numpy() - g["raw"]).max() < 1e-6


def test_lstm_numpy_oracle_matches_reference():
    """attbilstm2s (reference models.py:48-51): numpy LSTM restatement vs the reference's own forward on a seeded
    random initialisation (no checkpoint ships for this model type)."""
    from tests.conftest import load_npz
    g = load_npz("att2s_lstm.npz")
    sd = {k[3:]: v for k, v in g.items() if k.startswith("sd.")}
    logits, probs = att2s_numpy.forward_lstm(sd, *[g[k] for k in ("kmer", "kpass", "ipd", "pw", "kmer2", "kpass2", "ipd2", "pw2")],
                                             (g["h0_f"], g["c0_f"]), (g["h0_r"], g["c0_r"]), num_layers=2)
    assert np.abs(probs - g["probs"]).max() <= 1e-6
    assert np.abs(logits - g["logits"]).max() <= 1e-5


@pytest.mark.parametrize("cell", ["gru", "lstm"])
def test_2s2_numpy_oracle_matches_reference(cell):
    """ModelAttRNN2 (attbigru2s2 / attbilstm2s2, reference models.py:221-382): numpy restatement vs the reference's own
    forward on seeded random weights."""
    from tests.conftest import load_npz
    g = load_npz("att2s2.npz")
    sd = {k[len(cell) + 4:]: v for k, v in g.items() if k.startswith(cell + ".sd.")}
    h = (g[cell + ".h0"], g[cell + ".h1"]) if cell == "gru" else ((g[cell + ".h0"], g[cell + ".h1"]), (g[cell + ".h2"], g[cell + ".h3"]))
    logits, probs = att2s_numpy.forward_2s2(sd, *[g[k] for k in ("kmer", "kpass", "ipd", "pw", "kmer2", "kpass2", "ipd2", "pw2")],
                                            h[0], h[1], num_layers=2, cell=cell)
    assert np.abs(probs - g[cell + ".probs"]).max() <= 1e-6
    assert np.abs(logits - g[cell + ".logits"]).max() <= 1e-5


def test_transenc_numpy_oracle_matches_reference():
    """ModelTransEnc (transencoder2s, reference models.py:451-620): numpy restatement (conv stack, BatchNorm in eval
    mode, post-norm encoder layers) vs the reference's own forward on seeded random weights."""
    from tests.conftest import load_npz
    g = load_npz("transenc.npz")
    sd = {k[3:]: v for k, v in g.items() if k.startswith("sd.")}
    logits, probs = att2s_numpy.forward_transenc(sd, *[g[k] for k in ("kmer", "kpass", "ipd", "pw", "kmer2", "kpass2", "ipd2", "pw2")],
                                                 num_layers=2, nhead=4)
    assert np.abs(probs - g["probs"]).max() <= 1e-6
    assert np.abs(logits - g["logits"]).max() <= 1e-5


def test_norm_mad_oracle_matches_the_reference_normaliser():
    """`--norm mad` (reference extract_features.py:181-199 with statsmodels 0.14's robust.scale.mad restated: the package is
    not in this image).  Fixture: scripts/gen_golden.py norm_mad -- the reference's _normalize_signals around that formula."""
    from oracle.extract_numpy import _normalize_signals, statsmodels_mad, MAD_C
    from tests.conftest import load_npz
    g = load_npz("norm_mad.npz")
    n = 0
    while "sig%d" % n in g:
        out = np.asarray(_normalize_signals(g["sig%d" % n], "mad"), dtype=np.float64)
        assert np.array_equal(out, g["out%d" % n]), n
        n += 1
    assert n == 9
    # known answers: MAD of 1..9 is 2 (median 5, deviations 0 1 1 2 2 3 3 4 4), scaled by 1 / norm.ppf(3/4)
    assert statsmodels_mad(np.arange(1, 10)) == 2.0 / MAD_C
    assert abs(MAD_C - 0.6744897501960817) == 0.0
    assert np.isnan(statsmodels_mad(np.array([])))
