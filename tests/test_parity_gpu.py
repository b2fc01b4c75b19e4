"""Parity of the CUDA path (through the C ABI) against the reference fixtures and the oracle."""
import numpy as np
import pytest
import torch

from oracle import att2s_numpy, aggr_numpy

pytestmark = pytest.mark.gpu

FEATS = ("kmer", "kpass", "ipd", "pw", "kmer2", "kpass2", "ipd2", "pw2")
EDGE_CASES = ("n1", "n3", "allN", "extreme", "zeroh0", "fraccode")
TOL = 1e-4  # BASELINE.json north_star: max |dprob| <= 1e-4 vs the reference fp32 CPU path


def args16(g, pfx=""):
    n = g[pfx + "kmer"].shape[0]
    z = torch.zeros(n)
    t = lambda k: torch.from_numpy(g[pfx + k])
    return (t("kmer"), t("kpass"), t("ipd"), z, t("pw"), z, z, z, t("kmer2"), t("kpass2"), t("ipd2"), z, t("pw2"), z, z, z)


@pytest.fixture(scope="module")
def model(ckpt_att2s):
    from ccsmeth_b200.models import ModelAttRNN
    m = ModelAttRNN(21, 3, 2, 0, 256, is_npass=True, model_type="attbigru2s", device=0, precision="fp32")
    d = m.state_dict()
    d.update({k: torch.from_numpy(v) for k, v in ckpt_att2s.items()})
    m.load_state_dict(d)
    m = m.cuda(0)
    m.eval()
    return m


def run(model, g, pfx=""):
    h0 = (torch.from_numpy(g[pfx + "h0_f"]), torch.from_numpy(g[pfx + "h0_r"]))
    logits, probs = model(*[a.cuda() for a in args16(g, pfx)], h0=h0)
    assert logits.is_cuda and probs.is_cuda
    return logits.cpu().numpy(), probs.cpu().numpy()


def test_fp32_matches_reference_synth(model, golden_synth):
    model.set_precision("fp32")
    logits, probs = run(model, golden_synth)
    assert np.abs(probs - golden_synth["probs"]).max() <= 2e-5
    assert np.abs(logits - golden_synth["logits"]).max() <= 1e-4


@pytest.mark.parametrize("case", EDGE_CASES)
def test_fp32_edge_cases(model, golden_edge, case):
    model.set_precision("fp32")
    logits, probs = run(model, golden_edge, case + ".")
    assert probs.shape == golden_edge[case + ".probs"].shape
    assert np.abs(probs - golden_edge[case + ".probs"]).max() <= 2e-5


def test_empty_batch(model):
    z = torch.zeros(0, 21).cuda()
    e = torch.zeros(0).cuda()
    logits, probs = model(z, z, z, e, z, e, e, e, z, z, z, e, z, e, e, e)
    assert logits.shape == (0, 2) and probs.shape == (0, 2)


def test_default_h0_stream_matches_reference(model, golden_seeded):
    """h0=None must reproduce the reference: torch.randn on the CPU generator, strand 1 then strand 2."""
    model.set_precision("fp32")
    g = golden_seeded
    torch.manual_seed(int(g["tseed"]))
    _, probs = model(*[a.cuda() for a in args16(g)])
    assert np.abs(probs.cpu().numpy() - g["probs"]).max() <= 2e-5


def test_rnn_stack_matches_oracle_internals(model, ckpt_att2s, golden_synth):
    """Layer-stack output (n, 21, 512) of both strands vs the numpy oracle's intermediate."""
    import ctypes
    from ccsmeth_b200 import _lib
    model.set_precision("fp32")
    g = {k: (v[:, :32] if k.startswith("h0") else v[:32]) for k, v in golden_synth.items()}
    run(model, g)
    buf = np.empty(32 * 2 * 21 * 512, dtype=np.float32)
    got = _lib.load().ccsm_debug_last_rnn_out(model._handle, buf.ctypes.data_as(ctypes.c_void_p), buf.size)
    assert got == buf.size
    out = buf.reshape(32, 2, 21, 512)
    _, _, it = att2s_numpy.forward(ckpt_att2s, *[g[k] for k in FEATS], g["h0_f"], g["h0_r"], return_internals=True)
    assert np.abs(out[:, 0] - it["layers0"][-1]).max() <= 1e-4
    assert np.abs(out[:, 1] - it["layers1"][-1]).max() <= 1e-4


def test_host_entry_matches_device_entry(model, golden_synth):
    model.set_precision("fp32")
    g = golden_synth
    feats = {k: g[k] for k in FEATS}
    logits, probs = model.forward_host(feats, h0=(g["h0_f"], g["h0_r"]))
    assert np.abs(probs.numpy() - g["probs"]).max() <= 2e-5


def test_large_batch_chunking_is_consistent(model, golden_synth):
    """n > the library's internal chunk (8192 sites): tile the 256 golden sites 40x, every copy must agree."""
    model.set_precision("fp32")
    g = golden_synth
    rep = 40
    big = {k: np.concatenate([g[k]] * rep, axis=1 if k.startswith("h0") else 0) for k in FEATS + ("h0_f", "h0_r")}
    _, probs = run(model, big)
    probs = probs.reshape(rep, 256, 2)
    assert np.abs(probs - g["probs"][None]).max() <= 2e-5


def test_aggr_matches_reference(ckpt_aggr, golden_aggr):
    from ccsmeth_b200.models import AggrAttRNN
    g = golden_aggr
    m = AggrAttRNN(11, 1, 1, 0, 32, binsize=20, model_type="attbigru", device=0)
    m.load_state_dict({k: torch.from_numpy(v) for k, v in ckpt_aggr.items()})
    m = m.cuda(0).eval()
    pm, hm = aggr_numpy.build_windows(g["pos"], list(g["histos"]))
    out = m(torch.tensor(pm, dtype=torch.float).cuda(), torch.tensor(np.array(hm), dtype=torch.float).cuda(),
            h0=torch.from_numpy(g["h0"]))
    assert np.abs(out.cpu().numpy() - g["raw"]).max() <= 1e-5


@pytest.mark.parametrize("n", [1, 255, 257, 1000, 70001])
def test_aggr_fused_kernel_vs_oracle_and_layerwise_path(ckpt_aggr, n, monkeypatch):
    """The one-kernel aggregate forward (csrc/aggr_fused.cu) against the numpy oracle on random windows of ragged
    sizes, and against the layer-by-layer fp32 kernels (CCSM_AGGR_UNFUSED=1)."""
    from ccsmeth_b200.models import AggrAttRNN
    rng = np.random.default_rng(n)
    histos = rng.random((n, 11, 20)).astype(np.float32)
    histos = np.round(histos / np.linalg.norm(histos, axis=2, keepdims=True), 6).astype(np.float32)
    offsets = rng.integers(0, 1500, (n, 11)).astype(np.float32)
    h0 = rng.standard_normal((2, n, 32)).astype(np.float32)
    m = AggrAttRNN(11, 1, 1, 0, 32, binsize=20, model_type="attbigru", device=0)
    m.load_state_dict({k: torch.from_numpy(v) for k, v in ckpt_aggr.items()})
    m = m.cuda(0).eval()
    fused = m(torch.from_numpy(offsets), torch.from_numpy(histos), h0=torch.from_numpy(h0)).cpu().numpy()
    monkeypatch.setenv("CCSM_AGGR_UNFUSED", "1")
    layerwise = m(torch.from_numpy(offsets), torch.from_numpy(histos), h0=torch.from_numpy(h0)).cpu().numpy()
    monkeypatch.delenv("CCSM_AGGR_UNFUSED")
    assert np.abs(fused - layerwise).max() <= 1e-5
    k = min(n, 2000)
    ref = aggr_numpy.forward(ckpt_aggr, offsets[:k], histos[:k], h0[:, :k], dtype=np.float64)
    assert np.abs(fused[:k] - ref).max() <= 1e-5


@pytest.mark.parametrize("mt,hid,layers", [("attbilstm", 32, 1), ("attbilstm", 64, 2), ("attbigru", 48, 2)])
def test_aggr_other_cells_and_shapes_vs_oracle(mt, hid, layers):
    """AggrAttRNN outside the fused kernel -- the LSTM cell (ccsm_forward_aggr_lstm, initial state (h0, c0)) and other
    --hid_rnn / --layer_rnn -- on seeded random weights against the numpy oracle (itself pinned to the reference's
    region caller for these model types, tests/test_pileup_cpu.py)."""
    from ccsmeth_b200.models import AggrAttRNN
    torch.manual_seed(11)
    m = AggrAttRNN(11, layers, 1, 0, hid, binsize=20, model_type=mt, device=0)
    sd = {k: v.numpy().copy() for k, v in m.state_dict().items()}
    m = m.cuda(0).eval()
    rng = np.random.default_rng(7)
    n = 777
    histos = rng.random((n, 11, 20)).astype(np.float32)
    offsets = rng.integers(0, 1500, (n, 11)).astype(np.float32)
    h0 = rng.standard_normal((2 * layers, n, hid)).astype(np.float32)
    c0 = rng.standard_normal((2 * layers, n, hid)).astype(np.float32)
    lstm = mt == "attbilstm"
    state = (torch.from_numpy(h0), torch.from_numpy(c0)) if lstm else torch.from_numpy(h0)
    out = m(torch.from_numpy(offsets), torch.from_numpy(histos), h0=state).cpu().numpy()
    ref = aggr_numpy.forward(sd, offsets, histos, (h0, c0) if lstm else h0, num_layers=layers, dtype=np.float64)
    assert np.abs(out - ref).max() <= 2e-5
    if lstm:
        with pytest.raises(ValueError):
            m(torch.from_numpy(offsets), torch.from_numpy(histos), h0=torch.from_numpy(h0))


def test_set_weight_rejects_wrong_shape(model):
    import ctypes
    from ccsmeth_b200 import _lib
    lib = _lib.load()
    a = np.zeros((3, 3), dtype=np.float32)
    shp = (ctypes.c_int64 * 2)(3, 3)
    rc = lib.ccsm_set_weight(model._handle, b"fc1.weight", a.ctypes.data_as(ctypes.c_void_p), shp, 2)
    assert rc == _lib.EKEY and b"size mismatch" in lib.ccsm_last_error()
    rc = lib.ccsm_set_weight(model._handle, b"nonexistent.weight", a.ctypes.data_as(ctypes.c_void_p), shp, 2)
    assert rc == _lib.EKEY


def test_aggr_loop_matches_reference_loop(ckpt_aggr, golden_aggr):
    """_cal_modfreq_in_aggregate_mode vs the reference's own loop output (seeded randn h0 per 1024-slice)."""
    from ccsmeth_b200.models import AggrAttRNN
    from ccsmeth_b200.call_mods_freq_bam import _cal_modfreq_in_aggregate_mode
    g = golden_aggr
    m = AggrAttRNN(11, 1, 1, 0, 32, binsize=20, model_type="attbigru", device=0)
    m.load_state_dict({k: torch.from_numpy(v) for k, v in ckpt_aggr.items()})
    m = m.cuda(0).eval()
    torch.manual_seed(int(g["tseed"]))
    probs = _cal_modfreq_in_aggregate_mode(g["pos"], list(g["histos"]), m, 11, False)
    assert len(probs) == len(g["loop_probs"])
    assert np.abs(np.array(probs) - g["loop_probs"]).max() <= 2e-6


def test_lstm_variant_matches_reference():
    """ModelAttRNN(model_type="attbilstm2s") on the fp32 kernels (ccsm_forward_att2s_lstm) vs the reference's own forward
    (fixture att2s_lstm.npz: seeded random weights, hidden 64, 2 layers), with explicit (h0, c0) and with the
    reference's seeded draw order."""
    from ccsmeth_b200.models import ModelAttRNN
    from tests.conftest import load_npz
    g = load_npz("att2s_lstm.npz")
    m = ModelAttRNN(21, 2, 2, 0, 64, is_npass=True, model_type="attbilstm2s", device=0)
    m.load_state_dict({k[3:]: torch.from_numpy(v) for k, v in g.items() if k.startswith("sd.")})
    m = m.cuda(0).eval()
    assert m.get_precision() == "fp32" and m.rnn_cell == "lstm"
    a = args16(g)
    hc = ((torch.from_numpy(g["h0_f"]), torch.from_numpy(g["c0_f"])), (torch.from_numpy(g["h0_r"]), torch.from_numpy(g["c0_r"])))
    logits, probs = m(*a, h0=hc)
    assert np.abs(probs.cpu().numpy() - g["probs"]).max() <= 1e-5
    assert np.abs(logits.cpu().numpy() - g["logits"]).max() <= 1e-4
    torch.manual_seed(int(g["seed"]))
    _, probs = m(*a)
    assert np.abs(probs.cpu().numpy() - g["probs_seeded"]).max() <= 1e-5
    m.set_precision("bf16")  # ignored: the tensor-core kernels implement the GRU cell
    _, probs = m(*a, h0=hc)
    assert np.abs(probs.cpu().numpy() - g["probs"]).max() <= 1e-5


@pytest.mark.parametrize("cell,mt", [("gru", "attbigru2s2"), ("lstm", "attbilstm2s2")])
def test_2s2_variants_match_reference(cell, mt):
    """ModelAttRNN2 on the fp32 kernels (integer kinetics / pass-count embeddings, two-layer classifier) vs the
    reference's own forward (fixture att2s2.npz: seeded random weights, hidden 32, 2 layers)."""
    from ccsmeth_b200.models import ModelAttRNN2
    from tests.conftest import load_npz
    g = load_npz("att2s2.npz")
    m = ModelAttRNN2(21, 2, 2, 0, 32, is_npass=True, model_type=mt, device=0)
    m.load_state_dict({k[len(cell) + 4:]: torch.from_numpy(v) for k, v in g.items() if k.startswith(cell + ".sd.")})
    m = m.cuda(0).eval()
    t = lambda k: torch.from_numpy(g[k])
    h = (t(cell + ".h0"), t(cell + ".h1")) if cell == "gru" else ((t(cell + ".h0"), t(cell + ".h1")), (t(cell + ".h2"), t(cell + ".h3")))
    logits, probs = m(*args16(g), h0=h)
    assert np.abs(probs.cpu().numpy() - g[cell + ".probs"]).max() <= 1e-5
    assert np.abs(logits.cpu().numpy() - g[cell + ".logits"]).max() <= 1e-4


def test_transencoder_matches_reference():
    """ModelTransEnc on the fp32 kernels vs the reference's own forward (fixture transenc.npz: seeded random weights,
    d_model 64, 4 heads, dim_ff 128, 2 layers, randomised BatchNorm statistics), incl. ragged sizes."""
    from ccsmeth_b200.models import ModelTransEnc
    from tests.conftest import load_npz
    g = load_npz("transenc.npz")
    m = ModelTransEnc(21, 2, 2, 0, 64, 4, 128, is_npass=True, device=0)
    d = m.state_dict()
    d.update({k[3:]: torch.from_numpy(v) for k, v in g.items() if k.startswith("sd.")})
    m.load_state_dict(d)
    m = m.cuda(0).eval()
    logits, probs = m(*args16(g))
    assert np.abs(probs.cpu().numpy() - g["probs"]).max() <= 1e-5
    assert np.abs(logits.cpu().numpy() - g["logits"]).max() <= 1e-4
    sub = {k: v[:5] for k, v in g.items() if not k.startswith("sd.") and k not in ("logits", "probs")}
    _, p5 = m(*args16(sub))
    assert np.abs(p5.cpu().numpy() - g["probs"][:5]).max() <= 1e-5


@pytest.mark.parametrize("only_close", [False, True])
def test_aggr_in_kernel_windows_equal_materialised_windows(ckpt_aggr, golden_aggr, only_close):
    """ccsm_forward_aggr_sites (the kernel gathers each site's neighbourhood) == ccsm_forward_aggr on the windows
    call_mods_freq_bam.site_windows builds, including the reference's only_close flags."""
    from ccsmeth_b200.models import AggrAttRNN
    from ccsmeth_b200.call_mods_freq_bam import site_windows
    m = AggrAttRNN(11, 1, 1, 0, 32, binsize=20, model_type="attbigru", device=0)
    m.load_state_dict({k: torch.from_numpy(v) for k, v in ckpt_aggr.items()})
    m = m.cuda(0).eval()
    rng = np.random.default_rng(3)
    n = 3000
    pos = (np.cumsum(rng.integers(1, 3, size=n)) * 2).astype(np.int64)
    rows = rng.random((n, 20)).astype(np.float32)
    h0 = torch.randn(2, n, 32)
    off, win = site_windows(pos, rows, 11, only_close)
    a = m(torch.from_numpy(off), torch.from_numpy(win), h0=h0).cpu().numpy()
    b = m.forward_sites(pos, rows, h0=h0, only_close=only_close).cpu().numpy()
    assert np.abs(a - b).max() <= 1e-6
