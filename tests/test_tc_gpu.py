"""Tensor-core (tcgen05) path vs the reference fixtures / numpy oracle, through the C ABI."""
import ctypes

import numpy as np
import pytest
import torch

from oracle import att2s_numpy
from tests.test_parity_gpu import args16, FEATS, EDGE_CASES

pytestmark = pytest.mark.gpu

# parity modes must meet the north-star tolerance; single-pass modes are reported and only
# sanity-bounded here
TOL = {"fp16c8": 1e-4, "fp16x3": 1e-4, "bf16x3": 1e-4, "fp16": 5e-3, "bf16": 3e-2}


@pytest.fixture(scope="module")
def model(ckpt_att2s):
    from ccsmeth_b200.models import ModelAttRNN
    m = ModelAttRNN(21, 3, 2, 0, 256, is_npass=True, model_type="attbigru2s", device=0, precision="fp16x3")
    m.load_state_dict({k: torch.from_numpy(v) for k, v in ckpt_att2s.items()})
    return m.cuda(0).eval()


def run(model, g, pfx=""):
    h0 = (torch.from_numpy(g[pfx + "h0_f"]), torch.from_numpy(g[pfx + "h0_r"]))
    logits, probs = model(*[a.cuda() for a in args16(g, pfx)], h0=h0)
    return logits.cpu().numpy(), probs.cpu().numpy()


def layer_out(model, layer, n):
    from ccsmeth_b200 import _lib
    tiles = (2 * n + 127) // 128
    tiles += tiles & 1
    buf = np.empty(tiles * 128 * 21 * 512, dtype=np.float32)
    got = _lib.load().ccsm_debug_tc_layer_out(model._handle, layer, buf.ctypes.data_as(ctypes.c_void_p), buf.size)
    assert got == buf.size, got
    return buf.reshape(tiles * 128, 21, 512)[:2 * n].reshape(n, 2, 21, 512)


@pytest.mark.parametrize("prec", ["fp16c8", "fp16x3", "bf16x3", "fp16", "bf16"])
def test_tc_matches_reference_synth(model, golden_synth, prec):
    model.set_precision(prec)
    _, probs = run(model, golden_synth)
    err = np.abs(probs - golden_synth["probs"]).max()
    print("tc %s max|dprob| = %.3e" % (prec, err))
    assert err <= TOL[prec]


@pytest.mark.parametrize("prec", ["fp16c8", "fp16x3", "bf16x3"])
def test_tc_layers_match_oracle(model, ckpt_att2s, golden_synth, prec):
    model.set_precision(prec)
    n = 64
    g = {k: (v[:, :n] if k.startswith("h0") else v[:n]) for k, v in golden_synth.items()}
    run(model, g)
    _, _, it = att2s_numpy.forward(ckpt_att2s, *[g[k] for k in FEATS], g["h0_f"], g["h0_r"], return_internals=True)
    for l in range(3):
        out = layer_out(model, l, n)
        for s in range(2):
            err = np.abs(out[:, s] - it["layers%d" % s][l]).max()
            assert err <= 2e-4, (l, s, err)


@pytest.mark.parametrize("prec", ["fp16c8", "fp16x3"])
@pytest.mark.parametrize("case", EDGE_CASES)
def test_tc_edge_cases(model, golden_edge, case, prec):
    model.set_precision(prec)
    _, probs = run(model, golden_edge, case + ".")
    assert probs.shape == golden_edge[case + ".probs"].shape
    assert np.abs(probs - golden_edge[case + ".probs"]).max() <= 1e-4


@pytest.mark.parametrize("prec", ["fp16c8", "fp16x3"])
def test_tc_multi_tile_and_chunks(model, golden_synth, prec):
    """More sites than one pair of row tiles and than one library chunk (148*8*64 sites): every replica of the
    256 golden sites must reproduce the reference."""
    model.set_precision(prec)
    g = golden_synth
    rep = 300  # 76,800 sites > 75,776 per chunk
    big = {k: np.concatenate([g[k]] * rep, axis=1 if k.startswith("h0") else 0) for k in FEATS + ("h0_f", "h0_r")}
    _, probs = run(model, big)
    probs = probs.reshape(rep, 256, 2)
    assert np.abs(probs - g["probs"][None]).max() <= 1e-4


def test_tc_host_entry(model, golden_synth):
    model.set_precision("bf16x3")
    g = golden_synth
    _, probs = model.forward_host({k: g[k] for k in FEATS}, h0=(g["h0_f"], g["h0_r"]))
    assert np.abs(probs.numpy() - g["probs"]).max() <= 1e-4


def test_device_h0_same_noise_in_every_mode(model, golden_synth):
    """CCSM_H0_DEVICE_RANDOM: the Philox stream is defined per (site, strand, layer, dir, unit), so the fp32
    path and the tensor-core paths must see the same h0 and agree to parity tolerance; a second call draws
    fresh noise; zeros mode equals an explicit zero h0."""
    g = golden_synth
    a = [x.cuda() for x in args16(g)]
    outs = {}
    for prec in ("fp32", "fp16x3", "bf16x3", "fp16c8"):
        model.set_precision(prec)
        model.set_h0_mode("device", seed=77)
        _, p1 = model(*a)
        _, p2 = model(*a)
        outs[prec] = (p1.cpu().numpy(), p2.cpu().numpy())
    assert np.abs(outs["fp32"][0] - outs["fp16x3"][0]).max() <= 1e-4
    assert np.abs(outs["fp32"][0] - outs["bf16x3"][0]).max() <= 1e-4
    assert np.abs(outs["fp32"][0] - outs["fp16c8"][0]).max() <= 1e-4
    assert np.abs(outs["fp32"][1] - outs["fp16x3"][1]).max() <= 1e-4
    assert np.abs(outs["fp32"][0] - outs["fp32"][1]).max() > 1e-2       # fresh noise on every call
    assert np.abs(outs["fp32"][0] - g["probs"]).max() > 1e-2            # and not the fixture's h0
    # the spread over h0 draws is what the reference itself shows (mean 0.05)
    assert 0.005 < np.abs(outs["fp32"][0][:, 1] - outs["fp32"][1][:, 1]).mean() < 0.2
    model.set_precision("fp16x3")
    model.set_h0_mode("zeros")
    _, pz = model(*a)
    z = torch.zeros(6, 256, 256)
    _, pe = model(*a, h0=(z, z))
    assert np.abs(pz.cpu().numpy() - pe.cpu().numpy()).max() == 0.0
    model.set_h0_mode("reference")


@pytest.mark.parametrize("prec", ["fp16c8", "fp16x3", "bf16x3", "fp32"])
def test_parity_on_16k_reference_sites(model, prec):
    """16,384 synthetic sites (regenerated from the seed) against the unmodified reference's outputs
    (tests/golden/att2s_synth16k.npz, scripts/gen_golden.py att2s_16k): every parity mode within 1e-4 on every site."""
    from ccsmeth_b200 import synth
    from tests.conftest import load_npz
    g = load_npz("att2s_synth16k.npz")
    n = int(g["n"])
    b = synth.make_batch(n, seed=int(g["seed"]))
    chk = float(sum(float(v.double().sum()) for v in b.values()))
    assert abs(chk - float(g["input_checksum"])) < 1e-6 * abs(chk), "the input generator changed: regenerate the fixture"
    model.set_precision(prec)
    _, probs = model(*[a.cuda() for a in synth.to_forward_args(b)], h0=(b["h0_f"], b["h0_r"]))
    d = np.abs(probs.cpu().numpy() - g["probs"])
    print("%s on %d sites: max|dprob| %.3e mean %.3e" % (prec, n, d.max(), d.mean()))
    assert d.max() <= 1e-4
    model.set_precision("fp16x3")
