"""The batch loop replacement vs the reference's _call_mods2s output (fixture att2s_batchloop.npz)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _feature_list(g):
    code2base = "ACGTN"
    n = g["kmer"].shape[0]
    rows = []
    for i in range(n):
        fk = "".join(code2base[int(c)] for c in g["kmer"][i])
        rk = "".join(code2base[int(c)] for c in g["kmer2"][i])
        rows.append((".", -1, ".", "hole%d" % (i // 100), i * 7 + 3,
                     fk, int(g["kpass"][i, 0]), g["ipd"][i].astype(np.float64), ".", g["pw"][i].astype(np.float64), ".",
                     ".", ".", rk, int(g["kpass2"][i, 0]), g["ipd2"][i].astype(np.float64), ".",
                     g["pw2"][i].astype(np.float64), ".", ".", ".", 1))
    return rows


def test_call_mods2s_matches_reference_loop(ckpt_att2s, golden_batchloop):
    from ccsmeth_b200.models import ModelAttRNN
    from ccsmeth_b200 import call_modifications as cm
    g = golden_batchloop
    m = ModelAttRNN(21, 3, 2, 0, 256, is_npass=True, model_type="attbigru2s", device=0, precision="fp32")
    m.load_state_dict({k: torch.from_numpy(v) for k, v in ckpt_att2s.items()})
    m = m.cuda(0).eval()
    fb = cm._batch_feature_list2s(_feature_list(g))
    torch.manual_seed(int(g["tseed"]))  # same seed => same chunk-ordered h0 stream as the reference run
    pred, nb = cm._call_mods2s(fb, m, int(g["batch_size"]), 0)
    assert nb == int(g["batch_num"]) == 3
    assert [p[0] for p in pred] == list(g["holeids"]) and [p[1] for p in pred] == list(g["locs"])
    prob = np.array([p[2] for p in pred], dtype=np.float32)
    assert np.abs(prob - g["prob1"]).max() <= 1e-4
    assert np.abs(prob * 1e6 - np.round(prob * 1e6)).max() < 0.51  # 6-decimal rounding kept
