"""Row 8f-3 on the GPU: region pileup -> per-site (coverage, modified count, frequency) through ccsm_pileup_* and the
fused aggregate kernel with in-kernel windows, against the reference's own region caller (fixture pileup_region.npz,
same h0) and the numpy oracle on edge cases.  Integer outputs and all count-mode values must be identical; the
aggregate frequency is a float32 model output rounded to 6 decimals (tolerance 2e-6, as for the windowed forward)."""
import argparse

import numpy as np
import pytest
import torch

from ccsmeth_b200.call_mods_freq_bam import _call_modfreq_of_one_region
from ccsmeth_b200.models import AggrAttRNN
from oracle import pileup_numpy
from tests.conftest import load_npz

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def g():
    return load_npz("pileup_region.npz")


@pytest.fixture(scope="module")
def model(ckpt_aggr):
    m = AggrAttRNN(11, 1, 1, 0, 32, binsize=20, model_type="attbigru", device=0)
    m.load_state_dict({k: torch.from_numpy(v) for k, v in ckpt_aggr.items()})
    return m.cuda(0).eval()


def _args(**kw):
    a = argparse.Namespace(call_mode="aggregate", cov_cf=4, prob_cf=0.0, no_amb_cov=False, no_hap=False)
    for k, v in kw.items():
        setattr(a, k, v)
    return a


def _info(g):
    return {int(g["pos"][i]): [(int(g["ml"][k]), int(g["hap"][k])) for k in range(g["ptr"][i], g["ptr"][i + 1])]
            for i in range(len(g["pos"]))}


def _pack(res, n):
    out = np.full((3, n, 3), np.nan)
    for i, r in enumerate(res):
        for grp in range(3):
            if r[1 + grp] is not None:
                out[grp, i] = r[1 + grp]
    return out


def _same(a, b, tol=0.0):
    assert np.array_equal(np.isnan(a), np.isnan(b))
    m = ~np.isnan(a)
    if m.any():
        assert np.abs(a[m] - b[m]).max() <= tol


@pytest.mark.parametrize("tag,kw", [("count", {}), ("count_cf3", {"prob_cf": 0.3}),
                                    ("count_cf3_noamb", {"prob_cf": 0.3, "no_amb_cov": True})])
def test_count_mode_is_identical_to_the_reference(g, model, tag, kw):
    res = _call_modfreq_of_one_region(_info(g), _args(call_mode="count", **kw), model)
    assert [r[0] for r in res] == list(g["pos"])
    _same(_pack(res, len(g["pos"])), g[tag])


@pytest.mark.parametrize("tag,kw", [("aggr", {}), ("aggr_cf3", {"prob_cf": 0.3}), ("aggr_nohap", {"no_hap": True})])
def test_aggregate_mode_matches_the_reference_region_caller(g, model, tag, kw):
    sizes = g[tag + "_h0_sizes"]
    n_high = model.pileup_begin(g["pos"], g["ptr"], g["ml"], g["hap"], call_mode="aggregate", cov_cf=4,
                                prob_cf=kw.get("prob_cf", 0.0), no_hap=kw.get("no_hap", False))
    # the reference drew one h0 per 1024-site slice, group after group: cut the recorded stream the same way
    assert sum(-(-nh // 1024) for nh in n_high) == len(sizes) and sum(n_high) == g[tag + "_h0"].shape[1]
    h0, off = [], 0
    for nh in n_high:
        h0.append(torch.from_numpy(np.ascontiguousarray(g[tag + "_h0"][:, off:off + nh])) if nh else None)
        off += nh
    res = _call_modfreq_of_one_region(_info(g), _args(**kw), model, h0=h0)
    out, ref = _pack(res, len(g["pos"])), g[tag]
    _same(out[..., 0], ref[..., 0])              # coverage: identical
    _same(out[..., 2], ref[..., 2], tol=2e-6)    # frequency
    _same(out[..., 1], ref[..., 1], tol=0.0101)  # round(cov * freq, 2) may move by one step of 0.01
    low = np.diff(g["ptr"]) < 4                  # low-coverage sites take the count path: identical
    _same(out[0][low], ref[0][low])


def test_edge_regions_against_the_oracle(model, ckpt_aggr):
    rng = np.random.default_rng(5)
    cases = {
        "single_high": (np.array([100]), [9]),
        "single_low": (np.array([100]), [2]),
        "all_low": (np.cumsum(rng.integers(2, 50, 40)), rng.integers(1, 4, 40)),
        "short_high": (np.cumsum(rng.integers(2, 300, 7)), rng.integers(4, 30, 7)),   # fewer sites than one window
        "mixed": (np.cumsum(rng.integers(2, 2000, 1500)), rng.integers(1, 70, 1500)),
    }
    for name, (pos, cov) in cases.items():
        pos = pos.astype(np.int64)
        ptr = np.concatenate(([0], np.cumsum(cov))).astype(np.int64)
        ml = rng.integers(0, 256, int(ptr[-1])).astype(np.uint8)
        hap = rng.integers(0, 3, int(ptr[-1])).astype(np.uint8)
        for mode in ("count", "aggregate"):
            n_high = model.pileup_begin(pos, ptr, ml, hap, call_mode=mode, cov_cf=4, prob_cf=0.2)
            h0 = [None if nh == 0 else rng.standard_normal((2, nh, 32)).astype(np.float32) for nh in n_high]
            cov_d, cnt_d, freq_d = model.pileup_finish([None if h is None else torch.from_numpy(h) for h in h0])
            ref = pileup_numpy.call_region(pos, ptr, ml, hap, ckpt_aggr, call_mode=mode, cov_cf=4, prob_cf=0.2, h0=h0)
            none = cov_d < 0
            assert np.array_equal(none, np.isnan(ref[..., 0])), (name, mode)
            m = ~none
            assert np.array_equal(cov_d[m], ref[..., 0][m].astype(np.int32)), (name, mode)
            assert np.abs(freq_d[m] - ref[..., 2][m]).max() <= 2e-6, (name, mode)
            assert np.abs(cnt_d[m] - ref[..., 1][m]).max() <= 0.0101, (name, mode)
    # --only_close: the 21st input becomes the "directly follows its predecessor" flag
    pos = np.cumsum(rng.choice([2, 2, 3, 40, 700], size=900)).astype(np.int64)
    cov = rng.integers(4, 40, 900)
    ptr = np.concatenate(([0], np.cumsum(cov))).astype(np.int64)
    ml = rng.integers(0, 256, int(ptr[-1])).astype(np.uint8)
    n_high = model.pileup_begin(pos, ptr, ml, None, call_mode="aggregate", only_close=True)
    h0 = rng.standard_normal((2, n_high[0], 32)).astype(np.float32)
    _, _, freq_d = model.pileup_finish([torch.from_numpy(h0), None, None])
    ref = pileup_numpy.call_region(pos, ptr, ml, np.zeros(len(ml), np.uint8), ckpt_aggr, h0=(h0, None, None), only_close=True)
    assert np.abs(freq_d[0] - ref[0][:, 2]).max() <= 2e-6
    n_high = model.pileup_begin(pos, ptr, ml, None, call_mode="aggregate", only_close=False)
    _, _, freq_far = model.pileup_finish([torch.from_numpy(h0), None, None])
    assert np.abs(freq_far[0] - freq_d[0]).max() > 1e-3  # the flag really changes the model input
    assert model.pileup_begin(np.zeros(0, np.int64), np.zeros(1, np.int64), np.zeros(0, np.uint8)) == (0, 0, 0)
    assert model.pileup_finish()[0].shape == (3, 0)
    # hap = None: everything is haplotype 0, the hp groups are empty
    n_high = model.pileup_begin(np.array([5, 9], np.int64), np.array([0, 5, 11], np.int64), rng.integers(0, 256, 11).astype(np.uint8))
    assert n_high == (2, 0, 0)
    cov_d, _, _ = model.pileup_finish()
    assert list(cov_d[0]) == [5, 6] and (cov_d[1:] == -1).all()


@pytest.mark.parametrize("tag,sdtag,mt,hid,layers,kw", [
    ("lstm", "lstm", "attbilstm", 32, 1, {}),
    ("gru48x2", "gru48x2", "attbigru", 48, 2, {}),
    ("lstm_close", "lstm", "attbilstm", 32, 1, {"only_close": True}),
])
def test_other_aggregate_models_match_the_reference_region_caller(tag, sdtag, mt, hid, layers, kw):
    """--model_type attbilstm and --hid_rnn 48 --layer_rnn 2 (outside the fused kernel: materialised windows + the
    layer-by-layer fp32 kernels) against the reference's region caller on seeded random weights, same (h0, c0)."""
    v = load_npz("aggr_variants.npz")
    m = AggrAttRNN(11, layers, 1, 0, hid, binsize=20, model_type=mt, device=0)
    m.load_state_dict({k[len(sdtag) + 4:]: torch.from_numpy(v[k]) for k in v if k.startswith(sdtag + ".sd.")})
    m = m.cuda(0).eval()
    n_high = m.pileup_begin(v["pos"], v["ptr"], v["ml"], v["hap"], call_mode="aggregate", cov_cf=4, **kw)
    assert sum(n_high) == v[sdtag + "_h0"].shape[1]
    states, off = [], 0
    for nh in n_high:
        h0 = torch.from_numpy(np.ascontiguousarray(v[sdtag + "_h0"][:, off:off + nh]))
        if mt == "attbilstm":
            states.append((h0, torch.from_numpy(np.ascontiguousarray(v[sdtag + "_c0"][:, off:off + nh]))))
        else:
            states.append(h0)
        off += nh
    res = _call_modfreq_of_one_region(_info(v), _args(**kw), m, h0=states)
    out, ref = _pack(res, len(v["pos"])), v[tag]
    _same(out[..., 0], ref[..., 0])
    _same(out[..., 2], ref[..., 2], tol=2e-6)
    _same(out[..., 1], ref[..., 1], tol=0.0101)
    # the reference's own draw order reproduced from the seed: h0 (then c0) per 1024-site slice, group after group
    if not kw:
        from ccsmeth_b200.call_freqb import draw_region_h0
        a = argparse.Namespace(tseed=1234, seq_len=11, layer_rnn=layers, class_num=1, hid_rnn=hid, bin_size=20, model_type=mt)
        drawn = draw_region_h0(a, n_high)
        first = drawn[0][0] if mt == "attbilstm" else drawn[0]
        assert np.array_equal(first.numpy(), v[sdtag + "_h0"][:, :n_high[0]])


def test_windows_path_agrees_with_the_fused_kernel(g, model, monkeypatch):
    """The shipped model through the generic path (CCSM_AGGR_UNFUSED is honoured by the windowed forward only, so build
    the windows on the host and compare with the in-kernel windows of the pileup call)."""
    from oracle import aggr_numpy
    n_high = model.pileup_begin(g["pos"], g["ptr"], g["ml"], g["hap"], call_mode="aggregate", no_hap=True)
    h0 = torch.randn(2, n_high[0], 32, generator=torch.Generator().manual_seed(3))
    _, _, freq = model.pileup_finish([h0, None, None])
    cov = np.diff(g["ptr"])
    hi = np.nonzero(cov >= 4)[0]
    histos = [pileup_numpy.normalized_histo([pileup_numpy.cal_mod_prob(int(x)) for x in g["ml"][g["ptr"][i]:g["ptr"][i + 1]]])
              for i in hi]
    pm, hm = aggr_numpy.build_windows(g["pos"][hi], histos)
    monkeypatch.setenv("CCSM_AGGR_UNFUSED", "1")
    raw = model(torch.from_numpy(np.ascontiguousarray(pm, dtype=np.float32)),
                torch.from_numpy(np.ascontiguousarray(hm, dtype=np.float32)), h0=h0).cpu().numpy()
    assert np.abs(aggr_numpy.postprocess(raw)[:, 0] - freq[0][hi]).max() <= 2e-6
