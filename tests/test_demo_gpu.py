"""Config 3: demo BAM end to end on one B200 (BAM -> features -> kernels -> MM/ML -> modbam) against the
reference chain's per-site probabilities and MM/ML tags (fixture demo_callmods.npz, same tseed => same h0)."""
import os
from collections import OrderedDict

import numpy as np
import pytest
import torch

from ccsmeth_b200 import call_mods as cm
from ccsmeth_b200.bamio import BamReader
from tests.conftest import GOLDEN, load_npz

pytestmark = pytest.mark.gpu
DEMO = os.path.join(GOLDEN, "demo", "hg002.chr20_demo.hifi.bam")


@pytest.fixture(scope="module")
def ckpt_file(tmp_path_factory, ckpt_att2s):
    p = str(tmp_path_factory.mktemp("ckpt") / "model_v3.ckpt")
    torch.save(OrderedDict((k, torch.from_numpy(v)) for k, v in ckpt_att2s.items()), p)
    return p


@pytest.mark.parametrize("prec,max_ml_flips", [("fp16c8", 12), ("fp16x3", 12), ("fp32", 12), ("bf16x3", 20)])
def test_call_mods_demo_matches_reference_chain(tmp_path, ckpt_file, prec, max_ml_flips):
    g = load_npz("demo_callmods.npz")
    out = str(tmp_path / ("demo_" + prec))
    args = cm.build_parser().parse_args(["-i", DEMO, "-m", ckpt_file, "-o", out, "--mode", "denovo",
                                         "--tseed", str(int(g["tseed"])), "--precision", prec])
    counts, path = cm.call_mods(args)
    assert counts == {"sites": 12691, "model_batches": 26, "reads_written": 116,
                      "reads_with_mm": int((g["n_sites_per_read"] > 0).sum())}
    recs = list(BamReader(path))
    assert [r.query_name for r in recs] == list(g["names"])
    off, flips, n_tot = 0, 0, 0
    for r, n in zip(recs, g["n_sites_per_read"]):
        if n == 0:
            assert not r.has_tag("MM")
            continue
        mm = r.get_tag("MM")
        assert [int(x) for x in mm[5:-1].split(",")] == list(g["mm"][off:off + n])  # same sites called
        ml = r.get_tag("ML")
        flips += int((ml != g["ml"][off:off + n]).sum())
        # an ML byte may differ only where a <=1e-4 probability difference straddles a k/256 edge
        assert np.abs(ml.astype(int) - g["ml"][off:off + n].astype(int)).max() <= 1
        n_tot += n
        off += n
    assert n_tot == 12691 and flips <= max_ml_flips, flips


def test_call_holebatch_probabilities(ckpt_file):
    g = load_npz("demo_callmods.npz")
    args = cm.build_parser().parse_args(["-i", DEMO, "-m", ckpt_file, "-o", "x", "--precision", "fp16x3"])
    model = cm.load_model(ckpt_file, args, device=0, precision="fp16x3")
    reads = list(BamReader(DEMO))
    torch.manual_seed(int(g["tseed"]))
    probs = []
    for b0 in range(0, len(reads), 50):
        per_read, n, nb = cm.call_holebatch(model, reads[b0:b0 + 50], ["CG"], args)
        for pr in per_read:
            if pr is not None:
                order = np.argsort(pr[0], kind="stable")
                probs.append(pr[1][order])
    probs = np.concatenate(probs)
    assert probs.shape == g["prob1"].shape
    assert np.abs(probs - g["prob1"]).max() <= 1e-4  # north-star tolerance, whole demo, reference h0 stream


def _tags_by_name(path):
    out = {}
    for r in BamReader(path):
        out[r.query_name] = (r.get_tag("MM"), r.get_tag("ML").tobytes()) if r.has_tag("MM") else None
    return out


def test_pieces_and_rank_shards_give_the_same_modbam(tmp_path, ckpt_file, monkeypatch):
    """The native piece pipeline must not depend on where the file is cut (--device_batch) nor on how hole-batches
    are dealt to ranks: same MM/ML per read.  h0 = zeros so that every configuration sees the same initial state."""
    base = ["-i", DEMO, "-m", ckpt_file, "--precision", "fp16x3", "--h0", "zeros", "--holes_batch", "10"]
    one = cm.build_parser().parse_args(base + ["-o", str(tmp_path / "one")])
    counts1, p1 = cm.call_mods(one)
    ref = _tags_by_name(p1)
    assert counts1["sites"] == 12691 and counts1["reads_written"] == 116
    many = cm.build_parser().parse_args(base + ["-o", str(tmp_path / "many"), "--device_batch", "1"])
    counts2, p2 = cm.call_mods(many)  # 10 reads x 64 KiB per piece -> several pieces
    assert counts2 == counts1
    assert _tags_by_name(p2) == ref
    # two "ranks" run one after the other in this process (no process group: counts are per rank)
    from ccsmeth_b200 import parallel
    merged, sites = {}, 0
    for rank in (0, 1):
        monkeypatch.setattr(parallel, "init_from_env", lambda r=rank: (r, 2, 0))
        monkeypatch.setattr(parallel, "allreduce_counts", lambda c: list(c))
        a = cm.build_parser().parse_args(base + ["-o", str(tmp_path / "shard"), "--device_batch", "2", "--no_sort"])
        c, p = cm.call_mods(a)
        assert p.endswith(".rank%d.modbam.bam" % rank)
        merged.update(_tags_by_name(p))
        sites += c["sites"]
    assert sites == 12691 and merged == ref


@pytest.mark.parametrize("mt,extra", [("transencoder2s", ["--layer_trans", "2", "--d_model", "64", "--nhead", "4", "--dim_ff", "128"]),
                                      ("attbigru2s2", ["--layer_rnn", "2", "--hid_rnn", "32", "--h0", "zeros"]),
                                      ("attbilstm2s", ["--layer_rnn", "2", "--hid_rnn", "32", "--h0", "zeros", "--norm", "zscore"])])
def test_call_mods_pipeline_runs_the_other_model_types(tmp_path, mt, extra):
    """--model_type dispatch through the BAM pipeline (reference call_modifications.py:315-340): seeded random weights
    (no checkpoint ships for these types); the ML bytes written to the modbam must equal the ones obtained by calling
    the model's 16-tensor forward on the device-extracted features."""
    import ctypes
    from ccsmeth_b200.extract_features import extract_opts, pack_reads
    torch.manual_seed(11)
    base = ["-i", DEMO, "-o", str(tmp_path / mt), "--model_type", mt] + (["--norm", "none"] if "--norm" not in extra else []) + extra
    args = cm.build_parser().parse_args(base + ["-m", "unused"])
    from ccsmeth_b200.models import ModelAttRNN, ModelAttRNN2, ModelTransEnc
    if mt == "transencoder2s":
        ref_model = ModelTransEnc(21, 2, 2, 0, 64, 4, 128)
    elif mt == "attbigru2s2":
        ref_model = ModelAttRNN2(21, 2, 2, 0, 32, model_type=mt)
    else:
        ref_model = ModelAttRNN(21, 2, 2, 0, 32, model_type=mt)
    ckpt = str(tmp_path / (mt + ".ckpt"))
    torch.save(ref_model.state_dict(), ckpt)
    args.model_file = ckpt
    counts, path = cm.call_mods(args)
    assert counts["sites"] == 12691 and counts["reads_written"] == 116
    recs = list(BamReader(path))
    # independent route: device features of the first reads -> model.forward -> prob1 -> ML
    model = cm.load_model(ckpt, args, device=0)
    model.set_h0_mode("zeros")
    sub = list(BamReader(DEMO))[:12]
    batch = pack_reads(sub, args)
    n = model.extract_reads(batch, extract_opts(args, ["CG"]))
    site_read, _ = model.reads_sites()
    f = model.reads_features()
    order = ("kmer", "kpass", "ipd", None, "pw", None, None, None)
    a = [f[k] if k else torch.zeros(1) for k in order] + [f[k + "2"] if k else torch.zeros(1) for k in order]
    kw = {}
    if mt == "attbigru2s2":
        kw["h0"] = (torch.zeros(4, n, 32), torch.zeros(4, n, 32))
    elif mt == "attbilstm2s":
        z = torch.zeros(4, n, 32)
        kw["h0"] = ((z, z), (z, z))
    _, probs = model(*a, **kw)
    p = probs.cpu().numpy()
    ml = cm.convert_probs_to_mltag(np.round(p[:, 1] / (p[:, 0] + p[:, 1]), 6))
    for r in np.unique(site_read):
        assert np.array_equal(recs[batch.index[r]].get_tag("ML"), ml[site_read == r]), (mt, r)
