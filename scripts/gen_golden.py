#!/usr/bin/env python
"""Generate the committed parity fixtures under tests/golden/ by running the UNMODIFIED
reference (imported read-only from /root/reference with stubbed I/O deps, see
oracle/refimport.py) on seeded inputs.  Runs only in the build container.

    python scripts/gen_golden.py

Fixtures
  ckpt_att2s_v3.npz    the shipped v3 checkpoint's 30 tensors (the model the path is defined on)
  ckpt_aggr_v2p.npz    the shipped aggregate checkpoint's 13 tensors ("module." prefix kept)
  att2s_synth.npz      256 synthetic sites (config-2 generator) + explicit h0 -> logits, probs
  att2s_edge.npz       edge cases: n=1, n=3 (ragged tail), all-'N' kmers, huge kinetics, npass 0/200,
                       zero h0; each case -> logits, probs
  att2s_seeded.npz     reference default behaviour: torch.manual_seed(1234) then forward (h0 drawn by
                       the reference itself, strand 1 then strand 2) -> the h0 stream + outputs
  att2s_batchloop.npz  reference _call_mods2s over 1100 sites with batch_size=512 (3 batches incl. a
                       ragged tail), torch.manual_seed(1234): per-site (holeid, loc, prob_1_norm)
  aggr_synth.npz       1500 synthetic pileup sites -> windows, h0, raw outputs, clipped+rounded probs
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import refimport  # noqa: E402
from ccsmeth_b200 import synth  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def save_ckpts():
    sd = torch.load(refimport.V3_CKPT, map_location="cpu")
    np.savez_compressed(os.path.join(OUT, "ckpt_att2s_v3.npz"), **{k: v.numpy() for k, v in sd.items()})
    sd = torch.load(refimport.AGGR_CKPT, map_location="cpu")
    np.savez_compressed(os.path.join(OUT, "ckpt_aggr_v2p.npz"), **{k: v.numpy() for k, v in sd.items()})


def run_ref(model, b, h0_f, h0_r):
    args = synth.to_forward_args(b)
    with refimport.fixed_h0(model, [h0_f.clone().requires_grad_(True), h0_r.clone().requires_grad_(True)]):
        logits, probs = model(*args)
    return logits.detach().numpy(), probs.detach().numpy()


def gen_att2s_16k(n=16384, seed=synth.SEED + 16):
    """A larger parity sample (SURVEY.md 8d asks for a 16k-site slice): the unmodified reference forward over
    synth.make_batch(16384, seed) -- features and the explicit h0 are regenerated from the seed by the test (torch's CPU
    generator is deterministic), so only the reference's outputs are stored (probs float32 + logits float16-safe range
    as float32: 256 KB)."""
    torch.set_num_threads(os.cpu_count())
    m = refimport.load_ref_att2s()
    b = synth.make_batch(n, seed=seed)
    probs = np.empty((n, 2), np.float32)
    logits = np.empty((n, 2), np.float32)
    for s in range(0, n, 2048):   # the reference's forward has no chunking of its own; 2048-site calls bound the memory
        sl = {k: (v[:, s:s + 2048] if k.startswith("h0") else v[s:s + 2048]) for k, v in b.items()}
        lg, pr = run_ref(m, sl, sl["h0_f"].contiguous(), sl["h0_r"].contiguous())
        logits[s:s + 2048], probs[s:s + 2048] = lg, pr
    # a checksum of the regenerated inputs guards against a generator change going unnoticed
    chk = float(sum(float(v.double().sum()) for v in b.values()))
    np.savez_compressed(os.path.join(OUT, "att2s_synth16k.npz"), probs=probs, logits=logits, n=n, seed=seed, input_checksum=chk)
    print("att2s_synth16k: prob1 mean %.4f, input checksum %.6f" % (probs[:, 1].mean(), chk))


def gen_att2s():
    torch.set_num_threads(8)
    m = refimport.load_ref_att2s()

    # --- synthetic, explicit h0
    b = synth.make_batch(256, seed=synth.SEED)
    logits, probs = run_ref(m, b, b["h0_f"], b["h0_r"])
    np.savez_compressed(os.path.join(OUT, "att2s_synth.npz"),
                        **{k: v.numpy() for k, v in b.items()}, logits=logits, probs=probs)
    print("att2s_synth: prob1 mean %.4f" % probs[:, 1].mean())

    # --- edge cases
    edge = {}
    cases = {}
    b1 = synth.make_batch(1, seed=7)
    cases["n1"] = b1
    b3 = synth.make_batch(3, seed=8)
    cases["n3"] = b3
    bn = synth.make_batch(5, seed=9)
    bn["kmer"][:] = 4.0
    bn["kmer2"][:] = 4.0
    cases["allN"] = bn
    bk = synth.make_batch(6, seed=10)
    bk["ipd"][:, ::3] = 21.89
    bk["pw"][:, 1::4] = 17.33
    bk["ipd2"][:, 5] = -1.67
    bk["kpass"][:] = 0.0
    bk["kpass2"][:] = 200.0
    cases["extreme"] = bk
    bz = synth.make_batch(4, seed=11)
    bz["h0_f"].zero_()
    bz["h0_r"].zero_()
    cases["zeroh0"] = bz
    bf = synth.make_batch(2, seed=12)
    bf["kmer"] = bf["kmer"] + 0.7  # float codes are truncated by .int() (models.py:91)
    cases["fraccode"] = bf
    for name, bb in cases.items():
        logits, probs = run_ref(m, bb, bb["h0_f"], bb["h0_r"])
        for k, v in bb.items():
            edge[f"{name}.{k}"] = v.numpy()
        edge[f"{name}.logits"] = logits
        edge[f"{name}.probs"] = probs
    np.savez_compressed(os.path.join(OUT, "att2s_edge.npz"), **edge)

    # --- reference default: seeded randn h0 drawn by the reference itself
    b = synth.make_batch(64, seed=21, with_h0=False)
    torch.manual_seed(1234)
    logits, probs = m(*synth.to_forward_args(b))
    torch.manual_seed(1234)
    h0_f = torch.randn(6, 64, 256)
    h0_r = torch.randn(6, 64, 256)
    np.savez_compressed(os.path.join(OUT, "att2s_seeded.npz"),
                        **{k: v.numpy() for k, v in b.items()},
                        h0_f=h0_f.numpy(), h0_r=h0_r.numpy(),
                        logits=logits.detach().numpy(), probs=probs.detach().numpy(), tseed=1234)

    # --- the batch loop: reference _call_mods2s with batch_size 512 over 1100 sites
    ref = refimport.import_reference()
    import ccsmeth.call_modifications as rcm
    n = 1250
    b = synth.make_batch(n, seed=33, with_h0=False)
    code2base = "ACGTN"
    feature_list = []
    for i in range(n):
        fk = "".join(code2base[int(c)] for c in b["kmer"][i])
        rk = "".join(code2base[int(c)] for c in b["kmer2"][i])
        feature_list.append((".", -1, ".", "hole%d" % (i // 100), i * 7 + 3,
                             fk, int(b["kpass"][i, 0]), b["ipd"][i].double().numpy(), ".",
                             b["pw"][i].double().numpy(), ".", ".", ".",
                             rk, int(b["kpass2"][i, 0]), b["ipd2"][i].double().numpy(), ".",
                             b["pw2"][i].double().numpy(), ".", ".", ".", 1))
    fb = rcm._batch_feature_list2s(feature_list)
    torch.manual_seed(1234)
    pred, nb = rcm._call_mods2s(fb, m, 512, 0)
    np.savez_compressed(os.path.join(OUT, "att2s_batchloop.npz"),
                        **{k: v.numpy() for k, v in b.items()},
                        holeids=np.array([p[0] for p in pred]), locs=np.array([p[1] for p in pred]),
                        prob1=np.array([p[2] for p in pred], dtype=np.float32),
                        batch_num=nb, batch_size=512, tseed=1234)
    print("att2s_batchloop: %d preds in %d batches" % (len(pred), nb))


def gen_aggr():
    m = refimport.load_ref_aggr()
    import ccsmeth.call_mods_freq_bam as rfb
    n = 1500
    pos, histos, h0 = synth.make_aggr_batch(n, seed=synth.SEED)
    # direct forward with explicit h0 on materialised windows
    from oracle import aggr_numpy
    pm, hm = aggr_numpy.build_windows(pos, list(histos))
    with refimport.fixed_h0(m, [h0.clone().requires_grad_(True)]):
        raw = m(torch.tensor(pm, dtype=torch.float), torch.tensor(np.array(hm), dtype=torch.float))
    # the reference loop itself (batches of 1024, seeded randn h0)
    torch.manual_seed(1234)
    probs = rfb._cal_modfreq_in_aggregate_mode(pos, list(histos), m, 11, False)
    torch.manual_seed(1234)
    h0_b0 = torch.randn(2, 1024, 32)
    h0_b1 = torch.randn(2, n - 1024, 32)
    np.savez_compressed(os.path.join(OUT, "aggr_synth.npz"), pos=pos, histos=histos, h0=h0.numpy(),
                        pos_mat=pm, raw=raw.detach().numpy(),
                        loop_probs=np.array(probs, dtype=np.float32),
                        loop_h0_b0=h0_b0.numpy(), loop_h0_b1=h0_b1.numpy(), tseed=1234)
    print("aggr_synth: raw mean %.4f" % raw.mean().item())


def demo_args():
    import argparse
    return argparse.Namespace(mode="denovo", seq_len=21, motifs="CG", mod_loc=0, methy_label=1, norm="zscore",
                              no_decode=False, is_sn="no", is_map="no", is_stds="no", is_npass="yes", mapq=1,
                              identity=0.0, no_supplementary=False, skip_unmapped="yes", holes_batch=50,
                              batch_size=512, keep_pulse=False)


def gen_demo():
    """Config 1/3 golden: the reference chain  extract_features_from_double_strand_read ->
    _batch_feature_list2s -> _call_mods2s(batch 512)  over demo/hg002.chr20_demo.hifi.bam (--mode denovo),
    hole-batches of 50 reads, torch.manual_seed(1234) in-process; MM/ML through the reference's own
    _convert_locs_to_mmtag / _convert_probs_to_mltag.  Reads come from ccsmeth_b200.bamio (pysam is absent);
    the reference extractor only touches the duck-typed attributes listed in extract_features.py:88-126."""
    import shutil
    import time
    from ccsmeth_b200.bamio import BamReader
    ref = refimport.import_reference()
    import ccsmeth.extract_features as ref_ef
    import ccsmeth.call_modifications as rcm
    import ccsmeth._bam2modbam as rmb
    os.makedirs(os.path.join(OUT, "demo"), exist_ok=True)
    dst = os.path.join(OUT, "demo", "hg002.chr20_demo.hifi.bam")
    if not os.path.exists(dst):
        shutil.copyfile(refimport.DEMO_BAM, dst)
        os.chmod(dst, 0o644)
    m = refimport.load_ref_att2s()
    args = demo_args()
    reads = list(BamReader(dst))
    t0 = time.time()
    torch.manual_seed(1234)
    out = {"names": [], "n_sites_per_read": [], "locs": [], "prob1": [], "mm": [], "ml": [], "site_counts_per_batch": []}
    feat0 = None
    t_extract = t_model = 0.0
    for b0 in range(0, len(reads), args.holes_batch):
        hb = reads[b0:b0 + args.holes_batch]
        ta = time.time()
        feature_list, holeidx = [], []
        for i, r in enumerate(hb):
            f = ref_ef.extract_features_from_double_strand_read(r, ["CG"], None, None, None, args)
            feature_list += f
            holeidx += [i] * len(f)
        fb = rcm._batch_feature_list2s(feature_list)
        tb = time.time()
        pred, nb = rcm._call_mods2s(fb, m, 512, 0)
        tc = time.time()
        t_extract += tb - ta
        t_model += tc - tb
        out["site_counts_per_batch"].append(len(pred))
        if feat0 is None:  # feature-level golden for the first 3 reads
            k = sum(1 for h in holeidx if h < 3)
            feat0 = {"fkmer": np.array(fb[1][:k]), "fpass": np.array(fb[2][:k]), "fipd": np.array(fb[3][:k]),
                     "fpw": np.array(fb[5][:k]), "rkmer": np.array(fb[9][:k]), "rpass": np.array(fb[10][:k]),
                     "ripd": np.array(fb[11][:k]), "rpw": np.array(fb[13][:k]),
                     "locs": np.array([f[4] for f in feature_list[:k]]), "holeidx": np.array(holeidx[:k])}
        # per read MM/ML with the reference's converters
        for i, r in enumerate(hb):
            lp = sorted([(p[1], p[2]) for p, h in zip(pred, holeidx) if h == i], key=lambda x: x[0])
            out["names"].append(r.query_name)
            out["n_sites_per_read"].append(len(lp))
            if lp:
                locs, probs = zip(*lp)
                mm = rmb._convert_locs_to_mmtag(locs, r.get_forward_sequence())
                ml = rmb._convert_probs_to_mltag(probs)
                out["locs"] += list(locs)
                out["prob1"] += list(probs)
                out["mm"] += mm
                out["ml"] += ml
    dt = time.time() - t0
    n = len(out["prob1"])
    print("demo: %d reads, %d sites, batches %s; %.2f s total (extract+batch %.2f s, model %.2f s) -> %.0f sites/s"
          % (len(reads), n, out["site_counts_per_batch"], dt, t_extract, t_model, n / dt))
    np.savez_compressed(os.path.join(OUT, "demo_callmods.npz"), names=np.array(out["names"]),
                        n_sites_per_read=np.array(out["n_sites_per_read"]), locs=np.array(out["locs"]),
                        prob1=np.array(out["prob1"], dtype=np.float32), mm=np.array(out["mm"]),
                        ml=np.array(out["ml"], dtype=np.uint8),
                        site_counts_per_batch=np.array(out["site_counts_per_batch"]), tseed=1234,
                        ref_cpu_seconds=dt, ref_cpu_threads=torch.get_num_threads(),
                        **{"feat0." + k: v for k, v in feat0.items()})


def gen_pileup():
    """Config 5 / section 8f-3 golden: the reference's own region caller ``_call_modfreq_of_one_region``
    (call_mods_freq_bam.py:423-442 -> :308-420 in aggregate mode) on a synthetic region pileup, in aggregate and in
    count mode.  The model is built inside the reference function; its ``init_hidden`` draws are recorded so that the
    device path can be given the same h0 (groups in the order all reads / haplotype 1 / haplotype 2, one draw per
    1024 sites).  Also pins the two 256-entry tables (_cal_mod_prob, np.histogram bin)."""
    import argparse
    ref = refimport.import_reference()
    import ccsmeth.call_mods_freq_bam as rfb
    import ccsmeth.models as rmodels
    rng = np.random.default_rng(20261017)
    n = 2600
    pos = np.cumsum(rng.integers(2, 201, size=n)).astype(np.int64)
    cov = rng.integers(1, 41, size=n)
    cov[:5] = [1, 2, 3, 4, 5]
    ptr = np.concatenate(([0], np.cumsum(cov))).astype(np.int64)
    ml = np.floor(256 * rng.beta(0.3, 0.3, size=int(ptr[-1]))).clip(0, 255).astype(np.uint8)
    ml[:3] = [0, 128, 255]
    hap = rng.choice([0, 1, 2], size=int(ptr[-1]), p=[0.4, 0.3, 0.3]).astype(np.uint8)
    info = {int(pos[i]): [(rfb._cal_mod_prob(int(ml[k])), int(hap[k])) for k in range(ptr[i], ptr[i + 1])] for i in range(n)}

    def ns(**kw):
        a = argparse.Namespace(call_mode="aggregate", cov_cf=4, bin_size=20, prob_cf=0.0, no_amb_cov=False, no_hap=False,
                               seq_len=11, layer_rnn=1, class_num=1, hid_rnn=32, model_type="attbigru",
                               aggre_model=refimport.AGGR_CKPT, only_close=False, discrete=False, tseed=1234)
        for k, v in kw.items():
            setattr(a, k, v)
        return a

    def pack(res):
        out = np.full((3, n, 3), np.nan)
        assert [r[0] for r in res] == [int(p) for p in pos]
        for i, r in enumerate(res):
            for g in range(3):
                if r[1 + g] is not None:
                    out[g, i] = [float(x) for x in r[1 + g]]
        return out

    drawn = []
    orig = rmodels.AggrAttRNN.init_hidden

    def recording(self, *a, **k):
        h = orig(self, *a, **k)
        drawn.append(h.detach().clone())
        return h

    save = {"pos": pos, "ptr": ptr, "ml": ml, "hap": hap,
            "lut_prob": np.array([rfb._cal_mod_prob(v) for v in range(256)], dtype=np.float64),
            "lut_bin": np.array([int(np.argmax(np.histogram([rfb._cal_mod_prob(v)], bins=20, range=[0, 1])[0]))
                                 for v in range(256)], dtype=np.int32)}
    rmodels.AggrAttRNN.init_hidden = recording
    try:
        for tag, kw in (("aggr", {}), ("aggr_cf3", {"prob_cf": 0.3}), ("aggr_nohap", {"no_hap": True})):
            drawn.clear()
            res = rfb._call_modfreq_of_one_region(info, ns(**kw))
            save[tag] = pack(res)
            save[tag + "_h0"] = torch.cat(drawn, dim=1).numpy()   # (2, sum of high-coverage sites over groups, 32)
            save[tag + "_h0_sizes"] = np.array([h.shape[1] for h in drawn], dtype=np.int64)
    finally:
        rmodels.AggrAttRNN.init_hidden = orig
    for tag, kw in (("count", {}), ("count_cf3", {"prob_cf": 0.3}), ("count_cf3_noamb", {"prob_cf": 0.3, "no_amb_cov": True})):
        save[tag] = pack(rfb._call_modfreq_of_one_region(info, ns(call_mode="count", **kw)))
    np.savez_compressed(os.path.join(OUT, "pileup_region.npz"), **save)
    print("pileup_region: %d sites, %d calls, aggregate mean freq %.4f" % (n, len(ml), np.nanmean(save["aggr"][0, :, 2])))


def gen_aggr_variants():
    """Section 8f-3 beyond the shipped checkpoint: the reference's region caller with ``--model_type attbilstm`` and
    with another GRU shape (``--hid_rnn 48 --layer_rnn 2``).  No such checkpoints ship, so seeded random
    initialisations of the reference's own AggrAttRNN are saved to a temporary checkpoint and loaded by the reference
    function; every ``init_hidden`` draw (h0, for the LSTM also c0) is recorded."""
    import argparse
    import tempfile
    ref = refimport.import_reference()
    import ccsmeth.call_mods_freq_bam as rfb
    import ccsmeth.models as rmodels
    rng = np.random.default_rng(20261018)
    n = 1250
    pos = np.cumsum(rng.integers(2, 120, size=n)).astype(np.int64)
    cov = rng.integers(1, 31, size=n)
    ptr = np.concatenate(([0], np.cumsum(cov))).astype(np.int64)
    ml = np.floor(256 * rng.beta(0.3, 0.3, size=int(ptr[-1]))).clip(0, 255).astype(np.uint8)
    hap = rng.choice([0, 1, 2], size=int(ptr[-1]), p=[0.4, 0.3, 0.3]).astype(np.uint8)
    info = {int(pos[i]): [(rfb._cal_mod_prob(int(ml[k])), int(hap[k])) for k in range(ptr[i], ptr[i + 1])] for i in range(n)}
    save = {"pos": pos, "ptr": ptr, "ml": ml, "hap": hap}
    drawn = []
    orig = rmodels.AggrAttRNN.init_hidden

    def recording(self, *a, **k):
        h = orig(self, *a, **k)
        drawn.append(tuple(t.detach().clone() for t in h) if isinstance(h, tuple) else (h.detach().clone(),))
        return h

    rmodels.AggrAttRNN.init_hidden = recording
    try:
        for tag, mt, hid, layers, extra in (("lstm", "attbilstm", 32, 1, {}), ("gru48x2", "attbigru", 48, 2, {}),
                                            ("lstm_close", "attbilstm", 32, 1, {"only_close": True})):
            torch.manual_seed(97 + hid + layers)
            m = ref.models.AggrAttRNN(11, layers, 1, 0, hid, binsize=20, model_type=mt, device="cpu")
            with tempfile.NamedTemporaryFile(suffix=".ckpt") as f:
                torch.save(m.state_dict(), f.name)
                a = argparse.Namespace(call_mode="aggregate", cov_cf=4, bin_size=20, prob_cf=0.0, no_amb_cov=False,
                                       no_hap=False, seq_len=11, layer_rnn=layers, class_num=1, hid_rnn=hid, model_type=mt,
                                       aggre_model=f.name, only_close=False, discrete=False, tseed=1234)
                for k, v in extra.items():
                    setattr(a, k, v)
                drawn.clear()
                res = rfb._call_modfreq_of_one_region(info, a)
            out = np.full((3, n, 3), np.nan)
            for i, r in enumerate(res):
                for g in range(3):
                    if r[1 + g] is not None:
                        out[g, i] = [float(x) for x in r[1 + g]]
            save[tag] = out
            if extra:  # same seed, same constructor, same site counts: the draws repeat those of the plain case
                assert np.array_equal(torch.cat([d[0] for d in drawn], dim=1).numpy(), save["lstm_h0"])
                continue
            save[tag + "_h0"] = torch.cat([d[0] for d in drawn], dim=1).numpy()
            if mt == "attbilstm":
                save[tag + "_c0"] = torch.cat([d[1] for d in drawn], dim=1).numpy()
            save[tag + "_sizes"] = np.array([d[0].shape[1] for d in drawn], dtype=np.int64)
            if not extra:
                for k, v in m.state_dict().items():
                    save[tag + ".sd." + k] = v.detach().numpy()
            print("aggr_variants %s: mean freq %.4f" % (tag, np.nanmean(out[0, :, 2])))
    finally:
        rmodels.AggrAttRNN.init_hidden = orig
    np.savez_compressed(os.path.join(OUT, "aggr_variants.npz"), **save)


def gen_cli():
    """The reference's own argparse definitions of `call_mods` and `call_freqb` (ccsmeth/ccsmeth.py:195-330, 560-650):
    every flag with its aliases, default, type, choices and action, captured from the parser object the reference's
    main() builds.  tests/test_cli_cpu.py checks that the ccsmeth_b200 CLIs accept the same command lines."""
    import argparse
    import json
    refimport.import_reference()
    import ccsmeth.ccsmeth as rmain

    class Captured(Exception):
        pass

    orig = argparse.ArgumentParser.parse_args

    def capture(self, *a, **k):
        e = Captured()
        e.parser = self
        raise e

    argparse.ArgumentParser.parse_args = capture
    try:
        rmain.main()
    except Captured as e:
        parser = e.parser
    finally:
        argparse.ArgumentParser.parse_args = orig
    subs = next(a for a in parser._actions if isinstance(a, argparse._SubParsersAction)).choices
    out = {}
    for name in ("call_mods", "call_freqb"):
        flags = []
        for a in subs[name]._actions:
            if isinstance(a, argparse._HelpAction):
                continue
            flags.append({"flags": list(a.option_strings), "dest": a.dest, "default": a.default,
                          "type": getattr(a.type, "__name__", None), "choices": list(a.choices) if a.choices else None,
                          "action": type(a).__name__, "required": bool(a.required)})
        out[name] = flags
    with open(os.path.join(OUT, "cli_flags.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print("cli_flags: %s" % {k: len(v) for k, v in out.items()})


def gen_lstm():
    """Section 8f-4: the reference's ModelAttRNN(model_type="attbilstm2s") -- no checkpoint ships, so a seeded random
    initialisation of a small configuration (hidden 64, 2 layers) is the fixture -- with explicit (h0, c0)."""
    ref = refimport.import_reference()
    torch.manual_seed(4321)
    m = ref.models.ModelAttRNN(21, 2, 2, 0, 64, is_npass=True, model_type="attbilstm2s", device="cpu")
    m.eval()
    n = 128
    b = synth.make_batch(n, seed=synth.SEED + 5, with_h0=False)
    g = torch.Generator().manual_seed(99)
    hc = [(torch.randn(4, n, 64, generator=g), torch.randn(4, n, 64, generator=g)) for _ in range(2)]
    z = torch.zeros(n)
    args = (b["kmer"], b["kpass"], b["ipd"], z, b["pw"], z, z, z, b["kmer2"], b["kpass2"], b["ipd2"], z, b["pw2"], z, z, z)
    import ccsmeth.models as rmodels
    old = rmodels.use_cuda
    rmodels.use_cuda = False
    try:
        with refimport.fixed_h0(m, [(h.clone(), c.clone()) for h, c in hc]):
            logits, probs = m(*args)
        torch.manual_seed(777)  # the reference's own draw order: h0, c0 of strand 1, then of strand 2
        _, probs_seeded = m(*args)
    finally:
        rmodels.use_cuda = old
    save = {"sd." + k: v.detach().numpy() for k, v in m.state_dict().items()}
    save.update({k: b[k].numpy() for k in ("kmer", "kpass", "ipd", "pw", "kmer2", "kpass2", "ipd2", "pw2")})
    save.update({"h0_f": hc[0][0].numpy(), "c0_f": hc[0][1].numpy(), "h0_r": hc[1][0].numpy(), "c0_r": hc[1][1].numpy(),
                 "logits": logits.detach().numpy(), "probs": probs.detach().numpy(),
                 "probs_seeded": probs_seeded.detach().numpy(), "seed": 777})
    np.savez_compressed(os.path.join(OUT, "att2s_lstm.npz"), **save)
    print("att2s_lstm: mean p1 %.4f" % probs[:, 1].mean().item())


def gen_2s2():
    """Section 8f-4: the reference's ModelAttRNN2 ("attbigru2s2", "attbilstm2s2") on seeded random weights (hidden 64,
    2 layers; no checkpoint ships), raw integer kinetics as input (the model embeds them), explicit initial states."""
    ref = refimport.import_reference()
    import ccsmeth.models as rmodels
    rng = np.random.default_rng(12)
    n, L = 96, 21
    feats = {}
    for sfx in ("", "2"):
        feats["kmer" + sfx] = torch.from_numpy(rng.integers(0, 5, (n, L)).astype(np.float32))
        feats["ipd" + sfx] = torch.from_numpy(rng.integers(0, 953, (n, L)).astype(np.float32))
        feats["pw" + sfx] = torch.from_numpy(rng.integers(0, 953, (n, L)).astype(np.float32))
        feats["kpass" + sfx] = torch.from_numpy(np.repeat(rng.integers(0, 45, (n, 1)), L, axis=1).astype(np.float32))
    z = torch.zeros(n)
    args = (feats["kmer"], feats["kpass"], feats["ipd"], z, feats["pw"], z, z, z,
            feats["kmer2"], feats["kpass2"], feats["ipd2"], z, feats["pw2"], z, z, z)
    save = {k: v.numpy() for k, v in feats.items()}
    old = rmodels.use_cuda
    rmodels.use_cuda = False
    try:
        for tag, mt in (("gru", "attbigru2s2"), ("lstm", "attbilstm2s2")):
            torch.manual_seed(2468)
            m = rmodels.ModelAttRNN2(21, 2, 2, 0, 32, is_npass=True, model_type=mt, device="cpu")
            m.eval()
            g = torch.Generator().manual_seed(7)
            hs = [torch.randn(4, n, 32, generator=g) for _ in range(4)]
            fixed = [hs[0], hs[1]] if tag == "gru" else [(hs[0], hs[1]), (hs[2], hs[3])]
            with refimport.fixed_h0(m, [tuple(t.clone() for t in f) if isinstance(f, tuple) else f.clone() for f in fixed]):
                logits, probs = m(*args)
            save.update({"%s.sd.%s" % (tag, k): v.detach().numpy() for k, v in m.state_dict().items()})
            save.update({"%s.h%d" % (tag, i): h.numpy() for i, h in enumerate(hs[:2 if tag == "gru" else 4])})
            save["%s.logits" % tag], save["%s.probs" % tag] = logits.detach().numpy(), probs.detach().numpy()
            print("2s2", tag, "mean p1 %.4f" % probs[:, 1].mean().item())
    finally:
        rmodels.use_cuda = old
    np.savez_compressed(os.path.join(OUT, "att2s2.npz"), **save)


def gen_transenc():
    """Section 8f-4: the reference's ModelTransEnc ("transencoder2s") on seeded random weights (d_model 64, 4 heads,
    dim_ff 128, 2 layers; BatchNorm running statistics randomised so that the eval-mode affine is exercised)."""
    ref = refimport.import_reference()
    import ccsmeth.models as rmodels
    rng = np.random.default_rng(21)
    n, L = 64, 21
    feats = {}
    for sfx in ("", "2"):
        feats["kmer" + sfx] = torch.from_numpy(rng.integers(0, 5, (n, L)).astype(np.float32))
        feats["ipd" + sfx] = torch.from_numpy(rng.integers(0, 953, (n, L)).astype(np.float32))
        feats["pw" + sfx] = torch.from_numpy(rng.integers(0, 953, (n, L)).astype(np.float32))
        feats["kpass" + sfx] = torch.from_numpy(np.repeat(rng.integers(0, 45, (n, 1)), L, axis=1).astype(np.float32))
    z = torch.zeros(n)
    args = (feats["kmer"], feats["kpass"], feats["ipd"], z, feats["pw"], z, z, z,
            feats["kmer2"], feats["kpass2"], feats["ipd2"], z, feats["pw2"], z, z, z)
    torch.manual_seed(1357)
    m = rmodels.ModelTransEnc(21, 2, 2, 0, 64, 4, 128, is_npass=True, device="cpu")
    with torch.no_grad():
        for k, v in m.state_dict().items():
            if k.endswith("running_mean"):
                v.copy_(torch.randn(v.shape) * 0.05)
            elif k.endswith("running_var"):
                v.copy_(torch.rand(v.shape) * 0.5 + 0.05)
            elif ".conv_embed." in k and k.endswith(("1.weight", "5.weight", "1.bias", "5.bias")) and v.dim() == 1:
                v.copy_(torch.rand(v.shape) + 0.5 if k.endswith("weight") else torch.randn(v.shape) * 0.1)
    m.eval()
    with torch.no_grad():
        logits, probs = m(*args)
    save = {k: v.numpy() for k, v in feats.items()}
    save.update({"sd." + k: v.detach().numpy() for k, v in m.state_dict().items() if "num_batches_tracked" not in k})
    save["logits"], save["probs"] = logits.numpy(), probs.numpy()
    np.savez_compressed(os.path.join(OUT, "transenc.npz"), **save)
    print("transenc: mean p1 %.4f, logits spread %.4f" % (probs[:, 1].mean().item(), logits.std().item()))


class DuckBam:
    """What the reference's region worker needs from pysam.AlignmentFile: fetch(contig, start, stop) over records that
    overlap the interval, in file order (records are ccsmeth_b200.bamio.BamRecord)."""

    def __init__(self, path):
        from ccsmeth_b200.bamio import BamReader
        self.recs = list(BamReader(path))

    def fetch(self, contig=None, start=None, stop=None):
        for r in self.recs:
            if r.is_unmapped or r.reference_name != contig:
                continue
            if r.pos < stop and r.reference_end > start:
                yield r


def freqb_args(**kw):
    import argparse
    a = argparse.Namespace(call_mode="count", cov_cf=4, bin_size=20, prob_cf=0.0, no_amb_cov=False, no_hap=False, seq_len=11,
                           layer_rnn=1, class_num=1, hid_rnn=32, model_type="attbigru", aggre_model=refimport.AGGR_CKPT,
                           only_close=False, discrete=False, tseed=1234, modtype="5mC", mod_loc=0, motifs="CG",
                           no_comb=False, refsites_only=False, refsites_all=False, no_supplementary=False, mapq=1,
                           identity=0.0, hap_tag="HP", base_clip=0, contigs=None, chunk_len=10000, bed=False)
    for k, v in kw.items():
        setattr(a, k, v)
    return a


FREQB_CASES = (("count", {}), ("count_cf3", {"prob_cf": 0.3}), ("count_cf3_noamb", {"prob_cf": 0.3, "no_amb_cov": True}),
               ("count_nocomb", {"no_comb": True}), ("count_refsites", {"refsites_only": True}),
               ("count_clip_nosupp_ident", {"base_clip": 15, "no_supplementary": True, "identity": 0.995, "mapq": 20}),
               ("aggregate", {"call_mode": "aggregate"}), ("aggregate_nohap", {"call_mode": "aggregate", "no_hap": True}),
               ("aggregate_discrete", {"call_mode": "aggregate", "discrete": True, "no_hap": True}),
               ("aggregate_onlyclose", {"call_mode": "aggregate", "only_close": True, "no_hap": True}),
               ("count_refsites_all", {"refsites_all": True}),
               ("count_refsites_all_clip_nocomb", {"refsites_all": True, "base_clip": 40, "no_comb": True}),
               ("aggregate_refsites_all", {"call_mode": "aggregate", "refsites_all": True, "no_hap": True}))


def gen_freqb():
    """Section 8f-3 host half: the reference's region worker `_readmods_to_bed_of_one_region` + `_write_one_line`
    (call_mods_freq_bam.py:457-594, 626-634) over a synthetic aligned modbam (tests/bamsynth.make_aligned_modbam),
    one text file per flag combination and output group, in both output formats."""
    sys.path.insert(0, ROOT)
    from tests.bamsynth import make_aligned_modbam
    ref = refimport.import_reference()
    import ccsmeth.call_mods_freq_bam as rfb
    from ccsmeth.utils.ref_reader import DNAReference
    from ccsmeth.utils.process_utils import get_motif_seqs
    d = os.path.join(OUT, "freqb")
    os.makedirs(d, exist_ok=True)
    bam, fa = os.path.join(d, "synth.aligned.modbam.bam"), os.path.join(d, "synth.fa")
    make_aligned_modbam(bam, fa)
    dnacontigs = DNAReference(fa).getcontigs()
    reader = DuckBam(bam)
    import io
    texts = {}
    for tag, kw in FREQB_CASES:
        args = freqb_args(**kw)
        motifs = get_motif_seqs(args.motifs)
        mf = motifs if (args.refsites_only or args.refsites_all) else None
        chunks = rfb._get_reference_chunks(dnacontigs, args.contigs, args.chunk_len, args.motifs)
        outs = {(g, fmt): io.StringIO() for g in ("all", "hp1", "hp2") for fmt in ("bed", "freq.txt")}
        for region in chunks:
            beds = rfb._readmods_to_bed_of_one_region(reader, region, dnacontigs, mf, args)
            if len(beds[0]) == 0:
                continue
            for g, items in zip(("all", "hp1", "hp2"), beds):
                for it in items:
                    rfb._write_one_line(it, outs[(g, "bed")], True)
                    rfb._write_one_line(it, outs[(g, "freq.txt")], False)
        for (g, fmt), f in outs.items():
            if fmt == "bed" and not (g == "all" and tag in ("count", "aggregate")):
                continue  # the bed format of the same tuples: two samples are enough
            texts["%s.%s.%s" % (tag, g, fmt)] = np.frombuffer(f.getvalue().encode("ascii"), dtype=np.uint8)
        print("freqb", tag, {g: outs[(g, "bed")].getvalue().count("\n") for g in ("all", "hp1", "hp2")})
    texts["chunks"] = np.frombuffer("".join("%s\t%d\t%d\n" % c for c in
                                            rfb._get_reference_chunks(dnacontigs, None, 10000, "CG")).encode(), dtype=np.uint8)
    np.savez_compressed(os.path.join(d, "reference_outputs.npz"), **texts)


def gen_norm_mad():
    """`--norm mad` through the reference's own _normalize_signals (extract_features.py:181-199).  statsmodels is not in this
    image: `robust.scale.mad` is provided by oracle.extract_numpy.statsmodels_mad, the restated formula of statsmodels 0.14.0
    (the reference's environment.yml:13) -- so this fixture pins the reference's code AROUND the call (np.median as the
    shift, float(), the zero-scale branch, np.around(6)), not statsmodels itself."""
    import types
    refimport.import_reference()
    import ccsmeth.extract_features as ref_ef
    from oracle.extract_numpy import statsmodels_mad
    scale = types.ModuleType("statsmodels.robust.scale")
    scale.mad = statsmodels_mad
    ref_ef.robust.scale = scale
    rng = np.random.default_rng(41)
    out = {}
    cases = [rng.integers(0, 953, 301), rng.integers(0, 953, 300), rng.integers(0, 60, 4096), np.full(50, 17),
             np.array([3, 3, 3, 9, 9, 9]), np.array([5]), np.array([1, 2]), rng.integers(0, 4, 1001),
             np.concatenate([np.full(400, 20), rng.integers(0, 953, 399)])]
    for i, sig in enumerate(cases):
        sig = np.asarray(sig, dtype=np.int64)
        out["sig%d" % i] = sig
        out["out%d" % i] = np.asarray(ref_ef._normalize_signals(sig, "mad"), dtype=np.float64)
    np.savez_compressed(os.path.join(OUT, "norm_mad.npz"), **out)
    print("norm_mad", {k: v.shape for k, v in out.items() if k.startswith("out")})


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    if len(sys.argv) > 1 and sys.argv[1] == "norm_mad":
        gen_norm_mad()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "att2s_16k":
        gen_att2s_16k()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "transenc":
        gen_transenc()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "2s2":
        gen_2s2()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "lstm":
        gen_lstm()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "freqb":
        gen_freqb()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "demo":
        gen_demo()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "pileup":
        gen_pileup()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "cli":
        gen_cli()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "aggr_variants":
        gen_aggr_variants()
        sys.exit(0)
    save_ckpts()
    gen_att2s()
    gen_aggr()
    gen_pileup()
    gen_aggr_variants()
    gen_cli()
    gen_lstm()
    gen_2s2()
    gen_transenc()
    gen_freqb()
    gen_demo()
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))
