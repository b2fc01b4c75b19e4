#!/usr/bin/env python
"""CPU baseline arm -- TEST / MEASUREMENT INFRASTRUCTURE, not product code.

Times the reference's own CPU implementation of the hot path on this machine's host cores and prints ONE JSON line.
It runs in its own process with CUDA hidden, so that the reference's ``use_cuda`` switch
(reference ccsmeth/utils/constants_torch.py:5) stays off even on the GPU box.

    python oracle/ref_cpu_bench.py forward [--batches B] [--warmup W]   ModelAttRNN.forward, batch 512 (config 1/2)
    python oracle/ref_cpu_bench.py demo                                 extract -> _batch_feature_list2s -> _call_mods2s
                                                                         over the demo BAM (config 1/3)
    python oracle/ref_cpu_bench.py aggr [--batches B]                   AggrAttRNN.forward, batch 1024 (config 5)

``kind`` = "reference" when the unmodified reference package is importable (/root/reference in the build container,
oracle/_ref staged by oracle/stage_ref.py on the GPU box), else "port" (oracle/torch_port.py: the same ATen calls).
Weights and the demo BAM are the committed fixtures under tests/golden/ (tensor-for-tensor / byte-for-byte the
reference's shipped files).
"""
import os
os.environ["CUDA_VISIBLE_DEVICES"] = ""   # before torch: the reference must take its CPU path

import argparse
import json
import sys
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
G = os.path.join(ROOT, "tests", "golden")
FEATS = ("kmer", "kpass", "ipd", "pw", "kmer2", "kpass2", "ipd2", "pw2")


def _ref_att2s():
    """(model, kind): the reference ModelAttRNN(attbigru2s) with the v3 weights (construction mirrors reference
    call_modifications.py:316-369), or the torch port."""
    from oracle import refimport, torch_port
    ck = dict(np.load(os.path.join(G, "ckpt_att2s_v3.npz")))
    if refimport.package_root() is not None:
        ref = refimport.import_reference()
        m = ref.models.ModelAttRNN(21, 3, 2, 0, 256, is_npass=True, is_sn=False, is_map=False, is_stds=False,
                                   model_type="attbigru2s", device=0)
        d = m.state_dict()
        d.update({k: torch.from_numpy(v.copy()) for k, v in ck.items()})
        m.load_state_dict(d)
        m.eval()
        return m, "reference"
    m = torch_port.load_numpy_state(torch_port.Att2sPort(), ck)
    return m, "port"


def forward(args):
    from ccsmeth_b200 import synth
    m, kind = _ref_att2s()
    b = synth.make_batch(args.batch_size, seed=synth.SEED, with_h0=False)
    a = synth.to_forward_args(b) if kind == "reference" else [b[k] for k in FEATS]
    for _ in range(args.warmup):
        m(*a)
    t0 = time.perf_counter()
    for _ in range(args.batches):
        m(*a)      # like the reference's batch loop: no torch.no_grad(), h0 drawn per call (models.py:77-87)
    dt = time.perf_counter() - t0
    n = args.batches * args.batch_size
    return {"mode": "forward", "kind": kind, "value": n / dt, "unit": "sites/s", "seconds": dt, "sites": n,
            "sample": "%d batches x %d sites, %s ModelAttRNN.forward on CPU (torch %s)" %
                      (args.batches, args.batch_size, "reference" if kind == "reference" else "torch port of the reference",
                       torch.__version__)}


def demo(args):
    """The reference chain over the demo BAM in one process (the reference's CPU topology is 2 such worker processes,
    process_utils.py:77; this times one with every torch thread): extract_features_from_double_strand_read ->
    _batch_feature_list2s -> _call_mods2s(batch 512) -> MM/ML conversion.  Reads are parsed by ccsmeth_b200.bamio
    (pysam is absent); BAM writing is not included."""
    from oracle import refimport
    from ccsmeth_b200.bamio import BamReader
    if refimport.package_root() is None:
        return {"mode": "demo", "kind": "unavailable", "value": None}
    refimport.import_reference()
    import ccsmeth.extract_features as ref_ef
    import ccsmeth.call_modifications as rcm
    import ccsmeth._bam2modbam as rmb
    m, kind = _ref_att2s()
    a = argparse.Namespace(mode="denovo", seq_len=21, motifs="CG", mod_loc=0, methy_label=1, norm="zscore",
                           no_decode=False, is_sn="no", is_map="no", is_stds="no", is_npass="yes", mapq=1,
                           identity=0.0, no_supplementary=False, skip_unmapped="yes", holes_batch=50,
                           batch_size=512, keep_pulse=False)
    bam = os.path.join(G, "demo", "hg002.chr20_demo.hifi.bam")
    best, sites = None, 0
    for _ in range(args.repeats):
        t0 = time.perf_counter()
        reads = list(BamReader(bam))
        torch.manual_seed(1234)
        sites = 0
        for b0 in range(0, len(reads), a.holes_batch):
            hb = reads[b0:b0 + a.holes_batch]
            feats, holeidx = [], []
            for i, r in enumerate(hb):
                f = ref_ef.extract_features_from_double_strand_read(r, ["CG"], None, None, None, a)
                feats += f
                holeidx += [i] * len(f)
            pred, _nb = rcm._call_mods2s(rcm._batch_feature_list2s(feats), m, a.batch_size, 0)
            sites += len(pred)
            for i, r in enumerate(hb):
                lp = sorted([(p[1], p[2]) for p, h in zip(pred, holeidx) if h == i], key=lambda x: x[0])
                if lp:
                    locs, probs = zip(*lp)
                    rmb._convert_locs_to_mmtag(locs, r.get_forward_sequence())
                    rmb._convert_probs_to_mltag(probs)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return {"mode": "demo", "kind": kind, "value": sites / best, "unit": "sites/s", "seconds": best, "sites": sites,
            "sample": "demo BAM (116 reads, %d sites): reference extract -> batch -> _call_mods2s -> MM/ML in one "
                      "process, best of %d (torch %s)" % (sites, args.repeats, torch.__version__)}


def aggr(args):
    from oracle import refimport, torch_port
    from collections import OrderedDict
    ck = dict(np.load(os.path.join(G, "ckpt_aggr_v2p.npz")))
    ck = OrderedDict((k[7:] if k.startswith("module.") else k, torch.from_numpy(v.copy())) for k, v in ck.items())
    if refimport.package_root() is not None:
        ref = refimport.import_reference()
        m = ref.models.AggrAttRNN(11, 1, 1, 0, 32, binsize=20, model_type="attbigru", device="cpu")
        m.load_state_dict(ck)
        m.eval()
        kind = "reference"
    else:
        m = torch_port.AggrPort()
        m.load_state_dict(ck)
        kind = "port"
    g = torch.Generator().manual_seed(20261017)
    n = 1024  # the reference's aggregate batch (call_mods_freq_bam.py:295)
    histos = torch.rand((n, 11, 20), generator=g)
    histos = torch.round(histos / histos.norm(dim=2, keepdim=True) * 1e6) / 1e6
    offsets = torch.randint(0, 1200, (n, 11), generator=g).float()
    for _ in range(args.warmup):
        m(offsets, histos)
    t0 = time.perf_counter()
    for _ in range(args.batches):
        m(offsets, histos)
    dt = time.perf_counter() - t0
    return {"mode": "aggr", "kind": kind, "value": args.batches * n / dt, "unit": "sites/s", "seconds": dt,
            "sites": args.batches * n,
            "sample": "%d batches x %d sites, %s AggrAttRNN.forward on CPU" % (args.batches, n, kind)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("mode", choices=["forward", "demo", "aggr"])
    ap.add_argument("--batches", type=int, default=32)
    ap.add_argument("--batch-size", type=int, default=512)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--repeats", type=int, default=2)
    ap.add_argument("--threads", type=int, default=0)
    args = ap.parse_args()
    threads = args.threads or os.cpu_count()
    torch.set_num_threads(threads)
    out = {"forward": forward, "demo": demo, "aggr": aggr}[args.mode](args)
    out["cores"] = threads
    print(json.dumps(out))


if __name__ == "__main__":
    main()
