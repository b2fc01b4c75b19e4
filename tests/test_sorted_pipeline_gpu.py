"""call_mods on an aligned (unsorted) HiFi BAM -> coordinate-sorted, indexed modbam -> call_freqb, without samtools
(reference call_modifications.py:592-607 sorts and indexes through pysam so that call_freqb can fetch regions)."""
import os

import numpy as np
import pytest

from ccsmeth_b200 import call_freqb as cf, call_mods as cm
from ccsmeth_b200.bamio import BamReader, BamWriter
from tests.bamsynth import random_read

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ckpt_file(tmp_path_factory, ckpt_att2s):
    import torch
    from collections import OrderedDict
    p = str(tmp_path_factory.mktemp("ckpt") / "model_v3.ckpt")
    torch.save(OrderedDict((k, torch.from_numpy(v)) for k, v in ckpt_att2s.items()), p)
    return p


def _reads(rng, n_reads, ref_len):
    recs = []
    for k in range(n_reads):
        n = int(rng.integers(400, 1500))
        pos = int(rng.integers(0, ref_len - n))
        rev = bool(rng.random() < 0.5)
        rec, _ = random_read(rng, "m0/%d/ccs" % k, n, reverse=rev, flag=16 if rev else 0, cigar=((0, n),), mapq=60,
                             ref_id=int(rng.random() < 0.3), pos=pos)
        recs.append(rec)
    return recs


def test_unsorted_aligned_input_comes_out_sorted_indexed_and_feeds_call_freqb(tmp_path, ckpt_file):
    rng = np.random.default_rng(17)
    refs = [("ctgA", 30000), ("ctgB", 12000)]
    fa = str(tmp_path / "ref.fa")
    with open(fa, "w") as f:
        for name, ln in refs:
            f.write(">%s\n%s\n" % (name, "".join(np.array(list("ACGT"))[rng.integers(0, 4, ln)])))
    recs = _reads(rng, 60, 12000)
    hdr = "@HD\tVN:1.6\tSO:unsorted\n" + "".join("@SQ\tSN:%s\tLN:%d\n" % r for r in refs)
    order = sorted(range(len(recs)), key=lambda i: (recs[i].ref_id, recs[i].pos, recs[i].is_reverse))
    paths = {}
    for tag, idx in (("sorted", order), ("shuffled", list(rng.permutation(len(recs))))):
        paths[tag] = str(tmp_path / (tag + ".bam"))
        w = BamWriter(paths[tag], hdr, refs)
        for i in idx:
            w.write_raw(recs[i].raw)
        w.close()
    base = ["-m", ckpt_file, "--mode", "align", "--h0", "zeros", "--holes_batch", "7", "--threads", "2"]
    a = cm.build_parser().parse_args(["-i", paths["sorted"], "-o", str(tmp_path / "a"), "--no_sort"] + base)
    ca, pa = cm.call_mods(a)
    b = cm.build_parser().parse_args(["-i", paths["shuffled"], "-o", str(tmp_path / "b")] + base)
    cb, pb = cm.call_mods(b)
    assert ca == cb and ca["sites"] > 0
    assert not os.path.exists(pa + ".bai") and os.path.exists(pb + ".bai")
    ra, rb = list(BamReader(pa)), BamReader(pb)
    got = list(rb)
    assert "SO:coordinate" in rb.header_text.split("\n")[0]
    keys = [(r.ref_id, r.pos, r.is_reverse) for r in got]
    assert keys == sorted(keys)
    assert sorted(r.raw for r in got) == sorted(r.raw for r in ra)       # same records, tags included
    # call_freqb (count mode) streams the sorted output; the unsorted-input run gives the same frequencies as the sorted one
    outs = {}
    for tag, p in (("a", pa), ("b", pb)):
        o = str(tmp_path / ("freq_" + tag))
        cf.main(["--input_bam", p, "--ref", fa, "-o", o, "--threads", "2", "--chunk_len", "5000"])
        outs[tag] = open(o + ".count.all.freq.txt").read()
    assert outs["a"] == outs["b"] and len(outs["a"].splitlines()) > 10
    # and a sorted input is only indexed, not rewritten
    c = cm.build_parser().parse_args(["-i", paths["sorted"], "-o", str(tmp_path / "c")] + base)
    _, pc = cm.call_mods(c)
    assert os.path.exists(pc + ".bai") and [r.raw for r in BamReader(pc)] == [r.raw for r in ra]
