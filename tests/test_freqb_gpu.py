"""call_freqb end to end on the GPU (sorted modbam + FASTA in, frequency files out) against the files the reference's
own region worker + writer produce for the same input (tests/golden/freqb/, scripts/gen_golden.py gen_freqb).
Count mode: byte-identical files.  Aggregate mode: same sites, same coverage, low-coverage lines identical, model
frequencies within 2e-6 (float32 output rounded to 6 decimals) -- with the reference's own per-region seeded h0."""
import os
from collections import OrderedDict

import numpy as np
import pytest
import torch

from ccsmeth_b200 import call_freqb as cf
from tests.conftest import GOLDEN

pytestmark = pytest.mark.gpu
D = os.path.join(GOLDEN, "freqb")
BAM, FA = os.path.join(D, "synth.aligned.modbam.bam"), os.path.join(D, "synth.fa")


@pytest.fixture(scope="module")
def ref_out():
    with np.load(os.path.join(D, "reference_outputs.npz")) as z:
        return {k: z[k].tobytes().decode("ascii") for k in z.files}


@pytest.fixture(scope="module")
def aggr_ckpt(tmp_path_factory, ckpt_aggr):
    p = str(tmp_path_factory.mktemp("ckpt") / "aggr.ckpt")
    torch.save(OrderedDict((k, torch.from_numpy(v)) for k, v in ckpt_aggr.items()), p)
    return p


def _run(tmp_path, tag, extra, aggr_ckpt):
    out = str(tmp_path / tag)
    argv = ["--input_bam", BAM, "--ref", FA, "-o", out, "--chunk_len", "10000", "-m", aggr_ckpt] + extra
    counts, paths = cf.call_freqb(cf.build_parser().parse_args(argv))
    return counts, paths


def _read(path):
    return open(path).read() if os.path.exists(path) else ""


@pytest.mark.parametrize("tag,extra", [
    ("count", []), ("count_cf3", ["--prob_cf", "0.3"]), ("count_cf3_noamb", ["--prob_cf", "0.3", "--no_amb_cov"]),
    ("count_nocomb", ["--no_comb"]), ("count_refsites", ["--refsites_only"]),
    ("count_clip_nosupp_ident", ["--base_clip", "15", "--no_supplementary", "--identity", "0.995", "--mapq", "20"]),
    ("count_refsites_all", ["--refsites_all"]),
    ("count_refsites_all_clip_nocomb", ["--refsites_all", "--base_clip", "40", "--no_comb"])])
def test_count_mode_files_are_identical_to_the_reference(tmp_path, ref_out, aggr_ckpt, tag, extra):
    counts, paths = _run(tmp_path, tag, extra, aggr_ckpt)
    for name, p in zip(("all", "hp1", "hp2"), paths):
        assert p.endswith(".count.%s.freq.txt" % name)
        assert _read(p) == ref_out["%s.%s.freq.txt" % (tag, name)], (tag, name)
    if tag == "count":
        _, bed_paths = _run(tmp_path, tag + "_bed", extra + ["--bed"], aggr_ckpt)
        assert _read(bed_paths[0]) == ref_out["count.all.bed"]


@pytest.mark.parametrize("tag,extra", [("aggregate", []), ("aggregate_nohap", ["--no_hap"]),
                                       ("aggregate_discrete", ["--no_hap", "--discrete"]),
                                       ("aggregate_onlyclose", ["--no_hap", "--only_close"]),
                                       ("aggregate_refsites_all", ["--no_hap", "--refsites_all"])])
def test_aggregate_mode_files_match_the_reference(tmp_path, ref_out, aggr_ckpt, tag, extra):
    counts, paths = _run(tmp_path, tag, ["--call_mode", "aggregate"] + extra, aggr_ckpt)
    for name, p in zip(("all", "hp1", "hp2"), paths):
        mine, ref = _read(p).splitlines(), ref_out["%s.%s.freq.txt" % (tag, name)].splitlines()
        assert len(mine) == len(ref), (tag, name)
        n_model = n_same = 0
        for a, b in zip(mine, ref):
            fa, fb = a.split("\t"), b.split("\t")
            assert fa[:6] == fb[:6] and fa[8] == fb[8]  # contig, position, strand, coverage
            if int(fb[8]) < 4:
                assert a == b  # count path
            else:
                n_model += 1
                if tag == "aggregate_discrete":
                    # discretize_score snaps to whole reads: a 1e-6 difference can move the count by one read at most
                    assert abs(float(fa[6]) - float(fb[6])) <= 1.0 and abs(float(fa[9]) - float(fb[9])) <= 1.0 / int(fb[8]) + 1e-4
                    n_same += a == b
                    continue
                assert abs(float(fa[9]) - float(fb[9])) <= 1.01e-4  # printed with 4 decimals
                assert abs(float(fa[6]) - float(fb[6])) <= 0.0101   # round(cov * freq, 2)
        assert n_model > 100 or not ref
        if tag == "aggregate_discrete" and ref:
            assert n_same >= 0.99 * n_model
    if tag == "aggregate":
        _, bed_paths = _run(tmp_path, tag + "_bed", ["--call_mode", "aggregate", "--bed"], aggr_ckpt)
        mine, ref = _read(bed_paths[0]).splitlines(), ref_out["aggregate.all.bed"].splitlines()
        assert len(mine) == len(ref)
        diff = [abs(int(a.split("\t")[10]) - int(b.split("\t")[10])) for a, b in zip(mine, ref)]
        assert max(diff) <= 1 and np.mean(np.array(diff) > 0) < 0.01  # percent methylation, integer


def test_aggregate_identical_lines_when_h0_stream_is_reproduced(tmp_path, ref_out, aggr_ckpt):
    """With --h0 reference the per-region h0 stream is the reference's (seed, model construction, one randn per 1024
    sites): almost every printed line is then identical text."""
    _, paths = _run(tmp_path, "agg_same", ["--call_mode", "aggregate"], aggr_ckpt)
    mine, ref = _read(paths[0]).splitlines(), ref_out["aggregate.all.freq.txt"].splitlines()
    same = sum(a == b for a, b in zip(mine, ref))
    assert same >= 0.98 * len(ref), (same, len(ref))


def test_region_shards_over_ranks_cover_the_single_rank_output(tmp_path, ref_out, aggr_ckpt, monkeypatch):
    """Reference chunks dealt round-robin to two "ranks": the union of the rank files equals the one-rank file."""
    from ccsmeth_b200 import parallel
    lines = []
    for rank in (0, 1):
        monkeypatch.setattr(parallel, "init_from_env", lambda r=rank: (r, 2, 0))
        monkeypatch.setattr(parallel, "allreduce_counts", lambda c: list(c))
        _, paths = _run(tmp_path, "shard", [], aggr_ckpt)
        assert ".rank%d.count.all." % rank in paths[0]
        lines += _read(paths[0]).splitlines()
    ref = ref_out["count.all.freq.txt"].splitlines()
    assert len(lines) == len(ref) and sorted(lines) == sorted(ref)


def test_sort_and_gzip_outputs(tmp_path, ref_out, aggr_ckpt):
    import gzip
    _, paths = _run(tmp_path, "sorted", ["--sort", "--contigs", "chrB,chrA"], aggr_ckpt)
    lines = _read(paths[0]).splitlines()
    keys = [(ln.split("\t")[0], int(ln.split("\t")[1])) for ln in lines]
    assert keys == sorted(keys) and sorted(lines) == sorted(ref_out["count.all.freq.txt"].splitlines())
    _, paths = _run(tmp_path, "gz", ["--gzip"], aggr_ckpt)
    assert not os.path.exists(paths[0])
    with gzip.open(paths[0] + ".gz", "rt") as f:  # BGZF is a valid multi-member gzip stream
        assert sorted(f.read().splitlines()) == sorted(ref_out["count.all.freq.txt"].splitlines())
