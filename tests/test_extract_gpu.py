"""Device feature extraction (include/ccsm.h ccsm_reads_*, csrc/extract.cu) against
  * the reference extractor's own output on the demo BAM (tests/golden/demo_callmods.npz, made by
    scripts/gen_golden.py from the unmodified reference), and
  * the numpy oracle (oracle/extract_numpy.py, itself pinned on those fixtures) on the whole demo and on synthetic
    reads covering the edge cases: reverse-strand records, N bases, reads shorter than a window, reads without a
    CpG, constant kinetics (zero scale), every --norm, --no_decode, multi-motif, soft-clip windows.
Integer outputs (site lists, base codes, npass, MM deltas, ML bytes) must be identical.  The float features are
np.around(.,6) values: identical except where a last-ulp difference of the float64 std (numpy sums pairwise, the
kernel uses the exact integer sums) crosses a rounding boundary -- at most 1e-6 apart, and the tests also bound
how often that may happen."""
import os

import numpy as np
import pytest
import torch

from ccsmeth_b200 import call_mods as cm
from ccsmeth_b200.bamio import BamReader
from ccsmeth_b200.extract_features import extract_opts, pack_reads
from ccsmeth_b200.models import ModelAttRNN
from oracle.extract_numpy import batch_read_features, extract_read
from tests.bamsynth import make_record, random_read
from tests.conftest import GOLDEN, load_npz

pytestmark = pytest.mark.gpu
DEMO = os.path.join(GOLDEN, "demo", "hg002.chr20_demo.hifi.bam")
PAIRS = (("kmer", "kmer"), ("kmer2", "kmer2"), ("kpass", "kpass"), ("kpass2", "kpass2"), ("ipd", "ipd"), ("pw", "pw"),
         ("ipd2", "ipd2"), ("pw2", "pw2"))


def _args(**kw):
    a = cm.build_parser().parse_args(["-i", DEMO, "-m", "x.ckpt", "-o", "out"])
    for k, v in kw.items():
        setattr(a, k, v)
    return a


@pytest.fixture(scope="module")
def model(ckpt_att2s):
    m = ModelAttRNN(21, 3, 2, 0, 256, is_npass=True, model_type="attbigru2s", device=0, precision="fp16x3")
    m.load_state_dict({k: torch.from_numpy(v) for k, v in ckpt_att2s.items()})
    return m.cuda(0).eval()


@pytest.fixture(scope="module")
def reads():
    return list(BamReader(DEMO))


def _oracle(reads, args, motifs):
    feats = []
    for i, r in enumerate(reads):
        rf = extract_read(r, motifs, args)
        if rf is not None and len(rf):
            feats.append((i, rf))
    return batch_read_features(feats, args.seq_len)


def _compare(model, reads, args, motifs, exact_floats=True):
    arrays, holeidx, locs = _oracle(reads, args, motifs)
    batch = pack_reads(reads, args)
    n = model.extract_reads(batch, extract_opts(args, motifs)) if len(batch) else 0
    assert n == len(locs)
    if n == 0:
        return 0
    site_read, site_loc = model.reads_sites()
    assert np.array_equal(np.asarray(batch.index)[site_read], holeidx)
    assert np.array_equal(site_loc, locs)
    dev = {k: v.cpu().numpy() for k, v in model.reads_features().items()}
    flips = 0
    for mine, ref in PAIRS:
        a, b = dev[mine], arrays[ref]
        assert a.shape == b.shape, mine
        if mine.startswith(("kmer", "kpass")) or exact_floats:
            assert np.array_equal(a, b), mine
        else:
            d = np.abs(a.astype(np.float64) - b.astype(np.float64))
            assert d.max() <= 1.5e-6, (mine, d.max())
            flips += int((a != b).sum())
    return flips


def test_demo_sites_and_features_match_the_reference_extractor(model, reads):
    g = load_npz("demo_callmods.npz")
    args = _args()
    batch = pack_reads(reads, args)
    n = model.extract_reads(batch, extract_opts(args, ["CG"]))
    assert n == 12691 == len(g["locs"])
    site_read, site_loc = model.reads_sites()
    assert np.array_equal(site_loc, g["locs"])
    assert np.array_equal(np.bincount(site_read, minlength=len(reads)), g["n_sites_per_read"])
    # the first three reads' feature arrays, as produced by the reference's extract_features_from_double_strand_read
    n0 = len(g["feat0.locs"])
    dev = {k: v.cpu().numpy() for k, v in model.reads_features(0, n0).items()}
    for mine, ref in (("kmer", "fkmer"), ("kmer2", "rkmer"), ("kpass", "fpass"), ("kpass2", "rpass"),
                      ("ipd", "fipd"), ("pw", "fpw"), ("ipd2", "ripd"), ("pw2", "rpw")):
        assert np.array_equal(dev[mine], g["feat0." + ref].astype(np.float32)), mine


def test_demo_all_reads_match_the_oracle(model, reads):
    assert _compare(model, reads, _args(), ["CG"]) == 0


def test_demo_chunked_feature_ranges(model, reads):
    args = _args()
    batch = pack_reads(reads[:20], args)
    n = model.extract_reads(batch, extract_opts(args, ["CG"]))
    full = {k: v.cpu().numpy() for k, v in model.reads_features().items()}
    part = {k: v.cpu().numpy() for k, v in model.reads_features(100, 333).items()}
    for k in full:
        assert np.array_equal(full[k][100:433], part[k])
    with pytest.raises(Exception):
        model.reads_features(n - 5, 10)


@pytest.mark.parametrize("norm", ["zscore", "min-mean", "min-max", "none", "mad"])
@pytest.mark.parametrize("no_decode", [False, True])
def test_synthetic_reads_every_norm(model, norm, no_decode):
    rng = np.random.default_rng(7)
    recs = [random_read(rng, "r%d" % i, int(n))[0] for i, n in enumerate(rng.integers(30, 3000, 40))]
    flips = _compare(model, recs, _args(norm=norm, no_decode=no_decode), ["CG"], exact_floats=False)
    assert flips <= 2


def test_synthetic_edge_cases(model):
    rng = np.random.default_rng(11)
    recs = [
        random_read(rng, "short20", 20)[0],                      # shorter than one window: no sites
        random_read(rng, "short21", 21, p_cg=1.0)[0],            # exactly one window long
        random_read(rng, "len22", 22, p_cg=1.0)[0],
        random_read(rng, "nocg", 500, no_cg=True)[0],
        random_read(rng, "manyN", 800, p_n=0.2)[0],
        random_read(rng, "const_ipd", 600, const_sig=0)[0],      # zero variance -> all-zero feature
        random_read(rng, "const_rpw", 600, const_sig=3)[0],
        random_read(rng, "rev", 1500, reverse=True, flag=16)[0],  # stored reverse-complemented
        random_read(rng, "odd_len", 1001)[0],
        random_read(rng, "dense", 400, p_cg=0.5)[0],
    ]
    # CpG at the very edges: only sites with a full window on both strands survive
    edge = "CG" + "A" * 30 + "CG" + "T" * 8 + "CG" + "ACGT" * 3 + "CG"
    k = np.arange(len(edge), dtype=np.uint8)
    recs.append(make_record("edges", edge, k, k[::-1].copy(), k * 3 % 251, k * 7 % 241))
    assert _compare(model, recs, _args(), ["CG"], exact_floats=False) <= 1


def test_synthetic_align_mode_reverse_and_softclips(model):
    rng = np.random.default_rng(13)
    recs = []
    for i in range(12):
        n = int(rng.integers(200, 2500))
        lclip, rclip = int(rng.integers(0, 60)), int(rng.integers(0, 60))
        rev = bool(i % 2)
        recs.append(random_read(rng, "a%d" % i, n, reverse=rev, flag=(16 if rev else 0),
                                cigar=((4, lclip), (0, n - lclip - rclip), (4, rclip)), mapq=60)[0])
    recs.append(random_read(rng, "unmapped", 500, flag=4)[0])          # filtered out in align mode
    recs.append(random_read(rng, "lowq", 500, flag=0, cigar=((0, 500),), mapq=0)[0])
    recs.append(random_read(rng, "secondary", 500, flag=256, cigar=((0, 500),), mapq=60)[0])
    args = _args(mode="align")
    assert _compare(model, recs, args, ["CG"], exact_floats=False) <= 1
    args2 = _args(mode="align", skip_unmapped="no")
    assert _compare(model, recs, args2, ["CG"], exact_floats=False) <= 1


def test_multi_motif_and_mod_loc(model):
    rng = np.random.default_rng(17)
    recs = [random_read(rng, "m%d" % i, int(n))[0] for i, n in enumerate(rng.integers(100, 1500, 10))]
    motifs = cm.get_motif_seqs("CHG")
    assert sorted(motifs) == ["CAG", "CCG", "CTG"]
    assert _compare(model, recs, _args(motifs="CHG"), motifs, exact_floats=False) <= 1
    gatc = cm.get_motif_seqs("GATC")
    assert _compare(model, recs, _args(motifs="GATC", mod_loc=1), gatc, exact_floats=False) <= 1


def test_device_mm_ml_prob1_match_the_host_converters(model, reads):
    """prob_1_norm, MM deltas and ML bytes computed on the device == the reference-shaped host converters
    (call_mods.convert_*, pinned on the reference's own converters in tests/test_demo_cpu.py) applied to the
    device's raw probabilities."""
    args = _args()
    sub = reads[:25]
    batch = pack_reads(sub, args)
    n = model.extract_reads(batch, extract_opts(args, ["CG"]))
    site_read, site_loc = model.reads_sites()
    h0 = (torch.randn(6, n, 256, generator=torch.Generator().manual_seed(5)),
          torch.randn(6, n, 256, generator=torch.Generator().manual_seed(6)))
    res = model.reads_forward(h0=h0)
    p = res["probs"]
    prob1 = np.round(p[:, 1] / (p[:, 0] + p[:, 1]), 6)
    assert np.array_equal(res["prob1"], prob1)
    assert np.array_equal(res["ml"], cm.convert_probs_to_mltag(prob1))
    for r in np.unique(site_read):
        sel = site_read == r
        fwd = np.frombuffer(sub[batch.index[r]].get_forward_sequence().encode(), dtype=np.uint8)
        assert np.array_equal(res["mm"][sel], cm.convert_locs_to_mmtag(site_loc[sel].astype(np.int64), fwd))
    # and the forward itself: same features through the 16-tensor entry give the same probabilities
    f = model.reads_features()
    order = ("kmer", "kpass", "ipd", None, "pw", None, None, None)
    a = [f[k] if k else torch.zeros(1) for k in order] + [f[k + "2"] if k else torch.zeros(1) for k in order]
    _, probs2 = model(*a, h0=h0)
    assert np.array_equal(probs2.cpu().numpy(), p)


def test_mm_counts_reverse_strand_records(model):
    rng = np.random.default_rng(19)
    recs = [random_read(rng, "rv%d" % i, 900, reverse=True, flag=16)[0] for i in range(4)]
    args = _args()
    batch = pack_reads(recs, args)
    n = model.extract_reads(batch, extract_opts(args, ["CG"]))
    assert n > 0
    site_read, site_loc = model.reads_sites()
    res = model.reads_forward(h0=(torch.zeros(6, n, 256), torch.zeros(6, n, 256)))
    for r in np.unique(site_read):
        sel = site_read == r
        fwd = np.frombuffer(recs[r].get_forward_sequence().encode(), dtype=np.uint8)
        assert np.array_equal(res["mm"][sel], cm.convert_locs_to_mmtag(site_loc[sel].astype(np.int64), fwd))


def test_empty_and_bad_batches(model):
    args = _args()
    rng = np.random.default_rng(23)
    none = pack_reads([random_read(rng, "x", 300, no_cg=True)[0]], args)
    assert model.extract_reads(none, extract_opts(args, ["CG"])) == 0
    assert model.reads_forward()["prob1"].shape == (0,)
    batch = pack_reads([random_read(rng, "y", 300)[0]], args)
    batch.descs["fi_off"][0] = 10 ** 9  # outside the blob
    with pytest.raises(Exception, match="outside the blob"):
        model.extract_reads(batch, extract_opts(args, ["CG"]))
    with pytest.raises(Exception, match="ACGT"):
        model.extract_reads(pack_reads([random_read(rng, "z", 300)[0]], args),
                            {"mod_loc": 0, "norm": 0, "decode": 1, "motifs": ["CN"]})
