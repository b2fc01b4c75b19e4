"""Host side of call_mods on the reference's demo BAM (tests/golden/demo/, a copy of the reference's
demo/hg002.chr20_demo.hifi.bam) against fixtures produced by the reference's own extractor / tag converters
(scripts/gen_golden.py gen_demo).  No GPU needed."""
import os

import numpy as np
import pytest

from ccsmeth_b200 import call_mods as cm
from ccsmeth_b200.bamio import BamReader, BamWriter, add_pg_line
from ccsmeth_b200.extract_features import READ_SEQ_4BIT, pack_reads
from oracle.extract_numpy import CODE2FRAMES, batch_read_features, extract_read, to_feature_rows
from tests.conftest import GOLDEN, load_npz

DEMO = os.path.join(GOLDEN, "demo", "hg002.chr20_demo.hifi.bam")


@pytest.fixture(scope="module")
def golden():
    return load_npz("demo_callmods.npz")


@pytest.fixture(scope="module")
def reads():
    return list(BamReader(DEMO))


def _args(**kw):
    a = cm.build_parser().parse_args(["-i", DEMO, "-m", "x.ckpt", "-o", "out"])
    for k, v in kw.items():
        setattr(a, k, v)
    return a


def test_bam_reader_sees_the_demo(reads, golden):
    assert len(reads) == 116
    assert [r.query_name for r in reads] == list(golden["names"])
    r = reads[0]
    assert r.is_unmapped and not r.is_reverse  # the demo is unaligned CCS reads
    assert len(r.get_tag("fi")) == r.l_seq == len(r.query_sequence)
    assert r.get_tag("fi").dtype == np.uint8 and 1 <= r.get_tag("fn") <= 60
    with pytest.raises(KeyError):
        r.get_tag("XX")


def test_codecv1_table():
    # reference process_utils.py:426-449
    assert CODE2FRAMES[63] == 63 and CODE2FRAMES[64] == 64 and CODE2FRAMES[127] == 190
    assert CODE2FRAMES[128] == 192 and CODE2FRAMES[191] == 444 and CODE2FRAMES[192] == 448 and CODE2FRAMES[255] == 952


def test_extraction_matches_reference_extractor(reads, golden):
    args = _args()
    feats = [(i, extract_read(r, ["CG"], args)) for i, r in enumerate(reads[:3])]
    arrays, holeidx, locs = batch_read_features(feats, 21)
    assert np.array_equal(holeidx, golden["feat0.holeidx"])
    assert np.array_equal(locs, golden["feat0.locs"])
    for mine, ref in (("kmer", "fkmer"), ("kmer2", "rkmer"), ("kpass", "fpass"), ("kpass2", "rpass"),
                      ("ipd", "fipd"), ("pw", "fpw"), ("ipd2", "ripd"), ("pw2", "rpw")):
        assert np.array_equal(arrays[mine], golden["feat0." + ref].astype(np.float32)), mine


def test_site_counts_per_holebatch(reads, golden):
    args = _args()
    counts = []
    for b0 in range(0, len(reads), 50):
        n = 0
        for r in reads[b0:b0 + 50]:
            rf = extract_read(r, ["CG"], args)
            n += 0 if rf is None else len(rf)
        counts.append(n)
    assert counts == list(golden["site_counts_per_batch"]) == [5275, 5568, 1848]


def test_feature_rows_have_reference_shape(reads):
    args = _args()
    rf = extract_read(reads[0], ["CG"], args)
    rows = to_feature_rows(rf, args)
    assert len(rows) == len(rf) and len(rows[0]) == 22
    r0 = rows[0]
    assert r0[0] == "." and r0[1] == -1 and r0[3] == reads[0].query_name and len(r0[5]) == 21 and len(r0[13]) == 21
    assert r0[5][10:12] == "CG" and r0[13][10:12] == "CG"  # both strands centred on the CpG


def test_mm_ml_conversion_matches_reference(reads, golden):
    off = 0
    for r, n in zip(reads, golden["n_sites_per_read"]):
        if n == 0:
            continue
        locs = golden["locs"][off:off + n]
        fwd = np.frombuffer(r.get_forward_sequence().encode(), dtype=np.uint8)
        assert np.array_equal(cm.convert_locs_to_mmtag(locs, fwd), golden["mm"][off:off + n])
        assert np.array_equal(cm.convert_probs_to_mltag(golden["prob1"][off:off + n]), golden["ml"][off:off + n])
        off += n
    with pytest.raises(AssertionError):
        cm.convert_locs_to_mmtag(np.array([0, 1]), np.frombuffer(b"AAAA", dtype=np.uint8))


def test_ml_edge_values():
    assert list(cm.convert_probs_to_mltag([0.0, 0.5, 0.999999, 1.0])) == [0, 128, 255, 255]


def test_modbam_write_read_roundtrip(tmp_path, reads, golden):
    out = str(tmp_path / "o.modbam.bam")
    rd = BamReader(DEMO)
    wr = BamWriter(out, add_pg_line(rd.header_text, "t", "cmd"), rd.references)
    off = 0
    for r, n in zip(reads[:10], golden["n_sites_per_read"][:10]):
        pred = (golden["locs"][off:off + n], golden["prob1"][off:off + n]) if n else None
        raw, flag = cm.tag_read(r, pred, rm_pulse=True)
        assert flag == (1 if n else 0)
        wr.write_raw(raw)
        off += n
    wr.close()
    back = BamReader(out)
    assert "@PG\tPN:ccsmeth\tID:ccsmeth" in back.header_text
    recs = list(back)
    assert len(recs) == 10
    off = 0
    for a, b, n in zip(reads[:10], recs, golden["n_sites_per_read"][:10]):
        assert b.query_name == a.query_name and b.query_sequence == a.query_sequence and b.flag == a.flag
        for t in ("fi", "fp", "ri", "rp"):
            assert not b.has_tag(t)  # pulse tags dropped unless --keep_pulse (_bam2modbam.py:217-218)
        assert b.get_tag("fn") == a.get_tag("fn") and b.get_tag("zm") == a.get_tag("zm")
        if n:
            mm = b.get_tag("MM")
            assert mm.startswith("C+m?,") and mm.endswith(";")
            assert [int(x) for x in mm[5:-1].split(",")] == list(golden["mm"][off:off + n])
            assert np.array_equal(b.get_tag("ML"), golden["ml"][off:off + n])
        else:
            assert not b.has_tag("MM")
        off += n


def test_pack_reads_points_at_the_right_bytes(reads):
    """The descriptors handed to the device extractor must address each read's packed sequence and kinetics
    arrays inside the blob of raw records."""
    args = _args()
    batch = pack_reads(reads[:7], args)
    assert len(batch) == 7 and batch.index == list(range(7)) and batch.descs.dtype.itemsize == 80
    nib = "=ACMGRSVTWYHKDBN"
    for d, r in zip(batch.descs, reads[:7]):
        n = int(d["len"])
        assert n == r.l_seq and d["flags"] == READ_SEQ_4BIT and (d["win_lo"], d["win_hi"]) == (0, n)
        assert (d["fn"], d["rn"]) == (r.get_tag("fn"), r.get_tag("rn"))
        for key, tag in (("fi_off", "fi"), ("ri_off", "ri"), ("fp_off", "fp"), ("rp_off", "rp")):
            assert np.array_equal(batch.blob[d[key]:d[key] + n], r.get_tag(tag))
        packed = batch.blob[d["seq_off"]:d["seq_off"] + (n + 1) // 2]
        seq = "".join(nib[b >> 4] + nib[b & 15] for b in packed[:16])
        assert seq[:32] == r.query_sequence[:32]


def test_pack_reads_skips_reads_without_kinetics(reads):
    args = _args()
    r = reads[0]
    stripped = type(r)(r.with_tags({"fi"}))
    batch = pack_reads([stripped, reads[1]], args)
    assert batch.index == [1]


def test_native_bgzf_codec_roundtrip(tmp_path):
    """libccsm's BGZF thread team (include/ccsm.h ccsm_bgzf_*) against Python's zlib path, both directions."""
    a = list(BamReader(DEMO, threads=1))
    b = list(BamReader(DEMO, threads=4))
    assert [r.raw for r in a] == [r.raw for r in b]
    rd = BamReader(DEMO)
    for strategy in ("zlib", "rle"):
        outs = []
        for th in (1, 4):
            out = str(tmp_path / ("%s%d.bam" % (strategy, th)))
            wr = BamWriter(out, rd.header_text, rd.references, threads=th, strategy=strategy)
            for r in a[:40]:
                wr.write_raw(r.raw)
            wr.close()
            outs.append(out)
        b1, b4 = open(outs[0], "rb").read(), open(outs[1], "rb").read()
        if strategy == "zlib":
            assert b1 == b4  # same zlib, same level: identical bytes
        else:
            # Python's Z_RLE against the library's own run-length encoder: same token stream, near-identical sizes
            assert abs(len(b1) - len(b4)) <= 0.002 * len(b1)
        for out in outs:
            assert [r.raw for r in BamReader(out, threads=4)] == [r.raw for r in a[:40]]
            assert [r.raw for r in BamReader(out, threads=1)] == [r.raw for r in a[:40]]


def test_motif_expansion():
    assert cm.get_motif_seqs("CG") == ["CG"]
    assert sorted(cm.get_motif_seqs("CHG")) == ["CAG", "CCG", "CTG"]
