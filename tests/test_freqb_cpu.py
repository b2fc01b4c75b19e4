"""call_freqb host half without a GPU: reference chunks, FASTA reader, the native MM/ML -> reference projection
(ccsm_bam_modcalls) against the oracle restatement, and the region pileup builder + numpy pileup oracle against the
output of the reference's own region worker (tests/golden/freqb/reference_outputs.npz)."""
import argparse
import io
import os

import numpy as np
import pytest

from ccsmeth_b200 import _lib, call_freqb as cf
from ccsmeth_b200.bamio import BamReader
from ccsmeth_b200.bamstream import BamPieceReader
from oracle import freqb_numpy, pileup_numpy
from tests.conftest import GOLDEN

D = os.path.join(GOLDEN, "freqb")
BAM, FA = os.path.join(D, "synth.aligned.modbam.bam"), os.path.join(D, "synth.fa")


@pytest.fixture(scope="module")
def ref_out():
    with np.load(os.path.join(D, "reference_outputs.npz")) as z:
        return {k: z[k].tobytes().decode("ascii") for k in z.files}


def _args(**kw):
    a = cf.build_parser().parse_args(["--input_bam", BAM, "--ref", FA, "-o", "out", "--chunk_len", "10000"])
    for k, v in kw.items():
        setattr(a, k, v)
    return a


def test_fasta_and_chunks_match_the_reference(ref_out):
    contigs = cf.read_fasta(FA)
    assert sorted(contigs) == ["chrA", "chrB"] and contigs["chrA"] == contigs["chrA"].upper()
    chunks = cf.get_reference_chunks(contigs, None, 10000, "CG")
    assert "".join("%s\t%d\t%d\n" % c for c in chunks) == ref_out["chunks"]
    assert ("chrA", 0, 10001) in chunks and ("chrA", 10001, 20000) in chunks  # the CG on the boundary moved it


def _native_calls(**kw):
    o = _lib.ModcallOpts(kw.get("mapq", 1), int(kw.get("no_supplementary", False)), kw.get("base_clip", 0),
                         kw.get("hap_tag", "HP").encode(), kw.get("identity", 0.0), 0, 0, None, None, None)
    rd = BamPieceReader(BAM, _lib.BamFilter(0, 0, 0, 0, 0), threads=2, piece_bytes=40000, align_to=1)
    calls = cf.ModCalls()
    used = 0
    n_pieces = 0
    for piece in rd:
        used += calls.add_piece(piece, o)
        n_pieces += 1
    assert n_pieces > 1
    return calls.arrays(), used


@pytest.mark.parametrize("kw", [{}, {"base_clip": 15, "no_supplementary": True, "identity": 0.995, "mapq": 20},
                                {"hap_tag": "XX"}])
def test_native_modcalls_match_the_oracle(kw):
    (rid, pos, ml, hap, strand), used = _native_calls(**kw)
    exp, n_used = [], 0
    for rec in BamReader(BAM):
        c = freqb_numpy.read_calls(rec, **kw)
        if c is not None:
            n_used += 1
            exp += c
    assert used == n_used and len(exp) == len(rid) > 1000
    got = list(zip(rid.tolist(), pos.tolist(), ml.tolist(), hap.tolist(), strand.tolist()))
    assert got == exp  # same calls in the same order (file order, then along the read)
    if kw.get("hap_tag") == "XX":
        assert not hap.any()
    else:
        assert set(np.unique(hap)) == {0, 1, 2}


def test_native_refsites_all_matches_the_oracle():
    """--refsites_all: every aligned pair (insertions, soft clips, deletions included in the clip arithmetic), zero
    calls at uncalled reference motif sites, reads without MM/ML still counted."""
    import ctypes
    contigs = cf.read_fasta(FA)
    names = ["chrA", "chrB"]
    masks = [cf.motif_site_masks(contigs[n], ["CG"], 0) for n in names]
    # the masks themselves against a plain scan
    for n, (f, r) in zip(names, masks):
        seq = contigs[n]
        assert list(np.nonzero(f)[0]) == [i for i in range(len(seq) - 1) if seq[i:i + 2] == "CG"]
        assert list(np.nonzero(r)[0]) == [i + 1 for i in range(len(seq) - 1) if seq[i:i + 2] == "CG"]
    lens = [len(contigs[n]) for n in names]
    ref_off = np.concatenate(([0], np.cumsum(lens))).astype(np.int64)
    sf = np.concatenate([m[0] for m in masks] + [np.zeros(1, np.uint8)])
    sr = np.concatenate([m[1] for m in masks] + [np.zeros(1, np.uint8)])
    for clip in (0, 40):
        o = _lib.ModcallOpts(1, 0, clip, b"HP", 0.0, 1, 2, ref_off.ctypes.data, sf.ctypes.data, sr.ctypes.data)
        rd = BamPieceReader(BAM, _lib.BamFilter(0, 0, 0, 0, 0), threads=2, piece_bytes=40000, align_to=1)
        calls = cf.ModCalls()
        for piece in rd:
            calls.add_piece(piece, o)
        rid, pos, ml, hap, strand = calls.arrays()
        exp = []
        for rec in BamReader(BAM):
            if rec.is_unmapped:
                continue
            sets = (set(np.nonzero(masks[rec.ref_id][0])[0].tolist()), set(np.nonzero(masks[rec.ref_id][1])[0].tolist()))
            c = freqb_numpy.read_calls(rec, base_clip=clip, refsites=sets)
            if c is not None:
                exp += c
        got = list(zip(rid.tolist(), pos.tolist(), ml.tolist(), hap.tolist(), strand.tolist()))
        assert got == exp and (strand >= 2).sum() > 100


def test_malformed_mm_tags_yield_no_calls():
    from tests.bamsynth import make_record
    import struct
    seq = "ACGTCGCGTTCGAACCGG" * 3
    def rec(mm, ml):
        tags = b"MMZ" + mm.encode() + b"\x00" + b"MLBC" + struct.pack("<I", len(ml)) + bytes(ml)
        return make_record("r", seq, None, None, None, None, fn=None, flag=0, cigar=((0, len(seq)),), extra_tags=tags,
                           mapq=60, ref_id=0, pos=10)
    good = rec("C+m?,0,1;", [10, 200])
    assert [c[1:3] for c in freqb_numpy.read_calls(good)] == [(11, 10), (16, 200)]
    for bad in (rec("C+m?,0,1;", [10]), rec("C+m?,0,400;", [1, 2]), rec("C+h?,0,1;", [1, 2]), rec("C+m?;", [])):
        assert freqb_numpy.read_calls(bad) == []
    # the native walker agrees on each of them
    import ctypes
    from ccsmeth_b200.bamstream import REC_DTYPE
    from ccsmeth_b200.extract_features import READ_DTYPE
    lib = _lib.load()
    for r in (good, rec("C+m?,0,1;", [10]), rec("C+m?,0,400;", [1, 2]), rec("C+h?,0,1;", [1, 2]), rec("C+m?;", [])):
        buf = np.frombuffer(struct.pack("<i", len(r.raw)) + r.raw, dtype=np.uint8)
        recs, descs = np.zeros(4, REC_DTYPE), np.zeros(4, READ_DTYPE)
        nr, nd, cons = ctypes.c_int32(0), ctypes.c_int32(0), ctypes.c_int64(0)
        f = _lib.BamFilter(0, 0, 0, 0, 0)
        _lib.check(lib.ccsm_bam_index(buf.ctypes.data, len(buf), ctypes.byref(f), recs.ctypes.data, 4, descs.ctypes.data,
                                      ctypes.byref(nr), ctypes.byref(nd), ctypes.byref(cons)))
        o = _lib.ModcallOpts(1, 0, 0, b"HP", 0.0, 0, 0, None, None, None)
        out = [np.zeros(64, dt) for dt in (np.int32, np.int32, np.uint8, np.uint8, np.uint8)]
        used = ctypes.c_int32(0)
        n = lib.ccsm_bam_modcalls(buf.ctypes.data, recs.ctypes.data, 1, ctypes.byref(o), *[a.ctypes.data for a in out], 64,
                                  ctypes.byref(used))
        exp = freqb_numpy.read_calls(r)
        assert n == len(exp) and [(int(out[1][k]), int(out[2][k])) for k in range(n)] == [c[1:3] for c in exp]


@pytest.mark.parametrize("tag,kw", [("count", {}), ("count_cf3", {"prob_cf": 0.3}),
                                    ("count_cf3_noamb", {"prob_cf": 0.3, "no_amb_cov": True}),
                                    ("count_nocomb", {"no_comb": True}), ("count_refsites", {"refsites_only": True}),
                                    ("count_clip_nosupp_ident", {"base_clip": 15, "no_supplementary": True,
                                                                 "identity": 0.995, "mapq": 20}),
                                    ("count_refsites_all", {"refsites_all": True}),
                                    ("count_refsites_all_clip_nocomb", {"refsites_all": True, "base_clip": 40,
                                                                        "no_comb": True})])
def test_host_chain_with_the_pileup_oracle_reproduces_the_reference_files(ref_out, tag, kw):
    """Native projection -> region pileups -> (numpy pileup oracle instead of the device) -> text lines ==
    the reference region worker's output, byte for byte, in count mode."""
    args = _args(**kw)

    class OracleModel:  # stands in for AggrAttRNN.pileup_begin / pileup_finish
        def pileup_begin(self, refpos, ptr, ml, hap, **k):
            self.a = (refpos, ptr, ml, hap, k)
            return (0, 0, 0)

        def pileup_finish(self, h0, with_kind=False):
            refpos, ptr, ml, hap, k = self.a
            out = pileup_numpy.call_region(refpos, ptr, ml, hap, None, call_mode="count", cov_cf=k["cov_cf"],
                                           prob_cf=k["prob_cf"], no_amb_cov=k["no_amb_cov"], no_hap=k["no_hap"])
            n = len(refpos)
            cov = np.where(np.isnan(out[..., 0]), -1, out[..., 0]).astype(np.int32)
            kind = np.zeros((3, n), np.uint8)
            for g in range(3):
                for i in range(n):
                    if cov[g, i] >= 0:
                        sel = slice(ptr[i], ptr[i + 1])
                        probs = [pileup_numpy.cal_mod_prob(int(v)) for v, h in zip(ml[sel], hap[sel]) if g == 0 or h == g]
                        filt = sum(1 for p in probs if not abs(p - (1 - p)) < k["prob_cf"])
                        kind[g, i] = 1 if (k["no_amb_cov"] or filt == len(probs)) else 2
            return cov, np.nan_to_num(out[..., 1]), np.nan_to_num(out[..., 2]), kind

    contigs = cf.read_fasta(FA)
    bufs = [io.StringIO() for _ in range(3)]
    for _, *beds in cf.iter_region_results(args, OracleModel(), contigs, BAM):
        for g in range(3):
            for item in beds[g]:
                cf.write_one_line(item, bufs[g], False)
    for g, name in enumerate(("all", "hp1", "hp2")):
        assert bufs[g].getvalue() == ref_out["%s.%s.freq.txt" % (tag, name)], (tag, name)
    if tag == "count":
        # the same through many small pieces: references are flushed as soon as the sorted stream has moved past them
        b = io.StringIO()
        for _, *beds in cf.iter_region_results(args, OracleModel(), contigs, BAM, piece_bytes=30000):
            for item in beds[0]:
                cf.write_one_line(item, b, False)
        assert b.getvalue() == ref_out["count.all.freq.txt"]
        # reference chunks dealt round-robin to three ranks: disjoint, and together the one-rank result
        parts = []
        for rank in range(3):
            b = io.StringIO()
            for _, *beds in cf.iter_region_results(args, OracleModel(), contigs, BAM, rank, 3):
                for item in beds[0]:
                    cf.write_one_line(item, b, False)
            parts.append(b.getvalue().splitlines())
        assert all(parts) and sorted(sum(parts, [])) == sorted(ref_out["count.all.freq.txt"].splitlines())


def test_write_lines_prints_what_write_one_line_prints():
    """The list writer must keep the reference's str() formatting of every value type (int, float, np.float64,
    np.float32 -- the last one prints differently through an empty format spec)."""
    rng = np.random.default_rng(0)
    items = [("c", 0, "+", 5, 0, 0.0), ("c", 1, "+", 5, 5, 1.0), ("c", 2, "-", 7, np.float32(0.0), np.float32(1.0))]
    for i in range(3000):
        cov, f = int(rng.integers(1, 80)), float(rng.random())
        items.append(("chr1", 3 * i, "+", cov, int(f * cov), int(f * cov) / cov))
        items.append(("chr1", 3 * i + 1, "-", cov, np.float64(round(cov * f, 2)), f))
        items.append(("chrX", 3 * i + 2, "+", cov, np.float32(round(cov * np.float32(f), 2)), np.float32(round(f, 6))))
    for bed in (False, True):
        a, b = io.StringIO(), io.StringIO()
        for it in items:
            cf.write_one_line(it, a, bed)
        cf.write_lines(items, b, bed)
        assert a.getvalue() == b.getvalue()


def test_unsorted_modbam_is_refused_and_sorting_it_restores_the_result(tmp_path, ref_out):
    import struct
    """The streaming region caller needs coordinate order (the reference fetches through the index and fails loudly on
    other input): an unsorted modbam raises, and after ccsmeth_b200.bamsort the golden output comes back."""
    from ccsmeth_b200 import bamsort
    from ccsmeth_b200.bamio import BamReader, BamWriter
    rd = BamReader(BAM)
    recs = [r.raw for r in rd]
    rng = np.random.default_rng(1)
    shuffled = str(tmp_path / "shuffled.bam")
    w = BamWriter(shuffled, rd.header_text.replace("SO:coordinate", "SO:unsorted"), rd.references)
    for i in rng.permutation(len(recs)):
        w.write_raw(recs[i])
    w.close()
    args = _args()
    contigs = cf.read_fasta(FA)

    class NoModel:
        def pileup_begin(self, *a, **k):
            raise AssertionError("unsorted input must be refused before any region is called")

    with pytest.raises(ValueError, match="not coordinate-sorted"):
        for _ in cf.iter_region_results(args, NoModel(), contigs, shuffled, piece_bytes=30000):
            pass
    fixed = str(tmp_path / "fixed.bam")
    assert bamsort.sort_and_index(shuffled, fixed, threads=2) == len(recs)
    got = [r.raw for r in BamReader(fixed)]
    key = lambda raw: ((struct.unpack_from("<i", raw, 0)[0] & 0xFFFFFFFF) << 32) | ((struct.unpack_from("<i", raw, 4)[0] + 1) << 1) | \
        (1 if struct.unpack_from("<H", raw, 14)[0] & 16 else 0)
    assert [key(r) for r in got] == sorted(key(r) for r in recs) and sorted(got) == sorted(recs)
