"""Native piece-wise BAM streaming (ccsmeth_b200/bamstream.py over ccsm_bgzf_* / ccsm_bam_* in libccsm) against the
record-level Python implementation (bamio.BamRecord, extract_features.pack_reads, call_mods.tag_read), which is itself
pinned on the reference's fixtures in tests/test_demo_cpu.py.  No GPU needed: these are host helpers."""
import os

import numpy as np
import pytest

from ccsmeth_b200 import _lib, call_mods as cm
from ccsmeth_b200.bamio import BamReader, BamWriter
from ccsmeth_b200.bamstream import BamPieceReader, tag_records
from ccsmeth_b200.extract_features import pack_reads
from tests.bamsynth import random_read
from tests.conftest import GOLDEN, load_npz

DEMO = os.path.join(GOLDEN, "demo", "hg002.chr20_demo.hifi.bam")


def _args(**kw):
    a = cm.build_parser().parse_args(["-i", DEMO, "-m", "x.ckpt", "-o", "out"])
    for k, v in kw.items():
        setattr(a, k, v)
    return a


def _filter(args):
    return _lib.BamFilter(1 if args.mode == "align" else 0, args.mapq, 1 if args.no_supplementary else 0,
                          1 if args.skip_unmapped == "yes" else 0, 0,
                          identity=getattr(args, "identity", 0.0) if args.mode == "align" else 0.0)


def _check_pieces(path, args, piece_bytes, align_to):
    recs_py = list(BamReader(path))
    rd = BamPieceReader(path, _filter(args), threads=3, piece_bytes=piece_bytes, align_to=align_to)
    k = 0
    pieces = list(rd)
    for pi, p in enumerate(pieces):
        n = len(p.recs)
        assert p.first == k
        if pi + 1 < len(pieces):
            assert n % align_to == 0
        b = pack_reads(recs_py[k:k + n], args)
        assert len(b) == len(p.descs)
        assert list(np.nonzero(p.recs["read_idx"] >= 0)[0]) == b.index
        for name in ("len", "fn", "rn", "flags", "win_lo", "win_hi"):
            assert np.array_equal(b.descs[name], p.descs[name]), name
        for name in ("seq_off", "fi_off", "ri_off", "fp_off", "rp_off"):
            for i in range(len(b)):
                L = int(b.descs["len"][i])
                w = (L + 1) // 2 if name == "seq_off" else L
                assert np.array_equal(b.blob[b.descs[name][i]:b.descs[name][i] + w],
                                      p.buf[p.descs[name][i]:p.descs[name][i] + w])
        k += n
    assert k == len(recs_py)
    return pieces, recs_py


@pytest.mark.parametrize("piece_bytes,align_to", [(48 << 20, 50), (1 << 20, 50), (300000, 7), (70000, 1)])
def test_pieces_cover_the_demo_like_the_record_reader(piece_bytes, align_to):
    _check_pieces(DEMO, _args(), piece_bytes, align_to)


def test_retagging_matches_the_record_level_writer():
    g = load_npz("demo_callmods.npz")
    pieces, recs_py = _check_pieces(DEMO, _args(), 1 << 20, 50)
    off = 0
    k = 0
    for p in pieces:
        n = len(p.recs)
        counts = g["n_sites_per_read"][k:k + n]
        site_begin = np.concatenate(([0], np.cumsum(counts))).astype(np.int64)
        ns = int(site_begin[-1])
        mm = g["mm"][off:off + ns].astype(np.int32)
        ml = g["ml"][off:off + ns].astype(np.uint8)
        for keep in (False, True):
            out, with_mm = tag_records(p, p.recs, keep, site_begin, mm, ml)
            exp = []
            for j, r in enumerate(recs_py[k:k + n]):
                s, e = site_begin[j], site_begin[j + 1]
                pred = (np.zeros(e - s, dtype=np.int64), np.zeros(e - s, dtype=np.float32), mm[s:e], ml[s:e]) if e > s else None
                raw, _ = cm.tag_read(r, pred, rm_pulse=not keep)
                exp.append(len(raw).to_bytes(4, "little") + raw)
            assert bytes(out) == b"".join(exp)
            assert with_mm == int((counts > 0).sum())
        off += ns
        k += n


def test_retagged_pieces_written_through_the_bgzf_writer_read_back(tmp_path):
    g = load_npz("demo_callmods.npz")
    rd = BamPieceReader(DEMO, _filter(_args()), threads=3, piece_bytes=1 << 20, align_to=50)
    out = str(tmp_path / "o.bam")
    wr = BamWriter(out, rd.header_text, rd.references, threads=3)
    off = k = 0
    for p in rd:
        counts = g["n_sites_per_read"][k:k + len(p.recs)]
        sb = np.concatenate(([0], np.cumsum(counts))).astype(np.int64)
        ns = int(sb[-1])
        data, _ = tag_records(p, p.recs, False, sb, g["mm"][off:off + ns].astype(np.int32), g["ml"][off:off + ns])
        wr.bg.write(data)  # a uint8 numpy array, as the call_mods writer thread passes it
        off += ns
        k += len(p.recs)
    wr.close()
    back = list(BamReader(out, threads=2))
    assert [r.query_name for r in back] == list(g["names"])
    off = 0
    for r, n in zip(back, g["n_sites_per_read"]):
        if n:
            assert [int(x) for x in r.get_tag("MM")[5:-1].split(",")] == list(g["mm"][off:off + n])
            assert np.array_equal(r.get_tag("ML"), g["ml"][off:off + n])
        else:
            assert not r.has_tag("MM")
        assert not r.has_tag("fi")
        off += n


def test_index_applies_align_mode_filters_and_softclip_windows(tmp_path):
    rng = np.random.default_rng(3)
    recs = []
    for i in range(9):
        n = int(rng.integers(100, 900))
        lc, rc = int(rng.integers(0, 30)), int(rng.integers(0, 30))
        rev = bool(i % 2)
        recs.append(random_read(rng, "a%d" % i, n, reverse=rev, flag=16 if rev else 0,
                                cigar=((5, 3), (4, lc), (0, n - lc - rc), (4, rc)), mapq=60)[0])
    recs.append(random_read(rng, "unmapped", 300, flag=4)[0])
    recs.append(random_read(rng, "lowq", 300, flag=0, cigar=((0, 300),), mapq=0)[0])
    recs.append(random_read(rng, "dup", 300, flag=1024, cigar=((0, 300),), mapq=60)[0])
    # CIGAR identity (--identity, reference extract_features.py:283-286): 250 / 300 and 290 / 300
    recs.append(random_read(rng, "lowid", 300, flag=0, cigar=((0, 200), (1, 50), (0, 50)), mapq=60)[0])
    recs.append(random_read(rng, "highid", 300, flag=0, cigar=((7, 290), (8, 10)), mapq=60)[0])
    path = str(tmp_path / "syn.bam")
    wr = BamWriter(path, "@HD\tVN:1.6\n@SQ\tSN:chr1\tLN:100000\n", [("chr1", 100000)])
    for r in recs:
        wr.write_raw(r.raw)
    wr.close()
    for kw in ({"mode": "align"}, {"mode": "align", "skip_unmapped": "no"}, {"mode": "denovo"},
               {"mode": "align", "identity": 0.9}, {"mode": "denovo", "identity": 0.9}):
        pieces, _ = _check_pieces(path, _args(**kw), 1 << 20, 5)
        kept = sum(len(p.descs) for p in pieces)
        assert kept == ((10 if kw.get("identity") else 11) if kw["mode"] == "align" else 14)


def test_truncated_file_is_an_error(tmp_path):
    raw = open(DEMO, "rb").read()
    cut = str(tmp_path / "cut.bam")
    open(cut, "wb").write(raw[:len(raw) // 2])
    with pytest.raises(Exception):
        list(BamPieceReader(cut, _filter(_args()), threads=2, piece_bytes=1 << 20, align_to=1))


def test_header_only_and_zero_byte_files(tmp_path):
    path = str(tmp_path / "empty.bam")
    BamWriter(path, "@HD\tVN:1.6\n", [("chr1", 1000)]).close()
    rd = BamPieceReader(path, _filter(_args()), threads=2)
    assert rd.references == [("chr1", 1000)] and list(rd) == []
    rd.close()
    zero = str(tmp_path / "zero.bam")
    open(zero, "wb").close()
    with pytest.raises(ValueError):
        BamPieceReader(zero, _filter(_args()))


def test_writer_thread_indexes_while_it_writes(tmp_path):
    """call_mods' writer thread + StreamIndexer + _sort_and_index on the demo pieces (no GPU: the site lists come from the
    golden): the .bai written without reading the output back == the one index_sorted builds from the file."""
    import queue
    import threading
    from ccsmeth_b200 import bamsort
    g = load_npz("demo_callmods.npz")
    rd = BamPieceReader(DEMO, _filter(_args()), threads=2, piece_bytes=1 << 20, align_to=50)
    out = str(tmp_path / "o.modbam.bam")
    wr = BamWriter(out, rd.header_text, rd.references, threads=2)
    ix = bamsort.StreamIndexer(wr.header_bytes)
    q, counts, err = queue.Queue(), [0, 0, 0, 0], []
    th = threading.Thread(target=cm._writer_thread, args=(wr, q, False, True, counts, err, ix))
    th.start()
    off = k = 0
    for p in rd:
        c = g["n_sites_per_read"][k:k + len(p.recs)]
        sb = np.concatenate(([0], np.cumsum(c))).astype(np.int64)
        ns = int(sb[-1])
        q.put((p, p.recs, sb, g["mm"][off:off + ns].astype(np.int32), g["ml"][off:off + ns]))
        off += ns
        k += len(p.recs)
    q.put(None)
    th.join()
    assert not err and counts[2] == k
    wr.close()
    a = _args(output=str(tmp_path / "o"), no_sort=False)
    assert cm._sort_and_index(a, out, 0, 1, 2, ix, len(rd.references)) == out
    streamed = open(out + ".bai", "rb").read()
    os.remove(out + ".bai")
    assert bamsort.index_sorted(out, threads=2) == k
    assert open(out + ".bai", "rb").read() == streamed
    assert [r.query_name for r in BamReader(out)] == list(g["names"])
