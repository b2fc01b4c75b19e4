"""Coordinate sort + BAI of the modbam (reference call_modifications.py:592-607 runs samtools sort / index through pysam).
The index is checked the way a reader uses it: region queries through bins + linear index + virtual offsets must return
exactly the records a brute-force scan finds."""
import os
import struct
import zlib

import numpy as np
import pytest

from ccsmeth_b200 import bamsort
from ccsmeth_b200.bamio import BamReader, BamWriter
from tests.bamsynth import make_record


def _ref_len(cigar):
    return sum(ln for op, ln in cigar if op in (0, 2, 3, 7, 8))


def _make_unsorted(path, rng, n=1500, with_hd=True):
    refs = [("chrA", 400000), ("chrB", 90000), ("chrC", 1000)]
    recs = []
    for k in range(n):
        u = rng.random()
        if u < 0.05:
            r = make_record("u%d" % k, "ACGT" * 5, None, None, None, None, fn=None, flag=4)
            recs.append((r, -1, -1, 1, 4))
            continue
        tid = int(rng.choice([0, 0, 0, 1, 1, 2]))
        ln = int(rng.integers(30, 900)) if tid == 2 else int(rng.integers(200, 40000))
        ln = min(ln, refs[tid][1] - 1)
        pos = int(rng.integers(0, refs[tid][1] - ln))
        # duplicates of (tid, pos) exercise the strand tie-break and stability
        if k % 7 == 0 and recs and recs[-1][1] == tid:
            pos = recs[-1][2]
        cigar = ((4, 3), (0, ln // 2), (2, 5), (1, 4), (0, ln - ln // 2 - 5)) if ln > 20 else ((0, ln),)
        flag = (16 if rng.random() < 0.5 else 0) | (4 if rng.random() < 0.02 else 0)
        qlen = sum(l for op, l in cigar if op in (0, 1, 4))
        r = make_record("r%d" % k, "A" * min(qlen, 50), None, None, None, None, fn=None, flag=flag, cigar=cigar, ref_id=tid,
                        pos=pos, mapq=60)
        recs.append((r, tid, pos, 1 if flag & 4 else _ref_len(cigar), flag))
    hdr = ("@HD\tVN:1.6\tSO:unsorted\n" if with_hd else "") + "".join("@SQ\tSN:%s\tLN:%d\n" % x for x in refs) + \
        "@PG\tPN:ccsmeth\tID:ccsmeth\tVN:0.5.0\tCL:test\n"
    w = BamWriter(path, hdr, refs)
    for r in recs:
        w.write_raw(r[0].raw)
    w.close()
    return refs, recs


def _key(tid, pos, flag):
    return ((tid & 0xFFFFFFFF) << 32) | (((pos + 1) & 0xFFFFFFFF) << 1) | (1 if flag & 16 else 0)


class _Bgzf:
    """Random access by virtual offset, with plain zlib."""

    def __init__(self, path):
        self.d = open(path, "rb").read()
        self.cache = {}

    def block(self, coff):
        if coff not in self.cache:
            d = self.d
            xlen = struct.unpack_from("<H", d, coff + 10)[0]
            bsize = struct.unpack_from("<H", d, coff + 16)[0] + 1
            raw = zlib.decompress(d[coff + 12 + xlen:coff + bsize - 8], -15)
            self.cache[coff] = (raw, coff + bsize)
        return self.cache[coff]

    def read_records(self, vbeg, vend):
        """Records starting in [vbeg, vend)."""
        coff, uoff = vbeg >> 16, vbeg & 0xFFFF
        out = []
        while (coff << 16 | uoff) < vend:
            need = 4
            buf = b""
            c, u = coff, uoff
            while len(buf) < need:
                raw, nxt = self.block(c)
                take = raw[u:u + need - len(buf)]
                buf += take
                u += len(take)
                if u >= len(raw):
                    c, u = nxt, 0
                if len(buf) == 4 and need == 4:
                    need = 4 + struct.unpack("<i", buf)[0]
            out.append(buf[4:])
            coff, uoff = c, u
        return out


def _reg2bins(beg, end):
    end -= 1
    bins = [0]
    for shift, base in ((26, 1), (23, 9), (20, 73), (17, 585), (14, 4681)):
        bins += list(range(base + (beg >> shift), base + (end >> shift) + 1))
    return bins


def _load_bai(path):
    d = open(path, "rb").read()
    assert d[:4] == b"BAI\x01"
    n_ref = struct.unpack_from("<i", d, 4)[0]
    p = 8
    refs = []
    for _ in range(n_ref):
        n_bin = struct.unpack_from("<i", d, p)[0]
        p += 4
        bins = {}
        for _ in range(n_bin):
            b, nc = struct.unpack_from("<Ii", d, p)
            p += 8
            ch = struct.unpack_from("<%dQ" % (2 * nc), d, p)
            p += 16 * nc
            bins[b] = list(zip(ch[0::2], ch[1::2]))
        n_intv = struct.unpack_from("<i", d, p)[0]
        p += 4
        lin = struct.unpack_from("<%dQ" % n_intv, d, p)
        p += 8 * n_intv
        refs.append((bins, lin))
    n_no_coor = struct.unpack_from("<Q", d, p)[0] if p + 8 <= len(d) else None
    return refs, n_no_coor


@pytest.mark.parametrize("mem", [None, 200000])
def test_sort_and_index(tmp_path, mem):
    rng = np.random.default_rng(11)
    src = str(tmp_path / "in.bam")
    refs, recs = _make_unsorted(src, rng)
    out = str(tmp_path / "out.bam")
    n = bamsort.sort_and_index(src, out, threads=3, mem_bytes=mem)
    assert n == len(recs)
    rd = BamReader(out)
    got = list(rd)
    assert rd.header_text.startswith("@HD\tVN:1.6\tSO:coordinate\n") and "ID:ccsmeth_b200.sort" in rd.header_text
    assert rd.references == refs
    # samtools' order, stable for equal keys
    want = sorted(range(len(recs)), key=lambda i: _key(recs[i][1], recs[i][2], recs[i][4]))
    assert [g.raw for g in got] == [recs[i][0].raw for i in want]
    # the index answers region queries exactly
    bai, n_no_coor = _load_bai(out + ".bai")
    assert len(bai) == len(refs) and n_no_coor == sum(1 for r in recs if r[1] < 0)
    bg = _Bgzf(out)
    for tid in range(len(refs)):
        bins, lin = bai[tid]
        meta = bins.pop(37450)
        placed = [r for r in recs if r[1] == tid]
        assert meta[1] == (sum(1 for r in placed if not r[4] & 4), sum(1 for r in placed if r[4] & 4))
        for _ in range(25):
            beg = int(rng.integers(0, refs[tid][1] - 1))
            end = min(refs[tid][1], beg + int(rng.integers(1, 60000)))
            brute = sorted(r[0].raw for r in placed if r[2] < end and r[2] + r[3] > beg)
            min_off = lin[beg >> 14] if (beg >> 14) < len(lin) else None
            found = []
            for b in _reg2bins(beg, end):
                for cb, ce in bins.get(b, []):
                    if min_off is not None and ce <= min_off:
                        continue
                    for raw in bg.read_records(cb, ce):
                        t, p = struct.unpack_from("<ii", raw, 0)
                        ncig, fl = struct.unpack_from("<HH", raw, 12)
                        lname = raw[8]
                        rl = sum(v >> 4 for v in struct.unpack_from("<%dI" % ncig, raw, 32 + lname) if (v & 15) in (0, 2, 3, 7, 8))
                        rl = 1 if (fl & 4 or rl == 0) else rl
                        if t == tid and p < end and p + rl > beg:
                            found.append(raw)
            assert sorted(found) == brute, (tid, beg, end)


def test_sort_merges_rank_shards_and_adds_hd(tmp_path):
    rng = np.random.default_rng(5)
    a, b = str(tmp_path / "a.bam"), str(tmp_path / "b.bam")
    refs, ra = _make_unsorted(a, rng, n=300, with_hd=False)
    _, rb = _make_unsorted(b, rng, n=200, with_hd=False)
    out = str(tmp_path / "m.bam")
    assert bamsort.sort_and_index([a, b], out, threads=2) == 500
    rd = BamReader(out)
    got = [r.raw for r in rd]
    assert rd.header_text.split("\n")[0] == "@HD\tVN:1.6\tSO:coordinate"
    allr = ra + rb
    want = sorted(range(len(allr)), key=lambda i: _key(allr[i][1], allr[i][2], allr[i][4]))
    assert got == [allr[i][0].raw for i in want]
    assert os.path.exists(out + ".bai")


def test_unaligned_bam_keeps_its_order(tmp_path):
    """--mode denovo output (every record unplaced): the sort is the identity, the index lists only n_no_coor."""
    src = str(tmp_path / "u.bam")
    w = BamWriter(src, "@HD\tVN:1.5\tSO:unknown\n", [])
    raws = [make_record("z%d" % (97 - k), "ACGT", None, None, None, None, fn=None, flag=4).raw for k in range(40)]
    for r in raws:
        w.write_raw(r)
    w.close()
    out = str(tmp_path / "us.bam")
    assert bamsort.sort_and_index(src, out, threads=1) == 40
    assert [r.raw for r in BamReader(out)] == raws
    bai, n_no_coor = _load_bai(out + ".bai")
    assert bai == [] and n_no_coor == 40


def test_index_only_for_an_already_sorted_bam(tmp_path):
    """call_mods keeps the input order, so a sorted input needs no rewrite: the index is built over the file as it is
    (irregular BGZF blocks) and must answer queries like the rewritten one."""
    rng = np.random.default_rng(3)
    src = str(tmp_path / "in.bam")
    refs, recs = _make_unsorted(src, rng, n=900)
    out = str(tmp_path / "sorted.bam")
    bamsort.sort_and_index(src, out, threads=2)
    # rewrite the sorted records with the record-level writer (short, irregular blocks) and index in place
    rd = BamReader(out)
    again = str(tmp_path / "again.bam")
    w = BamWriter(again, rd.header_text, rd.references, threads=1)
    raws = [r.raw for r in rd]
    for i, r in enumerate(raws):
        w.write_raw(r)
        if i % 37 == 0:
            w.bg._flush(final=True) if hasattr(w.bg, "_flush") else None
    w.close()
    assert bamsort.index_sorted(again, threads=2) == len(raws)
    assert bamsort.index_sorted(src, threads=2) == -1 and not os.path.exists(src + ".bai")
    bai, n_no_coor = _load_bai(again + ".bai")
    bg = _Bgzf(again)
    for tid in range(len(refs)):
        bins, lin = bai[tid]
        bins.pop(37450)
        placed = [r for r in recs if r[1] == tid]
        for _ in range(20):
            beg = int(rng.integers(0, refs[tid][1] - 1))
            end = min(refs[tid][1], beg + int(rng.integers(1, 60000)))
            brute = sorted(r[0].raw for r in placed if r[2] < end and r[2] + r[3] > beg)
            found = []
            for b in _reg2bins(beg, end):
                for cb, ce in bins.get(b, []):
                    for raw in bg.read_records(cb, ce):
                        t, p = struct.unpack_from("<ii", raw, 0)
                        ncig, fl = struct.unpack_from("<HH", raw, 12)
                        rl = sum(v >> 4 for v in struct.unpack_from("<%dI" % ncig, raw, 32 + raw[8]) if (v & 15) in (0, 2, 3, 7, 8))
                        rl = 1 if (fl & 4 or rl == 0) else rl
                        if t == tid and p < end and p + rl > beg:
                            found.append(raw)
            assert sorted(found) == brute


def test_stream_indexer_matches_index_sorted(tmp_path):
    """The index built from the bytes the writer emits (no read-back) == the index built by re-reading the file, for
    every way of cutting the record stream into writes; an out-of-order stream is reported, not indexed."""
    rng = np.random.default_rng(5)
    src = str(tmp_path / "in.bam")
    refs, recs = _make_unsorted(src, rng, n=700)
    out = str(tmp_path / "sorted.bam")
    bamsort.sort_and_index(src, out, threads=2)
    rd = BamReader(out)
    raws = [struct.pack("<i", len(r.raw)) + r.raw for r in rd]
    for cut in (1, 13, 250, 10 ** 9):
        path = str(tmp_path / ("w%d.bam" % min(cut, 9999)))
        w = BamWriter(path, rd.header_text, rd.references, threads=2)
        ix = bamsort.StreamIndexer(w.header_bytes)
        for i in range(0, len(raws), cut):
            chunk = np.frombuffer(b"".join(raws[i:i + cut]), dtype=np.uint8)
            w.bg.write(chunk)
            ix.feed(chunk)
        w.close()
        assert ix.finish(path, len(rd.references)) == len(raws)
        streamed = open(path + ".bai", "rb").read()
        os.remove(path + ".bai")
        assert bamsort.index_sorted(path, threads=2) == len(raws)
        assert open(path + ".bai", "rb").read() == streamed
    # the unsorted input: reported as such, nothing written
    rd2 = BamReader(src)
    w = BamWriter(str(tmp_path / "u.bam"), rd2.header_text, rd2.references, threads=1)
    ix = bamsort.StreamIndexer(w.header_bytes)
    for r in rd2:
        b = np.frombuffer(struct.pack("<i", len(r.raw)) + r.raw, dtype=np.uint8)
        w.bg.write(b)
        ix.feed(b)
    w.close()
    assert ix.finish(str(tmp_path / "u.bam"), len(rd2.references)) == -1
    assert not os.path.exists(str(tmp_path / "u.bam") + ".bai")
    # a write that stops inside a record cannot be indexed
    ix = bamsort.StreamIndexer(100)
    with pytest.raises(ValueError):
        ix.feed(np.frombuffer(raws[0][:-3], dtype=np.uint8))
