"""Coordinate sort + BAI index of the modbam -- what the reference does with ``pysam.sort`` / ``pysam.index`` after calling
(reference ccsmeth/call_modifications.py:592-607, skipped with ``--no_sort``), without samtools / pysam.

Order = samtools': (refID as unsigned, so unplaced reads sort last; pos; forward before reverse strand), stable for
equal keys.  Sorted runs are built in memory (``mem_bytes`` of inflated records at a time, keys from the native record
walk ``ccsm_bam_scan_records``); several runs -- more records than fit, or one shard per rank -- are merged by key.
The output is written in uniform 65,280-byte BGZF blocks, so a record's virtual file offset follows from its position
in the inflated stream and the compressed block sizes; the ``.bai`` (SAM spec section 5.2: binning index with the
metadata pseudo-bin, 16 kb linear index, unplaced-read count) is derived from those.
"""
import ctypes
import heapq
import mmap
import os
import struct

import numpy as np

from . import _lib
from .bamio import _BGZF_EOF

BLOCK = 65280
_PSEUDO_BIN = 37450


class _Inflated:
    """Inflated byte stream of a BGZF file, produced piece by piece by the library's thread team."""

    def __init__(self, path, threads, piece_bytes=64 << 20):
        self.lib = _lib.load()
        self.f = open(path, "rb")
        self.size = os.fstat(self.f.fileno()).st_size
        self.mm = mmap.mmap(self.f.fileno(), 0, access=mmap.ACCESS_READ) if self.size else None
        self.src = np.frombuffer(self.mm, dtype=np.uint8) if self.size else np.zeros(0, dtype=np.uint8)
        self.pos = 0
        self.utotal = 0       # inflated bytes handed out so far
        self.threads = max(1, threads)
        self.piece_bytes = piece_bytes

    def more(self, carry):
        """carry + the next inflated piece, or None at end of file."""
        if self.pos >= self.size:
            return None
        window = self.piece_bytes
        while True:
            n = min(window, self.size - self.pos)
            base = self.src.ctypes.data + self.pos
            consumed = ctypes.c_int64(0)
            total = self.lib.ccsm_bgzf_inflated_size(base, n, ctypes.byref(consumed))
            if total < 0:
                _lib.check(int(total))
            if consumed.value > 0:
                break
            if self.pos + n >= self.size:
                raise ValueError("truncated BGZF block at end of file")
            window *= 2
        buf = np.empty(len(carry) + int(total), dtype=np.uint8)
        buf[:len(carry)] = carry
        got = self.lib.ccsm_bgzf_inflate(base, consumed.value, buf[len(carry):].ctypes.data, int(total), self.threads,
                                         ctypes.byref(consumed))
        if got < 0:
            _lib.check(int(got))
        self.pos += consumed.value
        self.utotal += int(total)
        return buf

    def close(self):
        self.src = None
        if self.mm is not None:
            self.mm.close()
        self.f.close()


def _read_header(stream):
    """-> (header_text, references, raw header bytes, leftover inflated bytes positioned at the first record)."""
    buf = np.zeros(0, dtype=np.uint8)

    def need(n):
        nonlocal buf
        while len(buf) < n:
            nxt = stream.more(buf)
            if nxt is None:
                raise ValueError("truncated BAM header")
            buf = nxt

    need(12)
    if bytes(buf[:4]) != b"BAM\x01":
        raise ValueError("not a BAM file")
    l_text = struct.unpack("<i", bytes(buf[4:8]))[0]
    need(12 + l_text)
    text = bytes(buf[8:8 + l_text]).rstrip(b"\x00").decode("utf-8", "replace")
    p = 8 + l_text
    n_ref = struct.unpack("<i", bytes(buf[p:p + 4]))[0]
    p += 4
    refs = []
    for _ in range(n_ref):
        need(p + 4)
        l_name = struct.unpack("<i", bytes(buf[p:p + 4]))[0]
        need(p + 8 + l_name)
        refs.append((bytes(buf[p + 4:p + 4 + l_name - 1]).decode("ascii"),
                     struct.unpack("<i", bytes(buf[p + 4 + l_name:p + 8 + l_name]))[0]))
        p += 8 + l_name
    return text, refs, buf[p:].copy()


def _scan(lib, buf):
    """Native walk over the complete records in buf -> dict of arrays, bytes consumed."""
    cap = max(16, len(buf) // 36 + 1)
    a = {"key": np.empty(cap, np.uint64), "off": np.empty(cap, np.int64), "len": np.empty(cap, np.int32),
         "ref": np.empty(cap, np.int32), "pos": np.empty(cap, np.int32), "end": np.empty(cap, np.int32),
         "flag": np.empty(cap, np.int32)}
    consumed = ctypes.c_int64(0)
    n = lib.ccsm_bam_scan_records(buf.ctypes.data, len(buf), cap, a["key"].ctypes.data, a["off"].ctypes.data,
                                  a["len"].ctypes.data, a["ref"].ctypes.data, a["pos"].ctypes.data,
                                  a["end"].ctypes.data, a["flag"].ctypes.data, ctypes.byref(consumed))
    if n < 0:
        _lib.check(int(n))
    return {k: v[:n] for k, v in a.items()}, consumed.value


def sorted_header(text):
    """@HD with SO:coordinate (added if the header has none), like `samtools sort`; a @PG line names this step."""
    lines = text.split("\n") if text else []
    lines = [l for l in lines if l != ""]
    hd = None
    for i, l in enumerate(lines):
        if l.startswith("@HD"):
            f = [x for x in l.split("\t") if not x.startswith("SO:") and not x.startswith("GO:")]
            hd = i
            lines[i] = "\t".join(f + ["SO:coordinate"])
            break
    if hd is None:
        lines.insert(0, "@HD\tVN:1.6\tSO:coordinate")
    pp = [l.split("\tID:")[1].split("\t")[0] for l in lines if l.startswith("@PG") and "\tID:" in l]
    lines.append("@PG\tID:ccsmeth_b200.sort\tPN:ccsmeth_b200%s\tCL:coordinate sort + bai (call_mods without --no_sort)"
                 % ("\tPP:" + pp[-1] if pp else ""))
    return "\n".join(lines) + "\n"


def _header_bytes(text, refs):
    t = text.encode("utf-8")
    out = [b"BAM\x01", struct.pack("<i", len(t)), t, struct.pack("<i", len(refs))]
    for name, l_ref in refs:
        nm = name.encode("ascii") + b"\x00"
        out += [struct.pack("<i", len(nm)), nm, struct.pack("<i", l_ref)]
    return b"".join(out)


class _BlockWriter:
    """Writes an inflated stream as uniform BLOCK-byte BGZF blocks and remembers every block's file offset."""

    def __init__(self, path, threads, level=6, strategy="rle"):
        self.lib = _lib.load()
        self.f = open(path, "wb")
        self.threads = max(1, threads)
        self.level = level | (_lib.BGZF_RLE if strategy == "rle" else 0)
        self.pending = []      # numpy uint8 chunks not yet deflated
        self.pending_n = 0
        self.upos = 0          # inflated bytes accepted so far
        self.cpos = 0          # compressed bytes written so far
        self.block_coff = []   # arrays of compressed offsets, one entry per block
        self.out = None

    def tell(self):
        return self.upos

    def write(self, arr):
        arr = np.frombuffer(arr, dtype=np.uint8) if not isinstance(arr, np.ndarray) else arr
        if len(arr) == 0:
            return
        self.pending.append(arr)
        self.pending_n += len(arr)
        self.upos += len(arr)
        if self.pending_n >= 256 * BLOCK:
            self._flush(False)

    def _flush(self, final):
        if self.pending_n == 0:
            return
        data = np.concatenate(self.pending) if len(self.pending) > 1 else self.pending[0]
        n = len(data) if final else len(data) // BLOCK * BLOCK
        if n:
            cap = int(self.lib.ccsm_bgzf_deflate_bound(n))
            if self.out is None or len(self.out) < cap:
                self.out = np.empty(cap, dtype=np.uint8)
            src = np.ascontiguousarray(data[:n])
            got = self.lib.ccsm_bgzf_deflate(src.ctypes.data, n, self.out.ctypes.data, len(self.out), self.level, self.threads)
            if got < 0:
                _lib.check(int(got))
            got = int(got)
            # block sizes from the BSIZE field of every block header (the library writes XLEN = 6: BSIZE at byte 16)
            nb = (n + BLOCK - 1) // BLOCK
            offs = np.empty(nb, dtype=np.int64)
            p = 0
            for i in range(nb):
                offs[i] = self.cpos + p
                p += (int(self.out[p + 16]) | (int(self.out[p + 17]) << 8)) + 1
            if p != got:
                raise RuntimeError("BGZF block walk ended at %d of %d bytes" % (p, got))
            self.block_coff.append(offs)
            self.f.write(memoryview(self.out)[:got])
            self.cpos += got
        rest = data[n:]
        self.pending = [rest.copy()] if len(rest) else []
        self.pending_n = len(rest)

    def close(self):
        self._flush(True)
        self.f.write(_BGZF_EOF)
        self.f.close()
        self.coff = np.concatenate(self.block_coff) if self.block_coff else np.zeros(0, dtype=np.int64)

    def voffset(self, upos):
        """Virtual file offsets of inflated stream positions (numpy int64 array); a position at the very end of the
        data maps to the EOF block."""
        upos = np.asarray(upos, dtype=np.int64)
        blk = upos // BLOCK
        coff = np.concatenate((self.coff, [self.cpos])).astype(np.uint64)
        inside = (upos - blk * BLOCK).astype(np.uint64)
        end = blk >= len(self.coff)
        return np.where(end, np.uint64(self.cpos) << np.uint64(16), (coff[np.minimum(blk, len(self.coff))] << np.uint64(16)) | inside)


def _reg2bin(beg, end):
    """SAM spec 5.3 (vectorised): smallest bin containing [beg, end)."""
    beg = np.asarray(beg, dtype=np.int64)
    end = np.asarray(end, dtype=np.int64) - 1
    out = np.zeros(len(beg), dtype=np.int64)
    done = np.zeros(len(beg), dtype=bool)
    for shift, base in ((14, 4681), (17, 585), (20, 73), (23, 9), (26, 1)):
        hit = ~done & ((beg >> shift) == (end >> shift))
        out[hit] = base + (beg[hit] >> shift)
        done |= hit
    return out


def write_bai(path, n_ref, ref, pos, end, flag, v_beg, v_end):
    """ref/pos/end/flag/v_beg/v_end: per record in file order (coordinate sorted)."""
    ref = np.asarray(ref)
    placed = ref >= 0
    n_no_coor = int((~placed).sum())
    with open(path, "wb") as f:
        f.write(b"BAI\x01" + struct.pack("<i", n_ref))
        order_ok = np.all(np.diff(ref[placed].astype(np.int64)) >= 0)
        if not order_ok:
            raise ValueError("records are not coordinate sorted")
        starts = np.searchsorted(ref[placed], np.arange(n_ref), side="left")
        stops = np.searchsorted(ref[placed], np.arange(n_ref), side="right")
        idx_placed = np.nonzero(placed)[0]
        for t in range(n_ref):
            sel = idx_placed[starts[t]:stops[t]]
            if len(sel) == 0:
                f.write(struct.pack("<i", 0) + struct.pack("<i", 0))
                continue
            p, e, fl = pos[sel].astype(np.int64), end[sel].astype(np.int64), flag[sel]
            p = np.maximum(p, 0)
            e = np.maximum(e, p + 1)
            vb, ve = v_beg[sel], v_end[sel]
            bins = _reg2bin(p, e)
            # a chunk = a run of consecutive records in the same bin (what `samtools index` emits)
            cut = np.concatenate(([True], bins[1:] != bins[:-1]))
            cs = np.nonzero(cut)[0]
            ce = np.concatenate((cs[1:], [len(bins)])) - 1
            cbin, cbeg, cend = bins[cs], vb[cs], ve[ce]
            order = np.argsort(cbin, kind="stable")
            cbin, cbeg, cend = cbin[order], cbeg[order], cend[order]
            ub, first = np.unique(cbin, return_index=True)
            last = np.concatenate((first[1:], [len(cbin)]))
            out = [struct.pack("<i", len(ub) + 1)]
            for b, a0, a1 in zip(ub, first, last):
                out.append(struct.pack("<Ii", int(b), int(a1 - a0)))
                ch = np.empty(2 * (a1 - a0), dtype=np.uint64)
                ch[0::2] = cbeg[a0:a1]
                ch[1::2] = cend[a0:a1]
                out.append(ch.tobytes())
            n_unmapped = int(((fl & 4) != 0).sum())
            out.append(struct.pack("<IiQQQQ", _PSEUDO_BIN, 2, int(vb[0]), int(ve[-1]), len(sel) - n_unmapped, n_unmapped))
            # linear index: smallest virtual offset of a record overlapping each 16 kb window
            n_intv = int(((e - 1) >> 14).max()) + 1
            lin = np.full(n_intv, np.iinfo(np.uint64).max, dtype=np.uint64)
            w0, w1 = p >> 14, (e - 1) >> 14
            span = (w1 - w0 + 1).astype(np.int64)
            rec = np.repeat(np.arange(len(sel)), span)
            win = np.repeat(w0, span) + (np.arange(span.sum()) - np.repeat(np.cumsum(span) - span, span))
            np.minimum.at(lin, win, vb[rec])
            for i in range(n_intv - 2, -1, -1):       # empty windows take the next one's offset (htslib)
                if lin[i] == np.iinfo(np.uint64).max:
                    lin[i] = lin[i + 1]
            out.append(struct.pack("<i", n_intv) + lin.tobytes())
            f.write(b"".join(out))
        f.write(struct.pack("<Q", n_no_coor))


def _block_table(path):
    """(compressed offset, inflated size) of every BGZF block of a file, from the block headers alone."""
    coff, isz = [], []
    with open(path, "rb") as f:
        mm = mmap.mmap(f.fileno(), 0, access=mmap.ACCESS_READ)
        p, n = 0, len(mm)
        while p + 18 <= n:
            xlen = mm[p + 10] | (mm[p + 11] << 8)
            bsize, i = None, 0
            while i + 4 <= xlen:
                e = p + 12 + i
                slen = mm[e + 2] | (mm[e + 3] << 8)
                if mm[e] == 66 and mm[e + 1] == 67 and slen == 2:
                    bsize = (mm[e + 4] | (mm[e + 5] << 8)) + 1
                i += 4 + slen
            if bsize is None or p + bsize > n:
                raise ValueError("malformed BGZF block at byte %d of %s" % (p, path))
            coff.append(p)
            isz.append(struct.unpack_from("<I", mm, p + bsize - 4)[0])
            p += bsize
        mm.close()
    return np.array(coff, dtype=np.int64), np.array(isz, dtype=np.int64)


def _write_bai_for(path, n_refs, c):
    """.bai of a written BAM from its records' (ref, pos, end, flag, inflated start, inflated end): virtual offsets come
    from the file's BGZF block table."""
    coff, isz = _block_table(path)
    ustart_of_block = np.concatenate(([0], np.cumsum(isz)))
    data_blk = np.nonzero(isz > 0)[0]           # empty blocks (EOF markers) hold no positions
    dstart = ustart_of_block[data_blk]
    eof_coff = int(coff[-1]) if len(coff) and isz[-1] == 0 else int(os.path.getsize(path))

    def voff(u):
        u = np.asarray(u, dtype=np.int64)
        j = np.searchsorted(dstart, u, side="right") - 1
        j = np.clip(j, 0, max(len(data_blk) - 1, 0))
        inside = u - dstart[j]
        at_end = inside >= isz[data_blk[j]]          # the end of the last record: the EOF block
        return np.where(at_end, np.uint64(eof_coff) << np.uint64(16),
                        (coff[data_blk[j]].astype(np.uint64) << np.uint64(16)) | inside.astype(np.uint64))

    write_bai(path + ".bai", n_refs, c["ref"], c["pos"], c["end"], c["flag"], voff(c["ustart"]), voff(c["uend"]))
    return len(c["ref"])


class StreamIndexer:
    """index_sorted without reading the file back: the writer feeds the record bytes it is about to write (whole records,
    in file order), and after the file is closed `finish` writes the .bai -- or reports that the records were not in
    coordinate order, in which case the caller sorts (call_mods: the output of a sorted or unaligned input is already in
    order, and re-inflating what was just deflated was a third of the demo run)."""

    def __init__(self, header_bytes):
        self.lib = _lib.load()
        self.u = int(header_bytes)   # inflated offset of the next record
        self.metas = []
        self.last_key = -1
        self.in_order = True

    def feed(self, data):
        data = np.ascontiguousarray(np.frombuffer(data, dtype=np.uint8) if not isinstance(data, np.ndarray) else data)
        if not len(data):
            return
        if self.in_order:
            a, used = _scan(self.lib, data)
            if used != len(data):
                raise ValueError("StreamIndexer.feed: the chunk does not end on a record boundary")
            k = a["key"]
            if len(k):
                if int(k[0]) < self.last_key or np.any(k[1:] < k[:-1]):
                    self.in_order = False
                    self.metas = []
                else:
                    self.last_key = int(k[-1])
                    a["ustart"] = self.u + a["off"]
                    a["uend"] = a["ustart"] + a["len"]
                    self.metas.append({f: a[f] for f in ("ref", "pos", "end", "flag", "ustart", "uend")})
        self.u += len(data)

    def finish(self, path, n_refs):
        """Record count, or -1 (nothing written) when the records were not sorted."""
        if not self.in_order:
            return -1
        c = {k: (np.concatenate([m[k] for m in self.metas]) if self.metas else np.zeros(0, np.int64))
             for k in ("ref", "pos", "end", "flag", "ustart", "uend")}
        return _write_bai_for(path, n_refs, c)


def index_sorted(path, threads=4):
    """Writes path + ".bai" for a BAM whose records are already in coordinate order (no rewrite).  Returns the record
    count, or -1 -- and writes nothing -- if the records turn out not to be sorted."""
    lib = _lib.load()
    st = _Inflated(path, threads)
    _text, refs, carry = _read_header(st)
    metas = []
    buf = carry
    last_key = -1
    while True:
        a, used = _scan(lib, buf)
        base = st.utotal - len(buf)   # inflated position of buf[0]
        if len(a["key"]):
            k = a["key"]
            if int(k[0]) < last_key or np.any(k[1:] < k[:-1]):
                st.close()
                return -1
            last_key = int(k[-1])
            a["ustart"] = base + a["off"]
            a["uend"] = a["ustart"] + a["len"]
            metas.append(a)
        nxt = st.more(buf[used:].copy())
        if nxt is None:
            break
        buf = nxt
    st.close()
    c = {k: (np.concatenate([m[k] for m in metas]) if metas else np.zeros(0, np.int64))
         for k in ("ref", "pos", "end", "flag", "ustart", "uend")}
    return _write_bai_for(path, len(refs), c)


class _Run:
    """A sorted run on disk: a headerless BGZF stream of records + its sorted keys."""

    def __init__(self, path, keys):
        self.path, self.keys = path, keys


def _iter_run(run, threads):
    """Yields (key, record bytes as a uint8 array) of a run file in order."""
    st = _Inflated(run.path, threads, piece_bytes=16 << 20)
    lib = st.lib
    carry = np.zeros(0, dtype=np.uint8)
    k = 0
    while True:
        buf = st.more(carry)
        if buf is None:
            break
        a, used = _scan(lib, buf)
        for off, ln in zip(a["off"].tolist(), a["len"].tolist()):
            yield int(run.keys[k]), buf[off:off + ln]
            k += 1
        carry = buf[used:].copy()
    st.close()


def sort_and_index(paths, out_path, threads=4, mem_bytes=None, write_index=True, bam_compress="rle", tmp_dir=None):
    """Coordinate-sorts the records of the BAM file(s) `paths` (same header; e.g. one shard per rank) into `out_path`
    and writes `out_path + ".bai"`.  Returns the number of records."""
    lib = _lib.load()
    if isinstance(paths, str):
        paths = [paths]
    mem_bytes = int(mem_bytes or os.environ.get("CCSM_SORT_MEM", 4 << 30))
    tmp_dir = tmp_dir or os.path.dirname(os.path.abspath(out_path))
    header = None
    runs, mem_pieces, mem_meta, mem_n = [], [], [], 0
    tmp_files = []

    def flush_run(final):
        nonlocal mem_pieces, mem_meta, mem_n
        if not mem_pieces:
            return None
        key = np.concatenate([m["key"] for m in mem_meta])
        piece = np.concatenate([np.full(len(m["key"]), i, dtype=np.int32) for i, m in enumerate(mem_meta)])
        cat = {k: np.concatenate([m[k] for m in mem_meta]) for k in ("off", "len", "ref", "pos", "end", "flag")}
        order = np.argsort(key, kind="stable")
        res = (order, key, piece, cat, mem_pieces)
        mem_pieces, mem_meta, mem_n = [], [], 0
        return res

    def emit(res, writer, collect):
        order, key, piece, cat, pieces = res
        ln = cat["len"][order].astype(np.int64)
        ustart = writer.tell() + np.concatenate(([0], np.cumsum(ln)[:-1]))
        # copy the records out piece-group by piece-group to keep it vectorised: gather byte ranges
        for i in order.tolist():
            pc = pieces[piece[i]]
            o = int(cat["off"][i])
            writer.write(pc[o:o + int(cat["len"][i])])
        if collect is not None:
            collect.append({"ref": cat["ref"][order], "pos": cat["pos"][order], "end": cat["end"][order],
                            "flag": cat["flag"][order], "ustart": ustart, "uend": ustart + ln})
        return key[order]

    # ---- pass 1: sorted runs
    single_pass = None
    for path in paths:
        st = _Inflated(path, threads)
        text, refs, carry = _read_header(st)
        if header is None:
            header = (text, refs)
        elif refs != header[1]:
            raise ValueError("shards have different reference dictionaries")
        buf = carry
        while True:
            a, used = _scan(lib, buf)
            if len(a["key"]):
                mem_pieces.append(buf)
                mem_meta.append(a)
                mem_n += used
            rest = buf[used:].copy()
            if mem_n >= mem_bytes:
                res = flush_run(False)
                rp = os.path.join(tmp_dir, ".ccsm_sort_run%d_%d.tmp" % (os.getpid(), len(runs)))
                w = _BlockWriter(rp, threads, strategy="rle")
                keys = emit(res, w, None)
                w.close()
                runs.append(_Run(rp, keys))
                tmp_files.append(rp)
            buf = st.more(rest)
            if buf is None:
                if len(rest):
                    raise ValueError("truncated BAM record at end of %s" % path)
                break
        st.close()
    text, refs = header
    writer = _BlockWriter(out_path + ".sorting.tmp", threads, strategy=bam_compress)
    writer.write(np.frombuffer(_header_bytes(sorted_header(text), refs), dtype=np.uint8))
    collect = []
    n_rec = 0
    if not runs:
        res = flush_run(True)
        if res is not None:
            n_rec = len(emit(res, writer, collect))
    else:
        res = flush_run(True)
        if res is not None:   # the tail becomes one more run so that the merge sees uniform inputs
            rp = os.path.join(tmp_dir, ".ccsm_sort_run%d_%d.tmp" % (os.getpid(), len(runs)))
            w = _BlockWriter(rp, threads, strategy="rle")
            keys = emit(res, w, None)
            w.close()
            runs.append(_Run(rp, keys))
            tmp_files.append(rp)
        # ---- pass 2: k-way merge by (key, run index): stable, since run r holds earlier input than run r + 1
        its = [((k, r, rec) for k, rec in _iter_run(run, max(1, threads // 2))) for r, run in enumerate(runs)]
        meta = {k: [] for k in ("ref", "pos", "end", "flag", "ustart", "uend")}
        for _k, _r, rec in heapq.merge(*its, key=lambda t: (t[0], t[1])):
            u0 = writer.tell()
            writer.write(rec)
            ref_id, pos = struct.unpack_from("<ii", rec, 4)
            meta["ustart"].append(u0)
            meta["uend"].append(u0 + len(rec))
            meta["ref"].append(ref_id)
            meta["pos"].append(pos)
            n_rec += 1
        # end / flag for the index: one more native walk over what was merged is cheaper than parsing CIGARs here
        collect = None
        merged_meta = meta
    writer.close()
    os.replace(out_path + ".sorting.tmp", out_path)
    for p in tmp_files:
        try:
            os.remove(p)
        except OSError:
            pass
    if write_index:
        if collect is None:
            # merged output: walk the written file once for end / flag (ref, pos, offsets are known from the merge)
            st = _Inflated(out_path, threads)
            _t, _r, carry = _read_header(st)
            ends, flags = [], []
            buf = carry
            while True:
                a, used = _scan(lib, buf)
                ends.append(a["end"])
                flags.append(a["flag"])
                buf = st.more(buf[used:].copy())
                if buf is None:
                    break
            st.close()
            m = merged_meta
            c = {"ref": np.array(m["ref"], np.int32), "pos": np.array(m["pos"], np.int32),
                 "end": np.concatenate(ends) if ends else np.zeros(0, np.int32),
                 "flag": np.concatenate(flags) if flags else np.zeros(0, np.int32),
                 "ustart": np.array(m["ustart"], np.int64), "uend": np.array(m["uend"], np.int64)}
        elif collect:
            c = {k: np.concatenate([x[k] for x in collect]) for k in collect[0]}
        else:
            c = {k: np.zeros(0, np.int64) for k in ("ref", "pos", "end", "flag", "ustart", "uend")}
        write_bai(out_path + ".bai", len(refs), c["ref"], c["pos"], c["end"], c["flag"],
                  writer.voffset(c["ustart"]), writer.voffset(c["uend"]))
    return n_rec
