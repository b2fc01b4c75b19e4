"""Minimal BAM reader/writer (BGZF via zlib) for the call_mods path -- no pysam/htslib dependency.

The reference does all BAM I/O through pysam (reader ccsmeth/extract_features.py:129-177, writer
ccsmeth/call_modifications.py:410-462).  pysam is not part of this environment, and the path only
needs: iterate records of a (possibly unaligned) HiFi BAM, read the kinetics tags, and write the same
records back with MM/ML tags added.  Records are kept as raw bytes; only the fields the path uses are
decoded, and untouched aux tags are copied byte-for-byte on output (types preserved exactly).

SAM/BAM spec section 4 (BGZF: 4.1; alignment record layout: 4.2; aux types: 4.2.4).
"""
import struct
import zlib

import numpy as np

_BGZF_EOF = bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000")
_SEQ_DECODE = "=ACMGRSVTWYHKDBN"
_SEQ_LUT = np.array([ord(_SEQ_DECODE[i >> 4]) for i in range(256)], dtype=np.uint8), \
    np.array([ord(_SEQ_DECODE[i & 15]) for i in range(256)], dtype=np.uint8)
_AUX_FIXED = {ord('A'): 1, ord('c'): 1, ord('C'): 1, ord('s'): 2, ord('S'): 2, ord('i'): 4, ord('I'): 4, ord('f'): 4}
_AUX_FMT = {ord('c'): '<b', ord('C'): '<B', ord('s'): '<h', ord('S'): '<H', ord('i'): '<i', ord('I'): '<I', ord('f'): '<f'}
_ARR_DTYPE = {ord('c'): np.int8, ord('C'): np.uint8, ord('s'): np.int16, ord('S'): np.uint16, ord('i'): np.int32,
              ord('I'): np.uint32, ord('f'): np.float32}


class BgzfReader:
    """Sequential reader over concatenated BGZF blocks.  ``threads > 1``: the file is read in large pieces and all
    complete blocks of a piece are inflated at once by libccsm's thread team (include/ccsm.h ccsm_bgzf_inflate;
    Python's own zlib holds the GIL in this interpreter, so Python threads do not scale)."""

    PIECE = 16 << 20

    def __init__(self, path, threads=1):
        self.f = open(path, "rb")
        self.buf = b""
        self.pos = 0
        self.threads = threads
        self.tail = b""
        if threads > 1:
            from . import _lib
            self.lib = _lib.load()

    def _fill_python(self):
        hdr = self.f.read(18)
        if len(hdr) < 18:
            return None
        if hdr[:4] != b"\x1f\x8b\x08\x04":
            raise ValueError("not a BGZF block")
        xlen = struct.unpack_from("<H", hdr, 10)[0]
        extra = hdr[12:18] + self.f.read(xlen - 6)
        bsize = None
        i = 0
        while i + 4 <= len(extra):
            si1, si2, slen = extra[i], extra[i + 1], struct.unpack_from("<H", extra, i + 2)[0]
            if si1 == 66 and si2 == 67:
                bsize = struct.unpack_from("<H", extra, i + 4)[0]
            i += 4 + slen
        if bsize is None:
            raise ValueError("BGZF block without BC subfield")
        cdata = self.f.read(bsize - xlen - 19)
        self.f.read(8)  # crc32 + isize
        return zlib.decompress(cdata, -15) if cdata else b""

    def _fill_native(self):
        import ctypes
        from . import _lib
        piece = self.f.read(self.PIECE)
        if not piece and not self.tail:
            return None
        src = self.tail + piece
        consumed = ctypes.c_int64(0)
        total = self.lib.ccsm_bgzf_inflated_size(src, len(src), ctypes.byref(consumed))
        if total < 0:
            _lib.check(int(total))
        if consumed.value == 0:
            if not piece:
                raise ValueError("truncated BGZF block at end of file")
            self.tail = src
            return b""
        dst = ctypes.create_string_buffer(int(total)) if total else None
        got = self.lib.ccsm_bgzf_inflate(src, len(src), dst, int(total), self.threads, ctypes.byref(consumed))
        if got < 0:
            _lib.check(int(got))
        self.tail = src[consumed.value:]
        return dst.raw if total else b""

    def _fill(self):
        data = self._fill_native() if self.threads > 1 else self._fill_python()
        if data is None:
            return False
        self.buf = self.buf[self.pos:] + data
        self.pos = 0
        return True

    def read(self, n):
        while len(self.buf) - self.pos < n:
            if not self._fill():
                break
        out = self.buf[self.pos:self.pos + n]
        self.pos += len(out)
        return out

    def close(self):
        self.f.close()


def _deflate_block(data, level, strategy=zlib.Z_DEFAULT_STRATEGY):
    c = zlib.compressobj(level, zlib.DEFLATED, -15, 9, strategy)
    cdata = c.compress(data) + c.flush()
    return (b"\x1f\x8b\x08\x04\x00\x00\x00\x00\x00\xff\x06\x00BC\x02\x00" + struct.pack("<H", len(cdata) + 25) + cdata +
            struct.pack("<II", zlib.crc32(data) & 0xffffffff, len(data)))


class BgzfWriter:
    """BGZF writer; ``threads > 1``: blocks are deflated by libccsm's thread team (ccsm_bgzf_deflate).

    strategy "zlib" (default): zlib's default strategy at `level` (what htslib does; right for text such as bed files);
    "rle": run-length matching + dynamic Huffman -- HiFi records are packed bases, qualities and kinetics bytes in which
    LZ77 finds next to nothing, so files stay within 3 % of the size; the thread team runs the library's own encoder
    (csrc/deflate_rle.h, 6-8x zlib level 6), the single-threaded Python path zlib's Z_RLE (BamWriter's default)."""

    BLOCK = 65280

    def __init__(self, path, level=6, threads=1, strategy="zlib"):
        if strategy not in ("rle", "zlib"):
            raise ValueError("BGZF strategy must be 'rle' or 'zlib'")
        self.f = open(path, "wb")
        self.level = level
        self.strategy = strategy
        self.buf = bytearray()
        self.threads = threads
        self.batch = 1
        self.out = None
        if threads > 1:
            from . import _lib
            self.lib = _lib.load()
            self.batch = 16 * threads

    def write(self, data):
        mv = memoryview(data).cast("B")  # bytes, bytearray or a uint8 numpy array
        if self.threads > 1 and len(mv) >= self.batch * self.BLOCK:
            # a large piece goes to the thread team straight from the caller's buffer: whatever is pending becomes its
            # own (short) blocks first, the piece's tail below one block waits for the next write
            if self.buf:
                self._deflate_native(self.buf)
                self.buf = bytearray()
            n_full = len(mv) // self.BLOCK * self.BLOCK
            self._deflate_native(mv[:n_full])
            mv = mv[n_full:]
        self.buf += mv
        if len(self.buf) >= self.batch * self.BLOCK:
            self._flush(final=False)

    def _deflate_native(self, view):
        import numpy as np
        from . import _lib
        src = np.frombuffer(view, dtype=np.uint8)
        cap = int(self.lib.ccsm_bgzf_deflate_bound(len(src)))
        if self.out is None or len(self.out) < cap:
            self.out = np.empty(cap, dtype=np.uint8)
        level = self.level | (_lib.BGZF_RLE if self.strategy == "rle" else 0)
        got = self.lib.ccsm_bgzf_deflate(src.ctypes.data, len(src), self.out.ctypes.data, len(self.out), level,
                                         self.threads)
        if got < 0:
            _lib.check(int(got))
        self.f.write(memoryview(self.out)[:int(got)])

    def _flush(self, final):
        n_full = len(self.buf) // self.BLOCK
        end = len(self.buf) if final else n_full * self.BLOCK
        if self.threads > 1:
            self._deflate_native(memoryview(self.buf)[:end])
        else:
            strategy = zlib.Z_RLE if self.strategy == "rle" else zlib.Z_DEFAULT_STRATEGY
            view = bytes(self.buf[:end])
            for i in range(0, len(view), self.BLOCK):
                self.f.write(_deflate_block(view[i:i + self.BLOCK], self.level, strategy))
        self.buf = self.buf[end:]

    def close(self):
        if self.buf:
            self._flush(final=True)
        self.f.write(_BGZF_EOF)
        self.f.close()


class BamRecord:
    """One alignment record; `raw` excludes the leading block_size field."""
    __slots__ = ("raw", "ref_id", "pos", "l_read_name", "mapq", "n_cigar", "flag", "l_seq", "_aux_off", "_tags",
                 "reference_name")

    def __init__(self, raw):
        self.raw = raw
        (self.ref_id, self.pos, self.l_read_name, self.mapq, _bin, self.n_cigar, self.flag, self.l_seq,
         _nref, _npos, _tlen) = struct.unpack_from("<iiBBHHHiiii", raw, 0)
        self._aux_off = 32 + self.l_read_name + 4 * self.n_cigar + (self.l_seq + 1) // 2 + self.l_seq
        self._tags = None
        self.reference_name = None  # filled in by BamReader for mapped records

    # ---- the attributes the extractor reads (reference extract_features.py:88-126)
    @property
    def query_name(self):
        return self.raw[32:32 + self.l_read_name - 1].decode("ascii")

    @property
    def is_unmapped(self):
        return bool(self.flag & 0x4)

    @property
    def is_reverse(self):
        return bool(self.flag & 0x10)

    @property
    def is_secondary(self):
        return bool(self.flag & 0x100)

    @property
    def is_duplicate(self):
        return bool(self.flag & 0x400)

    @property
    def is_supplementary(self):
        return bool(self.flag & 0x800)

    @property
    def cigartuples(self):
        off = 32 + self.l_read_name
        ops = np.frombuffer(self.raw, dtype="<u4", count=self.n_cigar, offset=off)
        return [(int(v & 0xf), int(v >> 4)) for v in ops]

    @property
    def query_sequence(self):
        off = 32 + self.l_read_name + 4 * self.n_cigar
        packed = np.frombuffer(self.raw, dtype=np.uint8, count=(self.l_seq + 1) // 2, offset=off)
        out = np.empty(2 * len(packed), dtype=np.uint8)
        out[0::2] = _SEQ_LUT[0][packed]
        out[1::2] = _SEQ_LUT[1][packed]
        return out[:self.l_seq].tobytes().decode("ascii")

    def get_forward_sequence(self):
        s = self.query_sequence
        if self.is_reverse:
            from .utils.process_utils import complement_seq
            return complement_seq(s)
        return s

    @property
    def query_alignment_start(self):
        """Leading soft clips (hard clips are not part of the stored sequence), like pysam."""
        start = 0
        for op, ln in self.cigartuples:
            if op == 5:
                continue
            if op == 4:
                start += ln
            else:
                break
        return start

    @property
    def query_alignment_end(self):
        end = self.l_seq
        for op, ln in reversed(self.cigartuples):
            if op == 5:
                continue
            if op == 4:
                end -= ln
            else:
                break
        return end

    # ---- pysam-compatible names the reference's extractor touches in align mode
    @property
    def mapping_quality(self):
        return self.mapq

    @property
    def reference_start(self):
        return self.pos

    @property
    def reference_end(self):
        if self.is_unmapped or self.n_cigar == 0:
            return None
        return self.pos + sum(l for op, l in self.cigartuples if op in (0, 2, 3, 7, 8))

    def get_aligned_pairs(self, matches_only=False):
        """(query_pos, ref_pos) pairs like pysam: M/=/X pair up; with matches_only insertions, soft clips,
        deletions and skips produce no pair (otherwise the missing side is None)."""
        out = []
        q, r = 0, self.pos
        for op, ln in self.cigartuples:
            if op in (0, 7, 8):
                out.extend(zip(range(q, q + ln), range(r, r + ln)))
                q += ln
                r += ln
            elif op in (1, 4):
                if not matches_only:
                    out.extend((i, None) for i in range(q, q + ln))
                q += ln
            elif op in (2, 3):
                if not matches_only:
                    out.extend((None, i) for i in range(r, r + ln))
                r += ln
        return out

    @property
    def modified_bases(self):
        """pysam >= 0.19 parses MM/ML itself; None makes the reference fall back to its own tag parser
        (call_mods_freq_bam.py:172-197)."""
        return None

    def get_cigar_stats(self):
        base = [0] * 11
        blocks = [0] * 11
        for op, l in self.cigartuples:
            base[op] += l
            blocks[op] += 1
        try:
            base[10] = int(self.get_tag("NM"))
        except KeyError:
            pass
        return base, blocks

    # ---- aux tags
    def _parse_tags(self):
        tags = {}
        raw, i, n = self.raw, self._aux_off, len(self.raw)
        while i + 3 <= n:
            tag = raw[i:i + 2].decode("ascii")
            ty = raw[i + 2]
            start = i
            i += 3
            if ty in _AUX_FIXED:
                ln = _AUX_FIXED[ty]
                val = raw[i:i + 1].decode("ascii") if ty == ord('A') else struct.unpack_from(_AUX_FMT[ty], raw, i)[0]
                i += ln
            elif ty in (ord('Z'), ord('H')):
                j = raw.index(b"\x00", i)
                val = raw[i:j].decode("ascii")
                i = j + 1
            elif ty == ord('B'):
                sub = raw[i]
                cnt = struct.unpack_from("<I", raw, i + 1)[0]
                dt = np.dtype(_ARR_DTYPE[sub]).newbyteorder("<")
                val = np.frombuffer(raw, dtype=dt, count=cnt, offset=i + 5)
                i += 5 + cnt * dt.itemsize
            else:
                raise ValueError("bad aux type %r in read %s" % (chr(ty), self.query_name))
            tags[tag] = (val, start, i)
        self._tags = tags

    def get_tag(self, name):
        if self._tags is None:
            self._parse_tags()
        if name not in self._tags:
            raise KeyError(name)
        return self._tags[name][0]

    def has_tag(self, name):
        if self._tags is None:
            self._parse_tags()
        return name in self._tags

    def with_tags(self, drop, mm=None, ml=None):
        """Record bytes (without block_size) with tags in `drop` removed and MM:Z / ML:B:C appended
        (reference _bam2modbam.py:211-226 `_refill_tags`)."""
        if self._tags is None:
            self._parse_tags()
        out = bytearray(self.raw[:self._aux_off])
        for tag, (_, s, e) in self._tags.items():
            if tag in drop:
                continue
            out += self.raw[s:e]
        if mm is not None:
            out += b"MMZ" + mm.encode("ascii") + b"\x00"
            out += b"MLBC" + struct.pack("<I", len(ml)) + (ml if isinstance(ml, bytes) else bytes(bytearray(ml)))
        return bytes(out)


class BamReader:
    def __init__(self, path, threads=1):
        self.bg = BgzfReader(path, threads)
        if self.bg.read(4) != b"BAM\x01":
            raise ValueError("%s is not a BAM file" % path)
        l_text = struct.unpack("<i", self.bg.read(4))[0]
        self.header_text = self.bg.read(l_text).rstrip(b"\x00").decode("utf-8", "replace")
        n_ref = struct.unpack("<i", self.bg.read(4))[0]
        self.references = []
        for _ in range(n_ref):
            l_name = struct.unpack("<i", self.bg.read(4))[0]
            name = self.bg.read(l_name)[:-1].decode("ascii")
            l_ref = struct.unpack("<i", self.bg.read(4))[0]
            self.references.append((name, l_ref))

    def __iter__(self):
        while True:
            b = self.bg.read(4)
            if len(b) < 4:
                return
            n = struct.unpack("<i", b)[0]
            rec = BamRecord(self.bg.read(n))
            if 0 <= rec.ref_id < len(self.references):
                rec.reference_name = self.references[rec.ref_id][0]
            yield rec

    def close(self):
        self.bg.close()


class BamWriter:
    def __init__(self, path, header_text, references, level=6, threads=1, strategy="rle"):
        self.bg = BgzfWriter(path, level, threads, strategy)
        text = header_text.encode("utf-8")
        self.bg.write(b"BAM\x01" + struct.pack("<i", len(text)) + text + struct.pack("<i", len(references)))
        self.header_bytes = 12 + len(text)  # inflated size of everything before the first record
        for name, l_ref in references:
            nm = name.encode("ascii") + b"\x00"
            self.bg.write(struct.pack("<i", len(nm)) + nm + struct.pack("<i", l_ref))
            self.header_bytes += 8 + len(nm)

    def write_raw(self, raw):
        self.bg.write(struct.pack("<i", len(raw)) + raw)

    def close(self):
        self.bg.close()


def add_pg_line(header_text, version, cmdline):
    """Input header + @PG ID:ccsmeth (reference call_modifications.py:445)."""
    if header_text and not header_text.endswith("\n"):
        header_text += "\n"
    return header_text + "@PG\tPN:ccsmeth\tID:ccsmeth\tVN:%s\tCL:%s\n" % (version, cmdline)
