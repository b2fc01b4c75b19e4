"""Piece-wise BAM streaming on libccsm's host helpers (include/ccsm.h ccsm_bgzf_*, ccsm_bam_*).

The call_mods pipeline never materialises per-read Python objects: a *piece* is a run of complete alignment
records inflated into ONE contiguous buffer, plus two arrays produced by ``ccsm_bam_index`` -- one ``ccsm_bam_rec``
per record and one ``ccsm_read`` per read that takes part in calling.  The buffer itself is the blob the device
extractor reads (no repacking), and ``ccsm_bam_tag_records`` writes the records back with MM/ML straight from
the device's per-site outputs.  Stands in for the pysam reader / writer processes of the reference
(extract_features.py:129-177, call_modifications.py:410-462).
"""
import ctypes
import mmap
import os
import struct

import numpy as np

from . import _lib
from .extract_features import READ_DTYPE

REC_DTYPE = np.dtype([("off", "<i8"), ("len", "<i4"), ("aux_off", "<i4"), ("flag", "<i4"), ("mapq", "<i4"),
                      ("l_seq", "<i4"), ("n_cigar", "<i4"), ("read_idx", "<i4"), ("pad", "<i4")], align=True)
assert REC_DTYPE.itemsize == 40


class Piece:
    """buf: uint8 array of inflated records; recs: REC_DTYPE array; descs: READ_DTYPE array (offsets into buf);
    first: index of recs[0] in the file's record order; last: True for the final piece."""
    __slots__ = ("buf", "recs", "descs", "first", "last")

    def __init__(self, buf, recs, descs, first, last):
        self.buf, self.recs, self.descs, self.first, self.last = buf, recs, descs, first, last

    def names(self):
        """Read names (bytes) of all records -- only needed for hole-id filters."""
        out = []
        b = self.buf
        for r in self.recs:
            s = int(r["off"]) + 4
            out.append(bytes(b[s + 32:s + 32 + int(b[s + 8]) - 1]))
        return out


class BamPieceReader:
    """Iterates Pieces of about `piece_bytes` compressed bytes.  `align_to` (--holes_batch): every piece but the
    last holds a whole number of hole-batches, so the reference's per-hole-batch bookkeeping (h0 stream, batch
    counter, rank ownership) is unchanged by how the file is cut."""

    def __init__(self, path, bam_filter, threads=4, piece_bytes=48 << 20, align_to=1):
        self.lib = _lib.load()
        self.f = open(path, "rb")
        self.threads = max(1, threads)
        self.piece_bytes = max(int(piece_bytes), 1)
        self.align_to = max(1, align_to)
        self.filter = bam_filter
        # the compressed file is mapped, not read: the thread team inflates straight out of the page cache
        self.size = os.fstat(self.f.fileno()).st_size
        self.mm = mmap.mmap(self.f.fileno(), 0, access=mmap.ACCESS_READ) if self.size else None
        self.src = np.frombuffer(self.mm, dtype=np.uint8) if self.size else np.zeros(0, dtype=np.uint8)
        self.pos = 0          # next compressed byte
        self.carry = np.zeros(0, dtype=np.uint8)  # inflated bytes not yet consumed
        self.eof = False
        self.n_seen = 0
        self._read_header()

    # -- inflated byte stream
    def _inflate_more(self):
        """Appends the next inflated piece to self.carry; returns False at end of file."""
        if self.pos >= self.size:
            self.eof = True
            return False
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
            window *= 2  # not even one complete block in the window
        buf = np.empty(len(self.carry) + int(total), dtype=np.uint8)
        buf[:len(self.carry)] = self.carry
        got = self.lib.ccsm_bgzf_inflate(base, consumed.value, buf[len(self.carry):].ctypes.data, int(total), self.threads,
                                         ctypes.byref(consumed))
        if got < 0:
            _lib.check(int(got))
        self.pos += consumed.value
        self.carry = buf
        return True

    def _need(self, n):
        while len(self.carry) < n:
            if not self._inflate_more():
                raise ValueError("truncated BAM header")

    def _read_header(self):
        self._need(12)
        if bytes(self.carry[:4]) != b"BAM\x01":
            raise ValueError("not a BAM file")
        l_text = struct.unpack("<i", bytes(self.carry[4:8]))[0]
        self._need(12 + l_text)
        self.header_text = bytes(self.carry[8:8 + l_text]).rstrip(b"\x00").decode("utf-8", "replace")
        p = 8 + l_text
        n_ref = struct.unpack("<i", bytes(self.carry[p:p + 4]))[0]
        p += 4
        self.references = []
        for _ in range(n_ref):
            self._need(p + 4)
            l_name = struct.unpack("<i", bytes(self.carry[p:p + 4]))[0]
            self._need(p + 4 + l_name + 4)
            name = bytes(self.carry[p + 4:p + 4 + l_name - 1]).decode("ascii")
            l_ref = struct.unpack("<i", bytes(self.carry[p + 4 + l_name:p + 8 + l_name]))[0]
            self.references.append((name, l_ref))
            p += 8 + l_name
        self.carry = self.carry[p:].copy()

    def __iter__(self):
        lib = self.lib
        while True:
            more = self._inflate_more() if not self.eof else False
            buf = self.carry
            if len(buf) == 0 and not more:
                return
            cap = max(4096, len(buf) // 512)
            recs = np.zeros(cap, dtype=REC_DTYPE)
            descs = np.zeros(cap, dtype=READ_DTYPE)
            n_recs, n_descs, consumed = ctypes.c_int32(0), ctypes.c_int32(0), ctypes.c_int64(0)
            _lib.check(lib.ccsm_bam_index(buf.ctypes.data, len(buf), ctypes.byref(self.filter), recs.ctypes.data, cap,
                                          descs.ctypes.data, ctypes.byref(n_recs), ctypes.byref(n_descs),
                                          ctypes.byref(consumed)))
            nr = n_recs.value
            last = self.eof and consumed.value == len(buf)
            if not last and self.align_to > 1:
                # keep whole hole-batches only; the partial one waits for the next piece
                nr_keep = nr - (self.n_seen + nr) % self.align_to
                if nr_keep <= 0:
                    if self.eof:
                        nr_keep = nr
                    else:
                        continue  # need more data to complete even one hole-batch
                if nr_keep < nr:
                    nr = nr_keep
                    consumed.value = int(recs["off"][nr])
                    last = False
            if nr == 0:
                if self.eof:
                    if len(buf):
                        raise ValueError("truncated BAM record at end of file")
                    return
                continue
            recs = recs[:nr]
            nd = int(recs["read_idx"].max()) + 1 if nr else 0
            self.carry = buf[consumed.value:].copy()
            piece = Piece(buf, recs, descs[:max(nd, 0)], self.n_seen, last and len(self.carry) == 0)
            self.n_seen += nr
            yield piece
            if self.eof and len(self.carry) == 0:
                return

    def close(self):
        self.src = None  # drop the exported buffer before the mapping goes away
        if self.mm is not None:
            try:
                self.mm.close()
            except BufferError:  # a caller still holds a view; the mapping is released with it
                pass
            self.mm = None
        self.f.close()


def tag_records(piece, recs, keep_pulse, site_begin, mm, ml):
    """-> (bytes of the re-tagged records incl. block_size prefixes, number of records that got MM/ML)."""
    lib = _lib.load()
    n_sites = int(site_begin[-1]) if len(site_begin) else 0
    cap = int(recs["len"].sum()) + 36 * len(recs) + 13 * n_sites + 64
    out = np.empty(cap, dtype=np.uint8)
    with_mm = ctypes.c_int32(0)
    recs = np.ascontiguousarray(recs)
    sb = np.ascontiguousarray(site_begin, dtype=np.int64)
    got = lib.ccsm_bam_tag_records(piece.buf.ctypes.data, recs.ctypes.data, len(recs), 1 if keep_pulse else 0,
                                   sb.ctypes.data, mm.ctypes.data if mm is not None else None,
                                   ml.ctypes.data if ml is not None else None, out.ctypes.data, cap,
                                   ctypes.byref(with_mm))
    if got < 0:
        _lib.check(int(got))
    return out[:int(got)], with_mm.value
