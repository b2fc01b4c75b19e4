"""Builds raw BAM alignment records (SAM/BAM spec 4.2) for synthetic reads, so the same bytes feed the device
extractor (ccsmeth_b200.extract_features.pack_reads) and the numpy oracle (oracle.extract_numpy.extract_read)."""
import struct

import numpy as np

from ccsmeth_b200.bamio import BamRecord

_NIB = {c: i for i, c in enumerate("=ACMGRSVTWYHKDBN")}


def make_record(name, seq, fi, ri, fp, rp, fn=5, rn=6, flag=4, cigar=(), sn=None, extra_tags=b"", mapq=255):
    """seq: the STORED query sequence (for flag 0x10 that is the reverse complement of the forward read).
    fi/ri/fp/rp: uint8 arrays or None (tag left out)."""
    l_seq = len(seq)
    nm = name.encode("ascii") + b"\x00"
    packed = bytearray((l_seq + 1) // 2)
    for i, c in enumerate(seq):
        packed[i >> 1] |= _NIB[c] << (4 if i % 2 == 0 else 0)
    cig = b"".join(struct.pack("<I", (ln << 4) | op) for op, ln in cigar)
    core = struct.pack("<iiBBHHHiiii", -1 if flag & 4 else 0, -1 if flag & 4 else 100, len(nm), mapq, 4680, len(cigar),
                       flag, l_seq, -1, -1, 0)
    aux = b""
    if fn is not None:
        aux += b"fnC" + struct.pack("<B", fn) + b"rnC" + struct.pack("<B", rn)
    for tag, arr in (("fi", fi), ("fp", fp), ("ri", ri), ("rp", rp)):
        if arr is not None:
            a = np.asarray(arr, dtype=np.uint8)
            aux += tag.encode() + b"BC" + struct.pack("<I", len(a)) + a.tobytes()
    if sn is not None:
        aux += b"snBf" + struct.pack("<I", 4) + np.asarray(sn, dtype="<f4").tobytes()
    aux += b"zmi" + struct.pack("<i", 1234) + extra_tags
    raw = core + nm + cig + bytes(packed) + b"\xff" * l_seq + aux
    return BamRecord(raw)


def random_read(rng, name, n, p_cg=0.08, p_n=0.002, reverse=False, const_sig=None, no_cg=False, **kw):
    """A random forward read of n bases with CpGs sprinkled in; returns the BamRecord (stored orientation
    follows `reverse`) and the forward sequence string."""
    bases = np.array(list("ACGT"))
    s = bases[rng.integers(0, 4, n)]
    for i in np.nonzero(rng.random(max(n - 1, 0)) < p_cg)[0]:
        s[i], s[i + 1] = "C", "G"
    s[rng.random(n) < p_n] = "N"
    if no_cg:
        for i in range(n - 1):
            if s[i] == "C" and s[i + 1] == "G":
                s[i + 1] = "A"
    fwd = "".join(s)
    comp = {"A": "T", "C": "G", "G": "C", "T": "A", "N": "N"}
    stored = "".join(comp[c] for c in reversed(fwd)) if reverse else fwd
    sig = [rng.integers(0, 256, n).astype(np.uint8) for _ in range(4)]
    if const_sig is not None:
        sig[const_sig] = np.full(n, 37, dtype=np.uint8)
    flag = kw.pop("flag", (16 if reverse else 4))
    rec = make_record(name, stored, sig[0], sig[1], sig[2], sig[3], flag=flag, **kw)
    return rec, fwd
