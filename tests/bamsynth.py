"""Builds raw BAM alignment records (SAM/BAM spec 4.2) for synthetic reads, so the same bytes feed the device
extractor (ccsmeth_b200.extract_features.pack_reads) and the numpy oracle (oracle.extract_numpy.extract_read)."""
import struct

import numpy as np

from ccsmeth_b200.bamio import BamRecord

_NIB = {c: i for i, c in enumerate("=ACMGRSVTWYHKDBN")}


def make_record(name, seq, fi, ri, fp, rp, fn=5, rn=6, flag=4, cigar=(), sn=None, extra_tags=b"", mapq=255, ref_id=None,
                pos=None):
    """seq: the STORED query sequence (for flag 0x10 that is the reverse complement of the forward read).
    fi/ri/fp/rp: uint8 arrays or None (tag left out)."""
    l_seq = len(seq)
    nm = name.encode("ascii") + b"\x00"
    packed = bytearray((l_seq + 1) // 2)
    for i, c in enumerate(seq):
        packed[i >> 1] |= _NIB[c] << (4 if i % 2 == 0 else 0)
    cig = b"".join(struct.pack("<I", (ln << 4) | op) for op, ln in cigar)
    if ref_id is None:
        ref_id = -1 if flag & 4 else 0
    if pos is None:
        pos = -1 if flag & 4 else 100
    core = struct.pack("<iiBBHHHiiii", ref_id, pos, len(nm), mapq, 4680, len(cigar), flag, l_seq, -1, -1, 0)
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


def make_aligned_modbam(bam_path, fasta_path, seed=7, n_reads=320, chunk_len=10000):
    """A small synthetic reference (two contigs) and a position-sorted modbam aligned to it: both strands, soft clips,
    insertions / deletions / mismatches, MM/ML ("C+m?") on most CpG C's and a few other C's, HP tags on 60 % of the
    reads, plus secondary / supplementary / low-mapq / unmapped records for the filters.  A CG straddles the first
    chunk boundary so that the reference's region adjustment is exercised."""
    from ccsmeth_b200.bamio import BamWriter
    rng = np.random.default_rng(seed)
    comp = {"A": "T", "C": "G", "G": "C", "T": "A", "N": "N"}
    contigs = []
    for name, n in (("chrA", 3 * chunk_len + 1234), ("chrB", chunk_len + 2100)):
        b = np.array(list("ACGT"))[rng.integers(0, 4, n)]
        for i in np.nonzero(rng.random(n - 1) < 0.06)[0]:
            b[i], b[i + 1] = "C", "G"
        if name == "chrA":
            b[chunk_len - 1], b[chunk_len] = "C", "G"
        contigs.append((name, "".join(b)))
    with open(fasta_path, "w") as f:
        for name, seq in contigs:
            f.write(">%s synthetic\n" % name)
            for i in range(0, len(seq), 70):
                f.write(seq[i:i + 70].lower() if (i // 70) % 5 == 0 else seq[i:i + 70])
                f.write("\n")
    recs = []
    for k in range(n_reads):
        cid = int(rng.random() < 0.3)
        ref = contigs[cid][1]
        ln = int(rng.integers(600, 2500))
        start = int(rng.integers(0, len(ref) - ln))
        seg = list(ref[start:start + ln])
        cigar = []
        # build the aligned query with a few edits
        for i in np.nonzero(rng.random(ln) < 0.01)[0]:
            seg[i] = "ACGT"[int(rng.integers(0, 4))]
        q = "".join(seg)
        if rng.random() < 0.4:
            cut = int(rng.integers(100, ln - 100))
            ins = "".join("ACGT"[int(x)] for x in rng.integers(0, 4, int(rng.integers(1, 6))))
            dele = int(rng.integers(1, 6))
            q = q[:cut] + ins + q[cut + dele:]
            cut2 = cut  # M(cut) I(len ins) D(dele) M(rest)
            cigar = [(0, cut2), (1, len(ins)), (2, dele), (0, ln - cut - dele)]
        else:
            cigar = [(0, ln)]
        lc, rc = (int(rng.integers(0, 30)), int(rng.integers(0, 30))) if rng.random() < 0.5 else (0, 0)
        q = "".join("ACGT"[int(x)] for x in rng.integers(0, 4, lc)) + q + "".join("ACGT"[int(x)] for x in rng.integers(0, 4, rc))
        cigar = ([(4, lc)] if lc else []) + cigar + ([(4, rc)] if rc else [])
        reverse = bool(rng.random() < 0.5)
        fwd = "".join(comp[c] for c in reversed(q)) if reverse else q
        # MM / ML on the forward (original) read
        cs = [i for i, c in enumerate(fwd) if c == "C"]
        called = [j for j, i in enumerate(cs)
                  if (i + 1 < len(fwd) and fwd[i + 1] == "G" and rng.random() < 0.9) or rng.random() < 0.01]
        tags = b""
        if called and rng.random() < 0.97:
            deltas = [called[0]] + [called[j] - called[j - 1] - 1 for j in range(1, len(called))]
            mlv = np.floor(256 * rng.beta(0.3, 0.3, size=len(called))).clip(0, 255).astype(np.uint8)
            style = ("C+m?,", "C+m,", "C+m.,")[int(rng.integers(0, 3))]
            tags += b"MMZ" + (style + ",".join(map(str, deltas)) + ";").encode() + b"\x00"
            tags += b"MLBC" + struct.pack("<I", len(mlv)) + mlv.tobytes()
        u = rng.random()
        if u < 0.3:
            tags += b"HPi" + struct.pack("<i", 1)
        elif u < 0.6:
            tags += b"HPC" + struct.pack("<B", 2)
        flag = 16 if reverse else 0
        mapq = 60
        v = rng.random()
        if v < 0.03:
            flag |= 256
        elif v < 0.06:
            flag |= 2048
        elif v < 0.09:
            mapq = 0
        elif v < 0.11:
            flag |= 1024
        recs.append((cid, start, make_record("read%d" % k, q, None, None, None, None, fn=None, flag=flag, cigar=tuple(cigar),
                                             extra_tags=tags, mapq=mapq, ref_id=cid, pos=start)))
    recs.sort(key=lambda t: (t[0], t[1]))
    unm = make_record("unmapped", "ACGTACGTCGCG", None, None, None, None, fn=None, flag=4)
    wr = BamWriter(bam_path, "@HD\tVN:1.6\tSO:coordinate\n" + "".join("@SQ\tSN:%s\tLN:%d\n" % (n, len(s)) for n, s in contigs),
                   [(n, len(s)) for n, s in contigs])
    for _, _, r in recs:
        wr.write_raw(r.raw)
    wr.write_raw(unm.raw)
    wr.close()
    return contigs
