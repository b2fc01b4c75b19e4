"""ORACLE (test infrastructure only): numpy restatement of the reference's per-read feature extraction.

The product extracts features on the device (ccsmeth_b200/csrc/extract.cu behind ccsm_reads_* in include/ccsm.h);
this file is the checker those kernels are compared with.  It is itself pinned against the reference's own
extractor run on the demo BAM (tests/golden/demo_callmods.npz, scripts/gen_golden.py gen_demo).

Mirrors ``extract_features_from_double_strand_read`` (reference ccsmeth/extract_features.py:261-406) for the
call_mods defaults (``--mode denovo|align`` site selection without ``--is_map``, ``--motifs CG``-style symmetric
motifs, ``--norm zscore|none|min-max|min-mean|mad``, CodecV1 decode unless ``--no_decode``): per read, decode the
``fi/ri/fp/rp`` kinetics (process_utils.py:426-449), normalise each over the whole read
(extract_features.py:181-199), scan the motif, and cut 21-base windows on the forward read and on its reverse
complement (kinetics of the reverse strand are NOT flipped, :316,319).

The reference builds one Python tuple per site; here a read's sites come out as stacked arrays
(``ReadFeatures``) that go straight into the batch loop, and ``to_feature_rows`` reproduces the reference's
22-field rows for parity tests.
"""
import numpy as np
from numpy.lib.stride_tricks import sliding_window_view

from ccsmeth_b200.utils.process_utils import base2code_dna, default_ref_loc, str2bool

# CodecV1: 8-bit code -> frame count (reference process_utils.py:426-449)
CODE2FRAMES = np.empty(256, dtype=np.int64)
CODE2FRAMES[0:64] = np.arange(0, 64)
CODE2FRAMES[64:128] = np.arange(64, 191, 2)
CODE2FRAMES[128:192] = np.arange(192, 445, 4)
CODE2FRAMES[192:256] = np.arange(448, 953, 8)

_CODE_LUT = np.full(256, 4, dtype=np.int64)
for _b, _c in base2code_dna.items():
    _CODE_LUT[ord(_b)] = _c
_COMP_LUT = np.full(256, ord('N'), dtype=np.uint8)
for _a, _b in zip("ACGTNWSMKRYBVDHZ", "TGCANWSKMYRVBHDZ"):
    _COMP_LUT[ord(_a)] = ord(_b)


class ReadFeatures:
    """All candidate sites of one read: locs (n,), kmer codes (n, L) int64, kinetics (n, L) float64, npass ints."""
    __slots__ = ("holeid", "locs", "fkmer", "rkmer", "fipd", "fpw", "ripd", "rpw", "npass_f", "npass_r", "sn",
                 "chrom", "chrom_pos", "strand", "fseq", "rseq")

    def __len__(self):
        return len(self.locs)


MAD_C = 0.6744897501960817  # scipy.stats.norm.ppf(3 / 4)


def statsmodels_mad(a, c=MAD_C):
    """statsmodels.robust.scale.mad of a 1-D array (statsmodels 0.14.0, the reference's environment.yml:13; the package
    is not in this image, so its published formula is restated): float64, ``median(|a - median(a)| / c)`` -- the division
    comes before the median."""
    a = np.asarray(a, dtype=np.double)
    if not a.size:
        return np.nan
    center = np.apply_over_axes(np.median, a, 0)
    return np.median(np.abs(a - center) / c, axis=0)


def _normalize_signals(signals, method="zscore"):
    """reference extract_features.py:181-199 (np.mean / population np.std / np.around 6)."""
    if method == "none":
        return np.around(signals, decimals=6)
    if method == "zscore":
        sshift, sscale = np.mean(signals), np.std(signals)
    elif method == "min-max":
        sshift, sscale = np.min(signals), np.max(signals) - np.min(signals)
    elif method == "min-mean":
        sshift, sscale = np.min(signals), np.mean(signals)
    elif method == "mad":
        sshift, sscale = np.median(signals), float(statsmodels_mad(signals))
    else:
        raise ValueError("--norm %s" % method)
    if sscale == 0.0:
        return np.zeros(len(signals), dtype=np.float64)
    return np.around((signals - sshift) / sscale, decimals=6)


def _motif_locs(seq_bytes, motifs, mod_loc):
    """All i with seq[i:i+len] in motifs, shifted by mod_loc (reference process_utils.py:121-138)."""
    mlen = len(motifs[0])
    n = len(seq_bytes)
    if n < mlen:
        return np.empty(0, dtype=np.int64)
    hit = np.zeros(n - mlen + 1, dtype=bool)
    for m in set(motifs):
        mb = np.frombuffer(m.encode("ascii"), dtype=np.uint8)
        h = np.ones(n - mlen + 1, dtype=bool)
        for k in range(mlen):
            h &= seq_bytes[k:n - mlen + 1 + k] == mb[k]
        hit |= h
    return np.nonzero(hit)[0].astype(np.int64) + mod_loc


def extract_read(read, motifs, args, holeids_e=None, holeids_ne=None):
    """One BAM record (ccsmeth_b200.bamio.BamRecord or any object with the attributes the reference reads,
    extract_features.py:88-126) -> ReadFeatures, or None if the read is skipped."""
    seq_name = read.query_name
    if holeids_e is not None and seq_name not in holeids_e:
        return None
    if holeids_ne is not None and seq_name in holeids_ne:
        return None
    if args.mode == "align":
        if read.is_unmapped or read.is_secondary or read.is_duplicate:
            return None
        if args.no_supplementary and read.is_supplementary:
            return None
        if read.mapq < args.mapq:
            return None
    try:
        tag_fi, tag_ri = read.get_tag("fi"), read.get_tag("ri")
        tag_fp, tag_rp = read.get_tag("fp"), read.get_tag("rp")
    except KeyError:
        return None
    try:
        npass_f, npass_r = read.get_tag("fn"), read.get_tag("rn")
    except KeyError:
        npass_f = npass_r = 0
    seq_seq = read.get_forward_sequence()
    n = len(seq_seq)
    if len(tag_fi) != n or len(tag_fp) != n or len(tag_ri) != n or len(tag_rp) != n:
        return None
    sig = []
    for tag in (tag_fi, tag_ri, tag_fp, tag_rp):
        v = np.asarray(tag).astype(np.int64)
        if not args.no_decode:
            v = CODE2FRAMES[v]
        sig.append(_normalize_signals(v, args.norm))
    ipd_f, ipd_r, pw_f, pw_r = sig

    fwd = np.frombuffer(seq_seq.encode("ascii"), dtype=np.uint8)
    rc = _COMP_LUT[fwd[::-1]]
    motif_len = len(motifs[0])
    rev_offset = (motif_len - 1 - args.mod_loc) - args.mod_loc
    nb = (args.seq_len - 1) // 2
    locs = _motif_locs(fwd, motifs, args.mod_loc)
    rev_in_rev = n - 1 - (locs + rev_offset)
    ok = (locs >= nb) & (locs < n - nb) & (rev_in_rev >= nb) & (rev_in_rev < n - nb)
    chrom, strand = ".", "."
    if args.mode == "align":
        # reference :296-301,374-392: only sites inside the aligned part of the query are kept when
        # --skip_unmapped yes (default); reference coordinates are not needed for the modbam output
        reverse = read.is_reverse
        qs, qe = read.query_alignment_start, read.query_alignment_end
        seq_start, seq_end = (n - qe, n - qs) if reverse else (qs, qe)
        if str2bool(args.skip_unmapped):
            ok &= (locs >= seq_start) & (locs < seq_end)
        strand = "-" if reverse else "+"
    locs = locs[ok]
    rev_in_rev = rev_in_rev[ok]
    rf = ReadFeatures()
    rf.holeid, rf.locs = seq_name, locs
    rf.npass_f, rf.npass_r = int(npass_f), int(npass_r)
    L = args.seq_len
    if len(locs) == 0 or n < L:
        z = np.empty((0, L))
        rf.fkmer = rf.rkmer = np.empty((0, L), dtype=np.int64)
        rf.fipd = rf.fpw = rf.ripd = rf.rpw = z
    else:
        fs, rs = locs - nb, rev_in_rev - nb
        rf.fkmer = _CODE_LUT[sliding_window_view(fwd, L)[fs]]
        rf.rkmer = _CODE_LUT[sliding_window_view(rc, L)[rs]]
        rf.fipd = sliding_window_view(ipd_f, L)[fs]
        rf.fpw = sliding_window_view(pw_f, L)[fs]
        rf.ripd = sliding_window_view(ipd_r, L)[rs]
        rf.rpw = sliding_window_view(pw_r, L)[rs]
    rf.sn = None
    if str2bool(args.is_sn):
        try:
            rf.sn = np.around(np.asarray(read.get_tag("sn"), dtype=float), decimals=6)
        except KeyError:
            rf.sn = np.zeros(4)
    rf.chrom, rf.chrom_pos, rf.strand = chrom, default_ref_loc, strand
    rf.fseq, rf.rseq = fwd, rc
    return rf


def to_feature_rows(rf, args):
    """ReadFeatures -> the reference's per-site 22-field rows (extract_features.py:400-405), for tests and for
    feeding the reference-shaped ``_batch_feature_list2s``."""
    rows = []
    nb = (args.seq_len - 1) // 2
    n = len(rf.fseq)
    rev_offset = (len(args.motifs.split(",")[0]) - 1 - args.mod_loc) - args.mod_loc
    sn = rf.sn if rf.sn is not None else "."
    for i, loc in enumerate(rf.locs):
        loc = int(loc)
        rl = n - 1 - (loc + rev_offset)
        fk = rf.fseq[loc - nb:loc + nb + 1].tobytes().decode("ascii")
        rk = rf.rseq[rl - nb:rl + nb + 1].tobytes().decode("ascii")
        rows.append([rf.chrom, rf.chrom_pos, rf.strand, rf.holeid, loc,
                     fk, rf.npass_f, rf.fipd[i], ".", rf.fpw[i], ".", sn, ".",
                     rk, rf.npass_r, rf.ripd[i], ".", rf.rpw[i], ".", sn, ".", args.methy_label])
    return rows


def batch_read_features(read_feats, seq_len):
    """Stack the ReadFeatures of one hole-batch into the float32 (n, L) arrays the model's host entry takes
    (same values as ``_batch_feature_list2s`` + the np.array/FloatTensor stacking of the reference loop).
    Returns (feats dict, holeidx (n,) index into the hole-batch, holeids list, locs (n,))."""
    idx, locs, cols = [], [], {k: [] for k in ("kmer", "kpass", "ipd", "pw", "kmer2", "kpass2", "ipd2", "pw2")}
    for i, rf in read_feats:
        n = len(rf)
        if n == 0:
            continue
        idx.append(np.full(n, i, dtype=np.int64))
        locs.append(rf.locs)
        cols["kmer"].append(rf.fkmer)
        cols["kmer2"].append(rf.rkmer)
        cols["ipd"].append(rf.fipd)
        cols["pw"].append(rf.fpw)
        cols["ipd2"].append(rf.ripd)
        cols["pw2"].append(rf.rpw)
        cols["kpass"].append(np.full((n, seq_len), rf.npass_f))
        cols["kpass2"].append(np.full((n, seq_len), rf.npass_r))
    if not idx:
        return None, np.empty(0, dtype=np.int64), np.empty(0, dtype=np.int64)
    feats = {k: np.ascontiguousarray(np.concatenate(v), dtype=np.float32) for k, v in cols.items()}
    return feats, np.concatenate(idx), np.concatenate(locs)
