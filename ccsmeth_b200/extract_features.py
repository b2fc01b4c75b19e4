"""Host glue of the device feature extractor (include/ccsm.h ``ccsm_reads_*``, kernels in csrc/extract.cu).

The reference extracts features per read in Python (ccsmeth/extract_features.py:261-406) and re-batches them
into 16 tensors (call_modifications.py:73-123).  Here the host only decides WHICH reads take part (the
read-level filters of extract_features.py:269-288,321-326 and the tag checks of :88-126) and tells the device
WHERE each read's sequence and kinetics arrays sit inside the raw (inflated) BAM records; CodecV1 decoding,
per-read normalisation, the motif scan, the 21-base windows of both strands and the MM/ML values are computed
on the device.  ``pack_reads`` produces that description; ``ModelAttRNN.extract_reads / reads_forward``
(ccsmeth_b200/models.py) hand it to the library.
"""
import numpy as np

from .utils.process_utils import str2bool

# mirrors `ccsm_read` in include/ccsm.h (80 bytes, natural alignment)
READ_DTYPE = np.dtype([("seq_off", "<i8"), ("fi_off", "<i8"), ("ri_off", "<i8"), ("fp_off", "<i8"), ("rp_off", "<i8"),
                       ("len", "<i4"), ("fn", "<i4"), ("rn", "<i4"), ("flags", "<i4"), ("win_lo", "<i4"),
                       ("win_hi", "<i4"), ("sn", "<f4", (4,))], align=True)
assert READ_DTYPE.itemsize == 80
READ_REVERSE, READ_SEQ_4BIT = 1, 2
NORM = {"zscore": 0, "min-mean": 1, "min-max": 2, "none": 3, "mad": 4}


class ReadBatch:
    """blob: contiguous uint8 array holding the raw records; descs: READ_DTYPE array, one per kept read;
    index: position of each kept read in the caller's record list."""
    __slots__ = ("blob", "descs", "index")

    def __init__(self, blob, descs, index):
        self.blob, self.descs, self.index = blob, descs, index

    def __len__(self):
        return len(self.descs)


def _uint8_tag_offset(rec, name):
    """Offset (inside rec.raw) of the data of a B:C aux array, and its element count."""
    val, start, _ = rec._tags[name]
    if rec.raw[start + 2] != ord('B') or rec.raw[start + 3] != ord('C'):
        raise ValueError("read %s: tag %s is not a B:C (uint8) array; ccsmeth_b200 reads CodecV1 kinetics"
                         % (rec.query_name, name))
    return start + 8, len(val)


def read_passes_filters(rec, args, holeids_e=None, holeids_ne=None):
    """Read-level filters of the reference extractor (extract_features.py:269-288)."""
    if holeids_e is not None or holeids_ne is not None:
        name = rec.query_name
        if holeids_e is not None and name not in holeids_e:
            return False
        if holeids_ne is not None and name in holeids_ne:
            return False
    if args.mode == "align":
        if rec.is_unmapped or rec.is_secondary or rec.is_duplicate:
            return False
        if args.no_supplementary and rec.is_supplementary:
            return False
        if rec.mapq < args.mapq:
            return False
        if getattr(args, "identity", 0.0) > 0.0 and cigar_identity(rec) < args.identity:
            return False
    return True


def cigar_identity(rec):
    """compute_pct_identity of the reference (process_utils.py:174-186) on a bamio.BamRecord: matches (M, =) over every
    aligned operation except clips; 0 for a CIGAR-less record."""
    nalign = nmatch = 0
    for op, ln in rec.cigartuples or ():
        if op not in (4, 5) and op <= 9:
            nalign += ln
        if op in (0, 7):
            nmatch += ln
    return nmatch / float(nalign) if nalign else 0.0


def pack_reads(records, args, holeids_e=None, holeids_ne=None):
    """list of bamio.BamRecord -> ReadBatch.  Reads the reference would skip (filters, missing or incomplete
    kinetics tags, extract_features.py:321-326) are left out; they still reach the writer untouched."""
    descs = np.zeros(len(records), dtype=READ_DTYPE)
    index, chunks = [], []
    base = 0
    k = 0
    for i, rec in enumerate(records):
        if not read_passes_filters(rec, args, holeids_e, holeids_ne):
            continue
        if rec._tags is None:
            rec._parse_tags()
        tags = rec._tags
        if not ("fi" in tags and "ri" in tags and "fp" in tags and "rp" in tags):
            continue
        n = rec.l_seq
        offs = []
        ok = True
        for t in ("fi", "ri", "fp", "rp"):
            off, cnt = _uint8_tag_offset(rec, t)
            ok &= cnt == n
            offs.append(base + off)
        if not ok or n == 0:
            continue
        d = descs[k]
        d["seq_off"] = base + 32 + rec.l_read_name + 4 * rec.n_cigar
        d["fi_off"], d["ri_off"], d["fp_off"], d["rp_off"] = offs
        d["len"] = n
        if "fn" in tags and "rn" in tags:  # extract_features.py:113-117: both or neither
            d["fn"], d["rn"] = int(tags["fn"][0]), int(tags["rn"][0])
        reverse = rec.is_reverse
        d["flags"] = READ_SEQ_4BIT | (READ_REVERSE if reverse else 0)
        lo, hi = 0, n
        if args.mode == "align" and str2bool(args.skip_unmapped):
            qs, qe = rec.query_alignment_start, rec.query_alignment_end
            lo, hi = (n - qe, n - qs) if reverse else (qs, qe)  # extract_features.py:296-301,374,388-390
        d["win_lo"], d["win_hi"] = lo, hi
        if str2bool(args.is_sn) and "sn" in tags:
            sn = np.asarray(tags["sn"][0], dtype=np.float32)
            d["sn"][:min(4, len(sn))] = sn[:4]
        chunks.append(rec.raw)
        base += len(rec.raw)
        index.append(i)
        k += 1
    blob = np.frombuffer(b"".join(chunks), dtype=np.uint8) if chunks else np.zeros(0, dtype=np.uint8)
    return ReadBatch(blob, descs[:k].copy(), index)


def extract_opts(args, motifs):
    """CLI args + expanded motif list -> the fields of `ccsm_extract_opts` (dict)."""
    if args.norm not in NORM:
        raise ValueError("--norm %s is not one of %s" % (args.norm, sorted(NORM)))
    mlen = len(motifs[0])
    if any(len(m) != mlen for m in motifs):
        raise ValueError("all --motifs must have the same length")
    return {"mod_loc": int(args.mod_loc), "norm": NORM[args.norm], "decode": 0 if args.no_decode else 1,
            "motifs": list(dict.fromkeys(motifs))}
