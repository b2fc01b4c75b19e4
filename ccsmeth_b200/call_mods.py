"""`call_mods` for BAM input on one process per GPU: BAM -> features -> model -> MM/ML tags -> modbam.

Keeps the reference's `ccsmeth call_mods` flag surface (ccsmeth/ccsmeth.py:196-326 ==
ccsmeth/call_modifications.py:616-752) and its output rules:
  * per read: predictions sorted by loc; ``MM:Z:C+m?,d0,d1,...;`` deltas = skipped C's of the FORWARD read
    sequence between called C's (_bam2modbam.py:187-203); ``ML:B:C`` = floor(p*256), 255 if p >= 1 (:206-208)
  * existing MM/ML dropped, fi/fp/ri/rp dropped unless --keep_pulse (:215-218)
  * reads without predictions are still written (call_modifications.py:239-242)
  * header gets an ``@PG ID:ccsmeth`` line (:445)
The reference wires reader / extractors / model workers / writer with multiprocessing queues
(call_modifications.py:520-590); here each rank runs a reader thread (BGZF inflate on a native thread team +
native record indexing, bamstream.py), the GPU calls (feature extraction + forward + MM/ML values on the device,
csrc/extract.cu) and a writer thread (native re-tagging + BGZF deflate on the thread team), takes the hole-batches
with ``batch_idx % world == rank`` and writes its own BAM shard.  No per-read Python objects on this path;
``call_reads`` / ``tag_read`` are the record-level equivalents kept for callers that hold BamRecords.  Unless
``--no_sort`` is given the output is coordinate-sorted and indexed (.bai) like the reference's (call_modifications.py:592-607
runs samtools sort / index through pysam; here bamsort.py), with the rank shards merged into one file by rank 0.

    python -m ccsmeth_b200.call_mods -i in.hifi.bam -m model.ckpt -o out_prefix [--mode denovo] ...
"""
import argparse
import os
import queue
import sys
import threading
import time

import numpy as np
import torch

from . import VERSION, _lib, parallel
from .bamio import BamWriter, add_pg_line
from .bamstream import BamPieceReader, tag_records
from .call_modifications import draw_h0_stream_batches, load_model
from .extract_features import ReadBatch, extract_opts, pack_reads
from .utils.process_utils import str2bool

TIMING = {}  # stage -> seconds, filled when CCSM_TIMING=1 (reader / pack / extract / forward / tag / write)


def _tic(key, t0):
    TIMING[key] = TIMING.get(key, 0.0) + (time.perf_counter() - t0)


IUPAC = {'A': 'A', 'C': 'C', 'G': 'G', 'T': 'T', 'R': 'AG', 'M': 'AC', 'S': 'CG', 'Y': 'CT', 'K': 'GT', 'W': 'AT',
         'B': 'CGT', 'D': 'AGT', 'H': 'ACT', 'V': 'ACG', 'N': 'ACGT'}


def get_motif_seqs(motifs):
    """IUPAC motif expansion (reference process_utils.py:140-170)."""
    out = []
    for m in motifs.strip().split(","):
        seqs = [""]
        for b in m.strip().upper():
            seqs = [s + x for s in seqs for x in IUPAC[b]]
        out += seqs
    return out


def convert_locs_to_mmtag(locs, fwd_seq_bytes, base=ord('C')):
    """MM deltas (reference _bam2modbam.py:187-203): order of each called loc among the read's C's, then
    first order followed by gaps - 1.  Raises AssertionError like the reference when a loc is not a C."""
    assert len(locs) > 0
    base_all = np.nonzero(fwd_seq_bytes == base)[0]
    orders = np.searchsorted(base_all, locs)
    assert orders[-1] < len(base_all) and np.array_equal(base_all[np.minimum(orders, len(base_all) - 1)], locs)
    mm = np.empty(len(orders), dtype=np.int64)
    mm[0] = orders[0]
    mm[1:] = np.diff(orders) - 1
    return mm


def convert_probs_to_mltag(probs):
    """floor(p * 256), 255 if p >= 1, in float32 like the reference's np.float32 scalars (_bam2modbam.py:206-208)."""
    p = np.asarray(probs, dtype=np.float32)
    return np.where(p < 1, np.floor(p * np.float32(256)), 255).astype(np.uint8)


def _announce_or_draw_h0(model, args, per_hb, h0):
    """The reference's h0 stream restarts its ``--batch_size`` slicing at every hole-batch: "reference" announces the
    hole-batch site counts to the library, which draws the stream on the device; "reference_host" draws it with
    torch.randn here (12 KB/site over PCIe); explicit h0 and the other modes pass through."""
    mode = getattr(args, "h0", "reference")
    if h0 is not None or getattr(model, "rnn_cell", None) != "gru":
        return h0
    if mode == "reference_host":
        return draw_h0_stream_batches(per_hb, args.batch_size, model.num_layers, model.hidden_size)
    if mode == "reference":
        model.set_h0_batching(per_hb, args.batch_size)
    return None


def call_reads(model, reads, motifs, args, holeids_e=None, holeids_ne=None, h0=None, holes_batch=None):
    """A run of consecutive hole-batches (list of BamRecord) -> (per-read list of (locs, prob_1_norm, mm, ml) or None,
    n_sites, n_model_batches).  Equivalent of process_one_holebatch + _batch_feature_list2s + _call_mods2s + the
    MM/ML conversion (reference extract_features.py:409-431, call_modifications.py:73-123,170-227,
    _bam2modbam.py:187-208) with the per-read work done on the device (csrc/extract.cu).

    ``holes_batch``: reads per hole-batch inside `reads` (default: all of `reads` is one hole-batch).  It only
    matters for the reference's h0 stream and batch counter, which restart their 512-slicing at every hole-batch."""
    per_read = [None] * len(reads)
    t0 = time.perf_counter()
    batch = pack_reads(reads, args, holeids_e, holeids_ne)
    _tic("pack", t0)
    if len(batch) == 0:
        return per_read, 0, 0
    t0 = time.perf_counter()
    n = model.extract_reads(batch, extract_opts(args, motifs))
    _tic("extract", t0)
    if n == 0:
        return per_read, 0, 0
    site_read, site_loc = model.reads_sites()
    hb = holes_batch or len(reads)
    rec_idx = np.asarray(batch.index, dtype=np.int64)[site_read]       # site -> position in `reads`
    per_hb = np.bincount(rec_idx // hb, minlength=(len(reads) + hb - 1) // hb)
    n_batches = int(sum((c + args.batch_size - 1) // args.batch_size for c in per_hb))
    h0 = _announce_or_draw_h0(model, args, per_hb, h0)
    t0 = time.perf_counter()
    res = model.reads_forward(h0=h0, want_probs=False)
    _tic("forward", t0)
    bounds = np.nonzero(np.diff(site_read))[0] + 1
    starts = np.concatenate(([0], bounds))
    ends = np.concatenate((bounds, [n]))
    for s, e in zip(starts, ends):
        per_read[batch.index[int(site_read[s])]] = (site_loc[s:e], res["prob1"][s:e], res["mm"][s:e], res["ml"][s:e])
    return per_read, n, n_batches


def call_holebatch(model, reads, motifs, args, holeids_e=None, holeids_ne=None, h0=None):
    """One hole-batch; returns (per-read list of (locs, prob_1_norm) or None, n_sites, n_model_batches)."""
    per_read, n, nb = call_reads(model, reads, motifs, args, holeids_e, holeids_ne, h0)
    return [None if p is None else (p[0], p[1]) for p in per_read], n, nb


def tag_read(rec, pred, rm_pulse, mod_base_is_c=True):
    """BamRecord + (locs, probs[, mm, ml]) -> (record bytes with MM/ML, mm_flag)
    (reference call_modifications.py:230-266).  With the device outputs (mm, ml) present nothing is recomputed;
    the 2-tuple form converts on the host like the reference does."""
    drop = {"MM", "ML"} | ({"fi", "fp", "ri", "rp"} if rm_pulse else set())
    if pred is None or len(pred[0]) == 0:
        return rec.with_tags(drop), 0
    if len(pred) == 4:
        if not mod_base_is_c:  # the reference counts C's only (_bam2modbam.py:187-199): assertion -> no tags
            return rec.with_tags(drop), 0
        mm, ml = pred[2], pred[3]
        return rec.with_tags(drop, "C+m?," + ",".join(map(str, mm.tolist())) + ";", ml.tobytes()), 1
    locs, probs = pred
    order = np.argsort(locs, kind="stable")
    locs, probs = locs[order], probs[order]
    fwd = np.frombuffer(rec.get_forward_sequence().encode("ascii"), dtype=np.uint8)
    try:
        mm = convert_locs_to_mmtag(locs, fwd)
    except AssertionError:
        return rec.with_tags(drop), 0  # reference writes the read without tags (:260-263)
    ml = convert_probs_to_mltag(probs)
    return rec.with_tags(drop, "C+m?," + ",".join(map(str, mm.tolist())) + ";", ml.tolist()), 1


def call_piece(model, piece, motifs, args, rank=0, world=1, holeids_e=None, holeids_ne=None):
    """One bamstream.Piece -> (recs of this rank, site_begin per read, mm, ml, n_sites, n_model_batches).
    Same work as ``call_reads`` without per-read Python objects: the piece's buffer is the device blob."""
    recs = piece.recs
    descs = piece.descs
    hb = args.holes_batch
    gidx = piece.first + np.arange(len(recs), dtype=np.int64)
    own = np.ones(len(recs), dtype=bool) if world == 1 else ((gidx // hb) % world) == rank
    use = own & (recs["read_idx"] >= 0)
    if holeids_e is not None or holeids_ne is not None:
        for i, nm in enumerate(piece.names()):
            nm = nm.decode("ascii", "replace")
            if (holeids_e is not None and nm not in holeids_e) or (holeids_ne is not None and nm in holeids_ne):
                use[i] = False
    recs = recs[own].copy()
    sel = piece.recs["read_idx"][use]                 # descs that take part, in record order
    remap = np.full(len(descs) + 1, -1, dtype=np.int32)
    remap[sel] = np.arange(len(sel), dtype=np.int32)
    recs["read_idx"] = np.where(use[own], remap[recs["read_idx"]], -1)
    site_begin = np.zeros(len(sel) + 1, dtype=np.int64)
    if len(sel) == 0:
        return recs, site_begin, None, None, 0, 0
    t0 = time.perf_counter()
    batch = ReadBatch(piece.buf, np.ascontiguousarray(descs[sel]), None)
    n = model.extract_reads(batch, extract_opts(args, motifs))
    _tic("extract", t0)
    if n == 0:
        return recs, site_begin, None, None, 0, 0
    site_read, _ = model.reads_sites()
    np.cumsum(np.bincount(site_read, minlength=len(sel)), out=site_begin[1:])
    # the reference's bookkeeping restarts at every hole-batch: batch counter and, in --h0 reference, the randn stream
    site_hb = (gidx[use] // hb)[site_read]
    per_hb = np.diff(np.concatenate(([0], np.nonzero(np.diff(site_hb))[0] + 1, [n])))
    n_batches = int(((per_hb + args.batch_size - 1) // args.batch_size).sum())
    h0 = _announce_or_draw_h0(model, args, per_hb, None)
    t0 = time.perf_counter()
    res = model.reads_forward(h0=h0, want_probs=False)
    _tic("forward", t0)
    return recs, site_begin, res["mm"], res["ml"], n, n_batches


def _reader_thread(rd, q):
    """Inflates and indexes the next pieces while the GPU works on the current one."""
    try:
        t0 = time.perf_counter()
        for piece in rd:
            _tic("reader", t0)
            q.put(("piece", piece))
            t0 = time.perf_counter()
        q.put(("done",))
    except Exception as e:  # surface reader failures to the main thread
        q.put(("error", e))


def _writer_thread(wr, q, keep_pulse, mod_base_is_c, counts, err, indexer=None):
    """Re-tags and writes the records of finished pieces (reference _worker_write_modbam,
    call_modifications.py:410-462).  `indexer` (bamsort.StreamIndexer) sees the same bytes, so that an output whose
    records are already in coordinate order gets its .bai without being read back."""
    try:
        while True:
            msg = q.get()
            if msg is None:
                return
            piece, recs, site_begin, mm, ml = msg
            t0 = time.perf_counter()
            if not mod_base_is_c:  # the reference counts C's only (_bam2modbam.py:187-199): assertion -> no tags
                mm = ml = None
            data, with_mm = tag_records(piece, recs, keep_pulse, site_begin, mm, ml)
            wr.bg.write(data)
            if indexer is not None:
                indexer.feed(data)
            counts[2] += len(recs)
            counts[3] += with_mm
            _tic("tag+write", t0)
    except Exception as e:
        err.append(e)
        while q.get() is not None:  # keep draining so the producer never blocks
            pass


def call_mods(args):
    """Runs the whole path; returns the summed run counters
    {sites, model_batches, reads_written, reads_with_mm} (all ranks)."""
    t0 = time.time()
    if args.seq_len % 2 == 0:
        raise ValueError("--seq_len must be odd")
    if not os.path.exists(args.model_file):
        raise ValueError("--model_file is not set right!")
    if not os.path.exists(args.input):
        raise ValueError("--input_file does not exist!")
    if not (args.input.endswith(".bam")):
        raise ValueError("ccsmeth_b200 call_mods takes BAM input (features.tsv input is out of scope)")
    if args.model_type in ("attbilstm2s", "attbilstm2s2") and getattr(args, "h0", "reference") in ("reference", "reference_host"):
        raise ValueError("--model_type %s: the BAM pipeline takes the LSTM initial state from the library; "
                         "pass --h0 device or --h0 zeros" % args.model_type)
    if args.model_type in ("attbigru2s2", "attbilstm2s2", "transencoder2s") and args.norm != "none":
        raise ValueError("--model_type %s embeds the kinetics as integers: run it with --norm none" % args.model_type)
    if str2bool(args.is_map) or str2bool(args.is_stds):
        raise ValueError("--is_map / --is_stds features are not extracted by ccsmeth_b200 (the reference's extractor itself "
                         "writes '.' for the std columns, extract_features.py:353-364; --is_map needs the reference FASTA walk "
                         "of extract_features.py:202-258)")
    rank, world, local = parallel.init_from_env()
    TIMING.clear()
    out_dir = os.path.dirname(os.path.abspath(args.output))
    os.makedirs(out_dir, exist_ok=True)
    out_modbam = args.output + (".modbam.bam" if world == 1 else ".rank%d.modbam.bam" % rank)
    t_load = time.perf_counter()
    model = load_model(args.model_file, args, device=local, precision=getattr(args, "precision", None))
    model._ensure_handle()
    _tic("model_load", t_load)
    # seed the process that draws h0, after model construction (which itself consumes the generator); the
    # reference seeds only its parent process (:479-481), so its workers' h0 streams are not reproducible
    torch.manual_seed(args.tseed + rank)
    model.set_h0_mode(getattr(args, "h0", "reference"), seed=args.tseed + rank)
    motifs = get_motif_seqs(args.motifs)
    mod_base_is_c = all(m[args.mod_loc] == "C" for m in motifs)
    holeids_e = _get_holes(args.holeids_e) if args.holeids_e else None
    holeids_ne = _get_holes(args.holeids_ne) if args.holeids_ne else None

    threads = max(1, args.threads)
    flt = _lib.BamFilter(1 if args.mode == "align" else 0, args.mapq, 1 if args.no_supplementary else 0,
                         1 if str2bool(args.skip_unmapped) else 0, 1 if str2bool(args.is_sn) else 0,
                         identity=args.identity if args.mode == "align" else 0.0)
    # a piece = `device_batch` hole-batches' worth of compressed bytes (about 3 MB per 50 HiFi reads)
    rd = BamPieceReader(args.input, flt, threads=threads,
                        piece_bytes=max(1, getattr(args, "device_batch", 8)) * args.holes_batch * 65536,
                        align_to=args.holes_batch)
    wr = BamWriter(out_modbam, add_pg_line(rd.header_text, VERSION, " ".join(sys.argv)), rd.references, threads=threads,
                   strategy=getattr(args, "bam_compress", "rle"))
    q = queue.Queue(maxsize=2)
    th = threading.Thread(target=_reader_thread, args=(rd, q), daemon=True)
    th.start()
    counts = [0, 0, 0, 0]
    wq, werr = queue.Queue(maxsize=2), []
    indexer = None
    if not args.no_sort and world == 1:
        from . import bamsort
        indexer = bamsort.StreamIndexer(wr.header_bytes)
    wth = threading.Thread(target=_writer_thread, args=(wr, wq, args.keep_pulse, mod_base_is_c, counts, werr, indexer),
                           daemon=True)
    wth.start()
    try:
        while True:
            msg = q.get()
            if msg[0] == "done":
                break
            if msg[0] == "error":
                raise msg[1]
            piece = msg[1]
            recs, site_begin, mm, ml, n_sites, n_batches = call_piece(model, piece, motifs, args, rank, world,
                                                                      holeids_e, holeids_ne)
            counts[0] += n_sites
            counts[1] += n_batches
            wq.put((piece, recs, site_begin, mm, ml))
            if werr:
                raise werr[0]
    finally:
        wq.put(None)
        wth.join()
    if werr:
        raise werr[0]
    wr.close()
    rd.close()
    total = parallel.allreduce_counts(counts)
    if not args.no_sort:
        # reference call_modifications.py:592-607: samtools sort + index of the modbam unless --no_sort
        t_sort = time.perf_counter()
        out_modbam = _sort_and_index(args, out_modbam, rank, world, threads, indexer, len(rd.references))
        _tic("sort_index", t_sort)
    if rank == 0:
        dt = time.time() - t0
        sys.stderr.write("[call_mods] %d sites in %d model batches(%d), wrote %d reads, in which %d were added mm "
                         "tags; %.1f s, %d rank(s)\n" % (total[0], total[1], args.batch_size, total[2], total[3], dt, world))
        if os.environ.get("CCSM_TIMING"):
            sys.stderr.write("[call_mods] stage seconds (threads overlap): %s\n" %
                             ", ".join("%s %.3f" % kv for kv in sorted(TIMING.items())))
    return dict(zip(("sites", "model_batches", "reads_written", "reads_with_mm"), total)), out_modbam


def _sort_and_index(args, out_modbam, rank, world, threads, indexer=None, n_refs=0):
    """Coordinate sort + .bai of the output (ccsmeth_b200/bamsort.py).  One rank: in place -- and when the records are
    already in coordinate order (a sorted input keeps its order; an unaligned input has only unplaced reads) only the
    index is written, from what the writer thread's StreamIndexer collected (no second pass over the file).  Several
    ranks (one node): rank 0 merges the sorted shards into <output>.modbam.bam."""
    from . import bamsort
    final = args.output + ".modbam.bam"
    comp = getattr(args, "bam_compress", "rle")
    if world == 1:
        done = indexer.finish(out_modbam, n_refs) if indexer is not None else bamsort.index_sorted(out_modbam, threads=threads)
        if done < 0:
            bamsort.sort_and_index(out_modbam, out_modbam, threads=threads, bam_compress=comp)
        return out_modbam
    parallel.barrier()
    if rank == 0:
        shards = [args.output + ".rank%d.modbam.bam" % r for r in range(world)]
        bamsort.sort_and_index(shards, final, threads=threads, bam_compress=comp)
        for p in shards:
            os.remove(p)
    parallel.barrier()
    return final


def _get_holes(path):
    with open(path) as f:
        return {ln.strip().split("\t")[0] for ln in f if ln.strip()}


def build_parser():
    """The reference's call_mods flags with the same defaults (call_modifications.py:616-752)."""
    p = argparse.ArgumentParser("ccsmeth_b200 call_mods", description="call modifications with the B200-native path")
    p.add_argument("--input", "-i", type=str, required=True)
    p.add_argument("--holes_batch", type=int, default=50)
    p.add_argument("--output", "-o", type=str, required=True)
    p.add_argument("--gzip", action="store_true", default=False)
    p.add_argument("--keep_pulse", action="store_true", default=False)
    p.add_argument("--no_sort", action="store_true", default=False)
    p.add_argument("--model_file", "-m", type=str, required=True)
    p.add_argument("--model_type", type=str, default="attbigru2s",
                   choices=["attbilstm2s", "attbigru2s", "transencoder2s", "attbilstm2s2", "attbigru2s2"])
    p.add_argument("--seq_len", type=int, default=21)
    p.add_argument("--is_npass", type=str, default="yes")
    p.add_argument("--is_stds", type=str, default="no")
    p.add_argument("--is_sn", type=str, default="no")
    p.add_argument("--is_map", type=str, default="no")
    p.add_argument("--class_num", type=int, default=2)
    p.add_argument("--dropout_rate", type=float, default=0)
    p.add_argument("--batch_size", "-b", type=int, default=512)
    p.add_argument("--layer_rnn", type=int, default=3)
    p.add_argument("--hid_rnn", type=int, default=256)
    p.add_argument("--layer_trans", type=int, default=6)
    p.add_argument("--nhead", type=int, default=4)
    p.add_argument("--d_model", type=int, default=256)
    p.add_argument("--dim_ff", type=int, default=512)
    p.add_argument("--mode", type=str, default="denovo", choices=["denovo", "align"])
    p.add_argument("--holeids_e", type=str, default=None)
    p.add_argument("--holeids_ne", type=str, default=None)
    p.add_argument("--motifs", type=str, default="CG")
    p.add_argument("--mod_loc", type=int, default=0)
    p.add_argument("--methy_label", type=int, default=1, choices=[1, 0])
    p.add_argument("--norm", type=str, default="zscore", choices=["zscore", "min-mean", "min-max", "mad", "none"])
    p.add_argument("--no_decode", action="store_true", default=False)
    p.add_argument("--ref", type=str, default=None)
    p.add_argument("--mapq", type=int, default=1)
    p.add_argument("--identity", type=float, default=0.0)
    p.add_argument("--no_supplementary", action="store_true", default=False)
    p.add_argument("--skip_unmapped", type=str, default="yes")
    p.add_argument("--threads", "-p", type=int, default=10)
    p.add_argument("--threads_call", type=int, default=3)
    p.add_argument("--tseed", type=int, default=1234)
    p.add_argument("--use_compile", type=str, default="no")
    p.add_argument("--h0", type=str, default="reference", choices=["reference", "reference_host", "device", "zeros"],
                   help="ccsmeth_b200 only: GRU initial state: the reference's torch.randn stream reproduced on the "
                        "device bit for bit (default), the same stream drawn by torch on the host (12 KB/site over "
                        "PCIe), N(0,1) from the device's own generator, or zeros")
    p.add_argument("--device_batch", type=int, default=8,
                   help="ccsmeth_b200 only: hole-batches per device call (features are extracted on the GPU for "
                        "this many x --holes_batch reads at once)")
    p.add_argument("--bam_compress", type=str, default="rle", choices=["rle", "zlib"],
                   help="ccsmeth_b200 only: BGZF block compression of the output modbam: 'rle' = run-length + Huffman "
                        "(zlib Z_RLE; 3-4x faster on HiFi records, size within 3 %%), 'zlib' = default strategy, level 6")
    p.add_argument("--precision", type=str, default=None, choices=["fp32", "fp16c8", "fp16x3", "bf16x3", "fp16", "bf16"],
                   help="ccsmeth_b200 only: arithmetic mode (default fp16c8, <= 1e-4 vs the fp32 reference)")
    return p


def main(argv=None):
    args = build_parser().parse_args(argv)
    counts, out = call_mods(args)
    parallel.finalize()
    return 0


if __name__ == "__main__":
    sys.exit(main())
