"""`call_freqb`: modification frequencies at genome level from an aligned, sorted modbam -- one process per GPU.

Keeps the reference's flag surface and output format (ccsmeth/call_mods_freq_bam.py:741-845, `_write_one_line`
:626-634, file names :639-642) for the default path: symmetric ``--motifs CG``-style sites, count or aggregate mode,
haplotype split by ``--hap_tag``, ``--refsites_only`` / ``--refsites_all``, ``--base_clip``, ``--discrete``,
``--only_close``, ``--sort``, ``--gzip`` (BGZF-compressed; the ``.tbi`` index is left to ``tabix -p bed``).

Where the reference forks region workers that each ``fetch`` their reads through pysam and pile calls up in Python
dictionaries (:457-540), this streams the sorted BAM once: native BGZF inflate + record index (bamstream.py), native
MM/ML-to-reference projection (``ccsm_bam_modcalls``), then per reference chunk (`_get_reference_chunks`, same
boundaries incl. the CG adjustment :62-82) one device call that turns the chunk's pileup into per-site results
(``ccsm_pileup_*``: count statistics, histograms, in-kernel windows, fused aggregate model).

    python -m ccsmeth_b200.call_freqb --input_bam aln.modbam.bam --ref genome.fa -o out --call_mode aggregate -m aggr.ckpt
"""
import argparse
import ctypes
import os
import sys
import time

import numpy as np
import torch

from . import _lib, parallel
from .bamstream import BamPieceReader
from .call_mods import get_motif_seqs
from .call_mods_freq_bam import draw_initial_states, load_aggr_model
from .models import AggrAttRNN
from .utils.process_utils import complement_seq


def read_fasta(path):
    """{contig: upper-case sequence} like the reference's DNAReference (utils/ref_reader.py:33-53)."""
    contigs, name, parts = {}, "", []
    with open(path) as f:
        for line in f:
            if line.startswith(">"):
                if name != "" and parts:
                    contigs[name] = "".join(parts)
                name = line.strip()[1:].split(" ")[0]
                parts = []
            else:
                parts.append(line.strip().upper())
    contigs[name] = "".join(parts)
    return contigs


def get_reference_chunks(dnacontigs, contig_str, chunk_len=500000, motifs="CG"):
    """Reference regions (call_mods_freq_bam.py:48-83): chunks of chunk_len per contig, boundaries moved by one base
    where they would split a CG."""
    if contig_str is not None:
        if os.path.isfile(contig_str):
            with open(contig_str) as f:
                contigs = sorted(set(f.read().splitlines()))
        else:
            contigs = sorted(set(contig_str.strip().split(",")))
    else:
        contigs = sorted(dnacontigs.keys())
    chunks = []
    for contig in contigs:
        n = len(dnacontigs[contig])
        for i in range(0, n, chunk_len):
            chunks.append((contig, i, i + chunk_len if i + chunk_len < n else n))
    if motifs == "CG":
        for idx in range(1, len(chunks)):
            pre_ref, pre_s, pre_e = chunks[idx - 1]
            cur_ref, cur_s, cur_e = chunks[idx]
            if pre_ref != cur_ref:
                continue
            if dnacontigs[pre_ref][(pre_e - 1):(pre_e + 1)] == "CG":
                chunks[idx - 1] = (pre_ref, pre_s, pre_e + 1)
                chunks[idx] = (cur_ref, cur_s + 1, cur_e)
    return chunks


class ModCalls:
    """All modification calls of the reads seen so far, as flat arrays (ref_id, ref_pos, ml, hap, strand)."""

    def __init__(self):
        self.parts = []

    def add_piece(self, piece, opts):
        lib = _lib.load()
        cap = int(piece.recs["l_seq"].sum()) // 8 + 1024
        used = ctypes.c_int32(0)
        while True:
            rid = np.empty(cap, dtype=np.int32)
            pos = np.empty(cap, dtype=np.int32)
            ml = np.empty(cap, dtype=np.uint8)
            hap = np.empty(cap, dtype=np.uint8)
            strand = np.empty(cap, dtype=np.uint8)
            recs = np.ascontiguousarray(piece.recs)
            n = lib.ccsm_bam_modcalls(piece.buf.ctypes.data, recs.ctypes.data, len(recs), ctypes.byref(opts),
                                      rid.ctypes.data, pos.ctypes.data, ml.ctypes.data, hap.ctypes.data,
                                      strand.ctypes.data, cap, ctypes.byref(used))
            if n < 0:
                _lib.check(int(n))
            if n <= cap:
                break
            cap = int(n)
        n = int(n)
        self.parts.append((rid[:n], pos[:n], ml[:n], hap[:n], strand[:n]))
        return used.value

    def arrays(self):
        if not self.parts:
            z = np.zeros(0, dtype=np.int32)
            return z, z, np.zeros(0, np.uint8), np.zeros(0, np.uint8), np.zeros(0, np.uint8)
        return tuple(np.concatenate([p[k] for p in self.parts]) for k in range(5))


def motif_site_masks(seq, motifs, mod_loc):
    """One byte per reference base: is it the modified base of a motif occurrence on the forward / on the reverse
    strand (what get_refloc_of_methysite_in_motif finds on the sequence and on its reverse complement,
    call_mods_freq_bam.py:448-455)."""
    b = np.frombuffer(seq.encode("ascii"), dtype=np.uint8)
    n, mlen = len(b), len(motifs[0])
    comp = np.full(256, ord("N"), dtype=np.uint8)
    for x, y in zip("ACGTNWSMKRYBVDHZ", "TGCANWSKMYRVBHDZ"):
        comp[ord(x)] = ord(y)
    rc = comp[b[::-1]]
    masks = []
    for arr in (b, rc):
        hit = np.zeros(max(n - mlen + 1, 0), dtype=bool)
        for m in set(motifs):
            mb = np.frombuffer(m.encode("ascii"), dtype=np.uint8)
            h = np.ones(len(hit), dtype=bool)
            for k in range(mlen):
                h &= arr[k:k + len(hit)] == mb[k]
            hit |= h
        mask = np.zeros(n, dtype=np.uint8)
        mask[np.nonzero(hit)[0] + mod_loc] = 1
        masks.append(mask)
    return masks[0], np.ascontiguousarray(masks[1][::-1])


def region_pileups(pos, ml, hap, strand, ref_start, ref_end, comb, zero_rule=None, idx=None):
    """Calls of one contig -> the region's CSR pileups.  Returns a list of (strand_char, refpos, ptr, ml, hap):
    one "+" pileup with the reverse-strand CpG calls folded onto the C of the forward strand (pos - 1) when `comb`
    (call_mods_freq_bam.py:542-551), else a "+" and a "-" pileup.  `strand` bit 1 marks --refsites_all zero calls;
    zero_rule = (motif_len, mod_loc): such a call only counts where its motif occurrence lies inside the region, because
    the reference searches the motif in the region's slice of the contig (:448-455).
    idx: the (ascending) indices of the calls with ref_start <= pos < ref_end when the caller already knows them (a
    contig's calls are sorted once and every region is a binary search, instead of one scan of the contig per region)."""
    sel = ((pos >= ref_start) & (pos < ref_end)) if idx is None else idx
    p, m, h, s = pos[sel].astype(np.int64), ml[sel], hap[sel], strand[sel]
    if zero_rule is not None and len(p):
        mlen, mod_loc = zero_rule
        zero, rev = s >= 2, (s & 1) == 1
        lo = np.where(rev, p + mod_loc - (mlen - 1), p - mod_loc)
        hi = np.where(rev, p + mod_loc, p - mod_loc + mlen - 1)
        keep = ~zero | ((lo >= ref_start) & (hi <= ref_end - 1))
        p, m, h, s = p[keep], m[keep], h[keep], s[keep]
    s = s & 1
    out = []
    if comb:
        keep = ~((s == 1) & (p == 0))          # a reverse call at position 0 has no forward partner (:544-545)
        p, m, h, s = p[keep], m[keep], h[keep], s[keep]
        groups = (("+", p - (s == 1), m, h),)
    else:
        groups = (("+", p[s == 0], m[s == 0], h[s == 0]), ("-", p[s == 1], m[s == 1], h[s == 1]))
    for ch, gp, gm, gh in groups:
        if len(gp) == 0:
            continue
        order = np.argsort(gp, kind="stable")
        gp, gm, gh = gp[order], gm[order], gh[order]
        refpos, first = np.unique(gp, return_index=True)
        ptr = np.concatenate((first, [len(gp)])).astype(np.int64)
        out.append((ch, refpos.astype(np.int64), ptr, gm, gh))
    return out


def draw_region_h0(args, n_high):
    """The h0 the reference's region caller would draw: it seeds torch with --tseed, BUILDS the model (parameter
    initialisation consumes the generator) and then draws one randn per 1024-site slice, group after group
    (call_mods_freq_bam.py:310-321, 295-301; models.py:661-671)."""
    torch.manual_seed(args.tseed)
    AggrAttRNN(args.seq_len, args.layer_rnn, args.class_num, 0, args.hid_rnn, binsize=args.bin_size,
               model_type=args.model_type, device="cpu")  # same constructor calls, same generator consumption
    return [draw_initial_states(nh, args.layer_rnn, args.hid_rnn, args.model_type == "attbilstm") if nh else None
            for nh in n_high]


def discretize_score(modprob, coverage):
    """--discrete (reference call_mods_freq_bam.py:240-262): push the model frequency of a site towards whole read
    counts.  Scalar post-processing of the float32 model output, same expressions as the reference."""
    if modprob > 0.66:
        mod_reads = int(np.ceil(modprob * float(coverage)))
    elif modprob <= 0.33:
        mod_reads = int(np.floor(modprob * float(coverage)))
    else:
        mod_reads = round(coverage * modprob, 2)
    unmod_reads = int(coverage) - mod_reads
    adjusted_score = 0.0 if mod_reads == 0 else float(mod_reads) / (mod_reads + unmod_reads)
    return mod_reads, unmod_reads, adjusted_score


def call_region(model, args, contig_seq, ref_name, pileups, motifs_filter):
    """One region's pileups -> (bed_all, bed_hp1, bed_hp2) lists of (ref_name, refpos, strand, cov, cnt, freq), the
    reference's `_readmods_to_bed_of_one_region` return value (:553-594)."""
    beds = ([], [], [])
    if motifs_filter is not None:
        mlen = len(motifs_filter[0])
        fwd_s, fwd_e = -args.mod_loc, mlen - args.mod_loc
        rev_s, rev_e = -(mlen - 1 - args.mod_loc), args.mod_loc + 1
        mset = set(motifs_filter)
    for ch, refpos, ptr, ml, hap in pileups:
        n_high = model.pileup_begin(refpos, ptr, ml, hap, call_mode=args.call_mode, cov_cf=args.cov_cf,
                                    prob_cf=args.prob_cf, no_amb_cov=args.no_amb_cov, no_hap=args.no_hap,
                                    only_close=args.only_close)
        h0 = (None, None, None)
        if args.call_mode == "aggregate" and getattr(args, "h0", "reference") == "reference":
            h0 = draw_region_h0(args, n_high)
        cov, cnt, freq, kind = model.pileup_finish(h0, with_kind=True)
        keep = np.ones(len(refpos), dtype=bool)
        if motifs_filter is not None:
            for i, p in enumerate(refpos.tolist()):
                if ch == "+":
                    keep[i] = contig_seq[p + fwd_s:p + fwd_e] in mset
                else:
                    keep[i] = complement_seq(contig_seq[p + rev_s:p + rev_e]) in mset
        for g in range(3):
            idx = np.nonzero(keep & (kind[g] != 0))[0]
            if len(idx) == 0:
                continue
            pos_l, cov_l, kind_l = refpos[idx].tolist(), cov[g, idx].tolist(), kind[g, idx].tolist()
            # the value types the reference holds (they decide how str() prints them in the output files):
            # kind 1 int count / float freq, kind 2 np.float64 count, kind 3 np.float32 count and freq (model path)
            cnt_g, freq_g = cnt[g, idx], freq[g, idx]
            cnt_i, freq_f = cnt_g.astype(np.int64).tolist(), freq_g.tolist()
            cnt32, freq32 = cnt_g.astype(np.float32), freq_g.astype(np.float32)
            out = beds[g]
            for j, k in enumerate(kind_l):
                if k == 1:
                    c, fr = cnt_i[j], freq_f[j]
                elif k == 2:
                    c, fr = cnt_g[j], freq_f[j]
                else:
                    c, fr = cnt32[j], freq32[j]
                    if args.discrete:
                        c, _, fr = discretize_score(fr, cov_l[j])
                out.append((ref_name, pos_l[j], ch, cov_l[j], c, fr))
    return beds


def write_one_line(beditem, wf, is_bed):
    """reference `_write_one_line` (:626-634)."""
    ref_name, refpos, strand, cov, met, metprob = beditem
    if is_bed:
        wf.write("\t".join([ref_name, str(refpos), str(refpos + 1), ".", str(cov), strand, str(refpos), str(refpos + 1),
                            "0,0,0", str(cov), str(int(round(metprob * 100 + 0.001, 0)))]) + "\n")
    else:
        wf.write("\t".join([ref_name, str(refpos), str(refpos + 1), strand, ".", ".", str(met), str(cov - met),
                            str(cov), str(round(metprob + 0.000001, 4)), "."]) + "\n")


def write_lines(beditems, wf, is_bed):
    """`write_one_line` for a whole list with one write call (same text, line for line; `!s` because an empty format
    spec prints NumPy float32 scalars with double precision digits, unlike the reference's str())."""
    if is_bed:
        wf.write("".join(
            f"{ref_name}\t{refpos}\t{refpos + 1}\t.\t{cov}\t{strand}\t{refpos}\t{refpos + 1}\t0,0,0\t{cov}\t"
            f"{int(round(metprob * 100 + 0.001, 0))}\n" for ref_name, refpos, strand, cov, met, metprob in beditems))
    else:
        wf.write("".join(
            f"{ref_name}\t{refpos}\t{refpos + 1}\t{strand}\t.\t.\t{met!s}\t{cov - met!s}\t{cov}\t"
            f"{round(metprob + 0.000001, 4)!s}\t.\n" for ref_name, refpos, strand, cov, met, metprob in beditems))


def iter_region_results(args, model, dnacontigs, bam_path, rank=0, world=1, piece_bytes=48 << 20):
    """Streams the sorted BAM and yields (region, bed_all, bed_hp1, bed_hp2) for every reference chunk that has calls.
    With world > 1 the chunks are dealt round-robin to the ranks (chunk index % world == rank): regions are
    independent, windows never cross a chunk boundary, so no data moves between ranks."""
    motifs = get_motif_seqs(args.motifs)
    motifs_filter = motifs if (args.refsites_only or args.refsites_all) else None
    comb = args.motifs == "CG" and not args.no_comb
    chunks = get_reference_chunks(dnacontigs, args.contigs, args.chunk_len, args.motifs)
    by_contig = {}
    for ci, c in enumerate(chunks):
        if ci % world == rank:
            by_contig.setdefault(c[0], []).append(c)
    flt = _lib.BamFilter(0, 0, 0, 0, 0)
    rd = BamPieceReader(bam_path, flt, threads=max(1, args.threads), piece_bytes=piece_bytes, align_to=1)
    ref_names = [r[0] for r in rd.references]
    opts = _lib.ModcallOpts(args.mapq, 1 if args.no_supplementary else 0, args.base_clip,
                            args.hap_tag.encode("ascii")[:2], float(args.identity), 0, 0, None, None, None)
    zero_rule = None
    if args.refsites_all:
        # reference motif sites of both strands, all references of the BAM header concatenated
        lens = [len(dnacontigs.get(nm, "")) for nm in ref_names]
        ref_off = np.concatenate(([0], np.cumsum(lens))).astype(np.int64)
        sites_fwd = np.zeros(int(ref_off[-1]) + 1, dtype=np.uint8)
        sites_rev = np.zeros(int(ref_off[-1]) + 1, dtype=np.uint8)
        for i, nm in enumerate(ref_names):
            if nm in by_contig and lens[i]:
                f, r = motif_site_masks(dnacontigs[nm], motifs, args.mod_loc)
                sites_fwd[ref_off[i]:ref_off[i + 1]] = f
                sites_rev[ref_off[i]:ref_off[i + 1]] = r
        opts.refsites_all, opts.n_refs = 1, len(ref_names)
        opts.ref_off, opts.sites_fwd, opts.sites_rev = ref_off.ctypes.data, sites_fwd.ctypes.data, sites_rev.ctypes.data
        zero_rule = (len(motifs[0]), args.mod_loc)
    calls = ModCalls()

    def flush(lo, hi):
        """All reads of references [lo, hi) have been seen (the BAM is coordinate-sorted): call their chunks, then
        forget their calls -- memory stays bounded by one reference sequence's worth of calls."""
        rid, pos, ml, hap, strand = calls.arrays()
        for ref_id in range(lo, hi):
            name = ref_names[ref_id]
            if name not in by_contig:
                continue
            sel = rid == ref_id
            if not sel.any():
                continue
            cpos, cml, chap, cstrand = pos[sel], ml[sel], hap[sel], strand[sel]
            order = np.argsort(cpos, kind="stable")
            spos = cpos[order]
            for region in by_contig[name]:
                _, s, e = region
                lo_i, hi_i = np.searchsorted(spos, np.array([s, e], dtype=spos.dtype))  # same dtype: no copy of spos
                if lo_i == hi_i:
                    continue
                idx = np.sort(order[lo_i:hi_i])  # back to read order: what a scan of the contig would select
                pile = region_pileups(cpos, cml, chap, cstrand, s, e, comb, zero_rule, idx=idx)
                if not pile:
                    continue
                beds = call_region(model, args, dnacontigs[name], name, pile, motifs_filter)
                if beds[0]:
                    yield (region,) + beds
        keep = rid >= hi
        calls.parts = [(rid[keep], pos[keep], ml[keep], hap[keep], strand[keep])] if keep.any() else []

    done = 0
    last_key = -1
    for piece in rd:
        # the streaming flush below relies on coordinate order; the reference fails loudly on other input (it fetches
        # regions through the index), so an unsorted modbam must not silently lose calls here
        offs = piece.recs["off"].astype(np.int64)
        rp = np.frombuffer(piece.buf, dtype=np.uint8)
        rid_pos = np.stack([rp[offs + 4 + k].astype(np.int64) << (8 * (k % 4)) for k in range(8)], axis=0)
        r_id = rid_pos[:4].sum(0).astype(np.uint32).astype(np.int32).astype(np.int64)
        r_id = np.where(r_id < 0, np.int64(1) << 30, r_id)               # unplaced (-1) sorts last
        r_pos = (rid_pos[4:].sum(0).astype(np.uint32).astype(np.int32) + 1).astype(np.int64)
        keys = (r_id << 32) | r_pos
        if len(keys) and (int(keys[0]) < last_key or np.any(keys[1:] < keys[:-1])):
            raise ValueError("%s is not coordinate-sorted: sort it first (samtools sort, or call_mods without --no_sort)"
                             % args.input_bam)
        if len(keys):
            last_key = int(keys[-1])
        calls.add_piece(piece, opts)
        off = int(piece.recs["off"][-1]) + 4
        last_ref = int(np.frombuffer(piece.buf[off:off + 4].tobytes(), dtype="<i4")[0])
        upto = len(ref_names) if last_ref < 0 else min(last_ref, len(ref_names))  # unmapped reads (-1) sort last
        if upto > done:
            yield from flush(done, upto)
            done = upto
    rd.close()
    yield from flush(done, len(ref_names))


def call_freqb(args):
    t0 = time.time()
    if args.call_mode == "aggregate" and not (args.aggre_model and os.path.exists(args.aggre_model)):
        raise ValueError("--aggre_model is not set right!")
    if not args.input_bam.endswith(".bam"):
        raise ValueError("--input_bam not a bam file!")
    if not os.path.exists(args.input_bam):
        raise ValueError("--input_bam does not exist!")
    if not os.path.exists(args.ref):
        raise ValueError("--ref does not exist!")
    os.makedirs(os.path.dirname(os.path.abspath(args.output)), exist_ok=True)
    rank, world, local = parallel.init_from_env()
    dnacontigs = read_fasta(args.ref)
    if args.call_mode == "aggregate":
        model = load_aggr_model(args.aggre_model, args, device=local)
    else:
        model = AggrAttRNN(args.seq_len, args.layer_rnn, args.class_num, 0, args.hid_rnn, binsize=args.bin_size,
                           model_type=args.model_type, device=local).cuda(local).eval()
    fext = "bed" if args.bed else "freq.txt"
    shard = "" if world == 1 else ".rank%d" % rank  # one file set per rank; concatenate (and sort) afterwards
    paths = [args.output + "%s.%s.%s.%s" % (shard, args.call_mode, g, fext) for g in ("all", "hp1", "hp2")]
    files = [open(p, "w") for p in paths]
    n_lines = [0, 0, 0]
    for _, *beds in iter_region_results(args, model, dnacontigs, args.input_bam, rank, world):
        for g in range(3):
            write_lines(beds[g], files[g], args.bed)
            n_lines[g] += len(beds[g])
    for f, p, n in zip(files, paths, n_lines):
        f.close()
        if n == 0:
            os.remove(p)  # the reference removes empty outputs (:663-666)
            continue
        if args.sort or args.gzip:
            # reference :667-676: bedtools sort (chromosome, then start), then bgzip + tabix.  Here: the same ordering
            # and BGZF compression through libccsm's thread team; the .tbi index is left to `tabix -p bed`.
            with open(p) as rf:
                lines = rf.readlines()
            lines.sort(key=lambda ln: (ln.split("\t", 2)[0], int(ln.split("\t", 2)[1])))
            if args.gzip:
                from .bamio import BgzfWriter
                wr = BgzfWriter(p + ".gz", threads=max(1, args.threads), strategy="zlib")  # text: LZ77 pays
                wr.write("".join(lines).encode("ascii"))
                wr.close()
                os.remove(p)
            else:
                with open(p, "w") as wf:
                    wf.writelines(lines)
    total = parallel.allreduce_counts(n_lines + [0])[:3]  # the run's only collective, like call_mods
    if rank == 0:
        sys.stderr.write("[call_freqb] %d / %d / %d sites (all / hp1 / hp2) in %.1f s, %d rank(s)\n"
                         % (*total, time.time() - t0, world))
    return dict(zip(("all", "hp1", "hp2"), total)), paths


def build_parser():
    """The reference's call_freqb flags with the same defaults (call_mods_freq_bam.py:741-840)."""
    p = argparse.ArgumentParser("ccsmeth_b200 call_freqb")
    p.add_argument("--threads", type=int, default=5)
    p.add_argument("--input_bam", type=str, required=True)
    p.add_argument("--ref", type=str, required=True)
    p.add_argument("--contigs", type=str, default=None)
    p.add_argument("--chunk_len", type=int, default=500000)
    p.add_argument("--output", "-o", type=str, required=True)
    p.add_argument("--bed", action="store_true", default=False)
    p.add_argument("--sort", action="store_true", default=False)
    p.add_argument("--gzip", action="store_true", default=False)
    p.add_argument("--modtype", type=str, default="5mC", choices=["5mC"])
    p.add_argument("--call_mode", type=str, default="count", choices=["count", "aggregate"])
    p.add_argument("--prob_cf", type=float, default=0.0)
    p.add_argument("--no_amb_cov", action="store_true", default=False)
    p.add_argument("--hap_tag", type=str, default="HP")
    p.add_argument("--mapq", type=int, default=1)
    p.add_argument("--identity", type=float, default=0.0)
    p.add_argument("--no_supplementary", action="store_true", default=False)
    p.add_argument("--motifs", type=str, default="CG")
    p.add_argument("--mod_loc", type=int, default=0)
    p.add_argument("--no_comb", action="store_true", default=False)
    p.add_argument("--refsites_only", action="store_true", default=False)
    p.add_argument("--refsites_all", action="store_true", default=False)
    p.add_argument("--no_hap", action="store_true", default=False)
    p.add_argument("--base_clip", type=int, default=0)
    p.add_argument("--aggre_model", "-m", type=str, default=None)
    p.add_argument("--model_type", type=str, default="attbigru", choices=["attbilstm", "attbigru"])
    p.add_argument("--seq_len", type=int, default=11)
    p.add_argument("--class_num", type=int, default=1)
    p.add_argument("--layer_rnn", type=int, default=1)
    p.add_argument("--hid_rnn", type=int, default=32)
    p.add_argument("--bin_size", type=int, default=20)
    p.add_argument("--cov_cf", type=int, default=4)
    p.add_argument("--only_close", action="store_true", default=False)
    p.add_argument("--discrete", action="store_true", default=False)
    p.add_argument("--tseed", type=int, default=1234)
    p.add_argument("--h0", type=str, default="reference", choices=["reference", "zeros"],
                   help="ccsmeth_b200 only: GRU initial state of the aggregate model: the reference's per-region "
                        "seeded torch.randn stream (default) or zeros")
    return p


def main(argv=None):
    call_freqb(build_parser().parse_args(argv))
    parallel.finalize()
    return 0


if __name__ == "__main__":
    sys.exit(main())
