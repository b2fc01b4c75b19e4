"""ORACLE (test infrastructure only): per-read modification calls projected onto the reference, restated from the
reference's ``_get_moddict_in_tags`` (call_mods_freq_bam.py:118-168) and the read loop of
``_readmods_to_bed_of_one_region`` (:466-520) over ``ccsmeth_b200.bamio.BamRecord`` objects.  The checker for the
native ``ccsm_bam_modcalls``.  The whole chain is pinned by tests/golden/freqb/ (outputs of the reference's own region
worker on the same synthetic modbam, scripts/gen_golden.py gen_freqb)."""
import re

import numpy as np


def moddict_from_tags(rec, modbase="C", modification="m"):
    """query position (alignment orientation) -> ML byte."""
    try:
        mmtag, mltag = rec.get_tag("MM"), rec.get_tag("ML")
    except KeyError:
        return {}
    fwd = rec.get_forward_sequence()
    deltas = None
    for x in mmtag.split(";"):
        if x.startswith(modbase + "+" + modification):
            start = len(modbase) + 1 + len(modification)
            if len(x) > start and x[start] in "?.":
                start += 1
            if len(x) > start and x[start] == ",":
                deltas = [int(y) for y in x[start + 1:].split(",")]
            break
    if deltas is None:
        return {}
    allpos = [m.start() for m in re.finditer(modbase, fwd)]
    out, count = {}, 0
    try:
        modpos = []
        for d in deltas:
            count += d + 1
            modpos.append(allpos[count - 1])
    except IndexError:
        return {}
    if len(modpos) != len(mltag):
        return {}
    for p, v in zip(modpos, mltag):
        out[len(fwd) - 1 - p if rec.is_reverse else p] = int(v)
    return out


def read_calls(rec, mapq=1, no_supplementary=False, base_clip=0, hap_tag="HP", identity=0.0, refsites=None):
    """-> list of (ref_id, ref_pos, ml, hap, strand) or None if the read is filtered out.  refsites = (fwd set, rev set) of
    reference motif positions switches to --refsites_all: all aligned pairs, uncalled reference sites give (ml 0,
    strand + 2)."""
    if rec.is_unmapped or rec.is_secondary or rec.is_duplicate:
        return None
    if no_supplementary and rec.is_supplementary:
        return None
    if rec.mapq < mapq:
        return None
    base = rec.get_cigar_stats()[0]
    nalign = sum(base[i] for i in range(10) if i not in (4, 5))
    ident = (base[0] + base[7]) / float(nalign) if nalign else 0.0
    if ident < identity:
        return None
    try:
        hap = int(rec.get_tag(hap_tag))
    except (KeyError, ValueError):
        hap = 0
    md = moddict_from_tags(rec)
    pairs = rec.get_aligned_pairs(matches_only=refsites is None)
    if base_clip > 0:
        pairs = pairs[base_clip:(-base_clip)]
    hv, st = hap if hap in (1, 2) else 0, 1 if rec.is_reverse else 0
    out = []
    for q, r in pairs:
        if r is None:
            continue
        if q is not None and q in md:
            out.append((rec.ref_id, r, md[q], hv, st))
        elif refsites is not None and r in refsites[st]:
            out.append((rec.ref_id, r, 0, hv, st + 2))
    return out
