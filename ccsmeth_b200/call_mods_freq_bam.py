"""Aggregate-mode frequency model, host side (the reference's ccsmeth/call_mods_freq_bam.py, hot part only).

Stands in for ``_cal_modfreq_in_aggregate_mode`` (reference call_mods_freq_bam.py:265-305): every site sees the
histograms of its 11-site neighbourhood (zero rows beyond the region's ends) and the distances to them, the model
runs, the output is clipped to [0, 1] and rounded to 6 places.  The reference builds the (n, 11, 20) windows on the
host, runs the model in batches of 1024 on the CPU and rebuilds + reloads it for every region (:308-342); here the
model is resident on the GPU, the kernel gathers the neighbourhoods from the per-site rows, the whole region goes
through in one call, and the h0 stream is still drawn per 1024-slice in the reference's order.
"""
from collections import OrderedDict

import numpy as np
import torch

from .models import AggrAttRNN

AGGR_BATCH = 1024  # reference call_mods_freq_bam.py:295


def _cal_modfreq_in_aggregate_mode(refposes, refposes_histos, model, seq_len=11, only_close=False, h0=None):
    """Same arguments and result as the reference's function (call_mods_freq_bam.py:265-305).  The shipped model
    configuration hands the per-site rows to the device, where the kernel gathers each site's neighbourhood
    (``AggrAttRNN.forward_sites``); other configurations get their windows from ``site_windows`` below."""
    n = len(refposes)
    if n == 0:
        return None
    pos = np.asarray(refposes, dtype=np.int64)
    rows = np.asarray(refposes_histos, dtype=np.float32).reshape(n, -1)
    if h0 is None:
        h0 = draw_initial_states(n, model.num_layers, model.hidden_size, model.rnn_cell == "lstm")
    if model.fused() and seq_len == model.seq_len:
        out = model.forward_sites(pos, rows, h0=h0, only_close=only_close)
    else:
        offsets, windows = site_windows(pos, rows, seq_len, only_close)
        out = model(torch.from_numpy(offsets), torch.from_numpy(windows), h0=h0)
    freq = np.round(np.clip(out.cpu().numpy(), 0, 1), 6)
    return [freq[i][0] for i in range(n)]


def site_windows(pos, rows, seq_len=11, only_close=False):
    """Neighbourhood tensors of a region's sites, by index gather: window slot k of site i is site i + k - seq_len // 2.
    Slots outside the region hold a zero histogram and sit 1000 bp before the first / after the last site
    (reference call_mods_freq_bam.py:272-283).  Returns (offsets (n, seq_len) float32, windows (n, seq_len, bins) float32):
    offsets = distance to the centre site, or with ``only_close`` 1.0 where a slot's site lies exactly 2 bp after the
    previous slot's site (:285-290)."""
    n, half = len(pos), seq_len // 2
    nb = np.arange(n)[:, None] + np.arange(-half, half + 1)[None, :]       # (n, seq_len) site index of every slot
    inside = (nb >= 0) & (nb < n)
    at = np.clip(nb, 0, n - 1)
    windows = np.where(inside[:, :, None], rows[at], np.float32(0))
    slot_pos = np.where(nb < 0, pos[0] - 1000, np.where(nb >= n, pos[-1] + 1000, pos[at]))
    if only_close:
        before = nb - 1                                                     # the slot one step to the left
        prev_pos = np.where(before < 0, pos[0] - 1000, np.where(before >= n, pos[-1] + 1000, pos[np.clip(before, 0, n - 1)]))
        offsets = (slot_pos - prev_pos == 2)
    else:
        offsets = np.abs(slot_pos - pos[:, None])
    return np.ascontiguousarray(offsets, dtype=np.float32), np.ascontiguousarray(windows, dtype=np.float32)


def draw_initial_states(n, num_layers, hidden_size, lstm=False):
    """The initial states the reference's batch loop draws for n sites: per 1024-site slice one ``torch.randn`` for h0
    and, for the LSTM cell, a second one for c0 (models.py:661-671, call_mods_freq_bam.py:295-301).
    -> h0 (2*layers, n, hidden), or the pair (h0, c0)."""
    # each slice is one contiguous torch.randn (the reference's call), placed with a plain memory copy
    h0 = np.empty((2 * num_layers, n, hidden_size), dtype=np.float32)
    c0 = np.empty((2 * num_layers, n, hidden_size), dtype=np.float32) if lstm else None
    for s in range(0, n, AGGR_BATCH):
        e = min(n, s + AGGR_BATCH)
        h0[:, s:e] = torch.randn(2 * num_layers, e - s, hidden_size).numpy()
        if lstm:
            c0[:, s:e] = torch.randn(2 * num_layers, e - s, hidden_size).numpy()
    h0 = torch.from_numpy(h0)
    c0 = torch.from_numpy(c0) if lstm else None
    return (h0, c0) if lstm else h0


def load_aggr_model(model_path, args, device=0):
    """Model lifecycle of the reference's per-region loader (call_mods_freq_bam.py:317-342), done once."""
    if args.model_type not in {"attbigru", "attbilstm"}:
        raise ValueError("--model_type not right!")
    model = AggrAttRNN(args.seq_len, args.layer_rnn, args.class_num, 0, args.hid_rnn, binsize=args.bin_size,
                       model_type=args.model_type, device=device)
    para_dict = torch.load(model_path, map_location=torch.device('cpu'))
    try:
        model_dict = model.state_dict()
        model_dict.update(para_dict)
        model.load_state_dict(model_dict)
    except RuntimeError:
        model.load_state_dict(OrderedDict((k[7:], v) for k, v in para_dict.items()))
    model = model.cuda(device)
    model.eval()
    return model


def _call_modfreq_of_one_region(refpos2modinfo, args, model, h0=None):
    """Drop-in for the reference's ``_call_modfreq_of_one_region`` (call_mods_freq_bam.py:423-442 and, in aggregate
    mode, :308-420) with the pileup statistics, histograms, window building and the model on the device.

    refpos2modinfo: {refpos: [(ML byte or probability, hap), ...]} as ``_readmods_to_bed_of_one_region`` builds it
    (:457-540); probabilities are mapped back to their ML byte (``_cal_mod_prob`` is a bijection on 0..255).
    Unlike the reference, the model is not rebuilt per region: pass the resident ``AggrAttRNN``.
    Returns the reference's list of ``(refpos, info_all, info_hp1, info_hp2)`` with ``info = (cov, cnt_mod, freq)``
    or None.  h0: optional per-group initial states; default = the reference's stream (one ``torch.randn`` per
    1024-site slice, groups in the order all / hp1 / hp2)."""
    refposes = np.array(sorted(refpos2modinfo.keys()), dtype=np.int64)
    counts = np.array([len(refpos2modinfo[p]) for p in refposes], dtype=np.int64)
    ptr = np.concatenate(([0], np.cumsum(counts))).astype(np.int64)
    ml = np.empty(int(ptr[-1]), dtype=np.uint8)
    hap = np.empty(int(ptr[-1]), dtype=np.uint8)
    k = 0
    for p in refposes:
        for v, h in refpos2modinfo[p]:
            ml[k] = v if isinstance(v, (int, np.integer)) else (0 if v <= 0 else int(round((v - 0.000001) * 256)))
            hap[k] = h if 0 <= h <= 255 else 0
            k += 1
    n_high = model.pileup_begin(refposes, ptr, ml, hap, call_mode=args.call_mode, cov_cf=args.cov_cf,
                                prob_cf=args.prob_cf, no_amb_cov=args.no_amb_cov, no_hap=args.no_hap,
                                only_close=getattr(args, "only_close", False))
    if h0 is None and args.call_mode == "aggregate":
        h0 = [draw_initial_states(nh, model.num_layers, model.hidden_size, model.rnn_cell == "lstm") if nh else None
              for nh in n_high]
    cov, cnt, freq = model.pileup_finish(h0 if h0 is not None else (None, None, None))
    out = []
    for i, p in enumerate(refposes):
        infos = [None if cov[g, i] < 0 else (int(cov[g, i]), float(cnt[g, i]), float(freq[g, i])) for g in range(3)]
        out.append((int(p), infos[0], infos[1], infos[2]))
    return out
