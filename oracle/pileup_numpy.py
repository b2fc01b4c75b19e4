"""ORACLE (test infrastructure only): numpy restatement of the reference's per-region frequency caller.

Follows reference ccsmeth/call_mods_freq_bam.py:
  * ``_cal_mod_prob``                      :102-107
  * ``_cal_modfreq_in_count_mode``         :200-217
  * ``_get_normalized_histo``              :221-237
  * ``_call_modfreq_of_one_region`` (count :423-442, aggregate :308-420) -- group split all / hp1 / hp2, low / high
    coverage split, ``round(cov * modprob, 2)``.
The model forward is ``oracle.aggr_numpy`` (windows + AggrAttRNN).  Pinned against the reference's own output in
tests/golden/pileup_region.npz (scripts/gen_golden.py gen_pileup).
"""
import numpy as np

from . import aggr_numpy


def cal_mod_prob(ml_value):
    return round(ml_value / float(256) + 0.000001, 6) if ml_value > 0 else 0


def count_mode(modprobs, prob_cf=0, no_amb_cov=False):
    cnt_all_filtered, cnt_mod = 0, 0
    for p in modprobs:
        if abs(p - (1 - p)) < prob_cf:
            continue
        cnt_all_filtered += 1
        if p > 0.5:
            cnt_mod += 1
    modfreq = cnt_mod / float(cnt_all_filtered) if cnt_all_filtered > 0 else 0.
    if no_amb_cov:
        return cnt_all_filtered, cnt_mod, modfreq
    if cnt_all_filtered != len(modprobs):
        cnt_mod = np.round(len(modprobs) * modfreq, 2)
    return len(modprobs), cnt_mod, modfreq


def normalized_histo(probs, binsize=20):
    hist = np.histogram(probs, bins=binsize, range=[0, 1])[0]
    return np.round(hist / np.linalg.norm(hist), 6)


def call_region(pos, ptr, ml, hap, sd, call_mode="aggregate", cov_cf=4, prob_cf=0.0, no_amb_cov=False, no_hap=False,
                h0=(None, None, None), seq_len=11, only_close=False, num_layers=1):
    """-> (3, n, 3) array of (cov, cnt_mod, freq) per group (all, hp1, hp2), NaN where the reference returns None.
    h0[g]: (2*layers, n_high_g, hidden) initial states of group g's high-coverage sites (zeros if None); an (h0, c0)
    pair selects the LSTM cell (model_type="attbilstm")."""
    n = len(pos)
    out = np.full((3, n, 3), np.nan)
    for g in range(3):
        if g > 0 and no_hap:
            continue
        hi_idx, hi_hist, hi_cov = [], [], []
        for i in range(n):
            sel = slice(ptr[i], ptr[i + 1])
            probs = [cal_mod_prob(int(v)) for v, h in zip(ml[sel], hap[sel]) if g == 0 or h == g]
            if not probs:
                continue
            if call_mode == "aggregate" and len(probs) >= cov_cf:
                hi_idx.append(i)
                hi_hist.append(normalized_histo(probs))
                hi_cov.append(len(probs))
            else:
                out[g, i] = count_mode(probs, prob_cf, no_amb_cov)
        if hi_idx:
            pm, hm = aggr_numpy.build_windows(pos[hi_idx], hi_hist, seq_len, only_close)
            hidden = next(np.asarray(w).shape[1] for k, w in sd.items() if k.endswith("rnn.weight_hh_l0"))
            hh = h0[g] if h0[g] is not None else np.zeros((2 * num_layers, len(hi_idx), hidden), dtype=np.float32)
            raw = aggr_numpy.forward(sd, pm.astype(np.float32), hm.astype(np.float32), hh, num_layers, dtype=np.float32)
            p = aggr_numpy.postprocess(raw)[:, 0]
            for k, i in enumerate(hi_idx):
                out[g, i] = (hi_cov[k], round(hi_cov[k] * p[k], 2), p[k])
    return out
