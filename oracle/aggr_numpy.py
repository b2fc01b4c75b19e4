"""numpy restatement of the aggregate-mode model -- TEST ORACLE, not product code.

Follows:
  * ``AggrAttRNN.forward``             reference ccsmeth/models.py:673-694
  * ``_cal_modfreq_in_aggregate_mode`` reference ccsmeth/call_mods_freq_bam.py:265-305
    (window construction + clip/round of the regression output).

Pinned against the reference through tests/golden/aggr_*.npz.
"""
import numpy as np

from .att2s_numpy import bigru_stack, bilstm_stack, attention


def forward(sd, offsets, histos, h0, num_layers=1, dtype=np.float64):
    """offsets (n, L), histos (n, L, B), h0 (2*layers, n, H) -> out (n, 1) raw regression (no softmax).

    x = cat(histos, offsets[..., None])   (models.py:675-677).  h0 given as an (h0, c0) pair selects the LSTM cell of
    model_type="attbilstm" (models.py:640-643, 684).
    """
    sd = {k[7:] if k.startswith("module.") else k: np.asarray(v, dtype=dtype) for k, v in sd.items()}
    x = np.concatenate([np.asarray(histos, dtype=dtype), np.asarray(offsets, dtype=dtype)[:, :, None]], axis=2)
    if isinstance(h0, (tuple, list)):
        out, h_n = bilstm_stack(x, np.asarray(h0[0], dtype=dtype), np.asarray(h0[1], dtype=dtype), sd, num_layers)
    else:
        out, h_n = bigru_stack(x, np.asarray(h0, dtype=dtype), sd, num_layers)
    q = np.concatenate([h_n[2 * (num_layers - 1)], h_n[2 * (num_layers - 1) + 1]], axis=1)
    ctx, _ = attention(q, out, sd)
    return ctx @ sd["fc1.weight"].T + sd["fc1.bias"]


def build_windows(refposes, refposes_histos, seq_len=11, only_close=False):
    """Sliding windows over neighbouring CpG sites (call_mods_freq_bam.py:272-290).

    Returns pos_mat (n, seq_len) |pos_j - pos_center| and histos_mat (n, seq_len, B).
    """
    from numpy.lib.stride_tricks import sliding_window_view
    refposes = np.asarray(refposes)
    pad = seq_len // 2
    hm = np.pad(np.stack(refposes_histos), pad_width=((pad, pad), (0, 0)), mode="constant", constant_values=0)
    hm = np.swapaxes(sliding_window_view(hm, seq_len, axis=0), 1, 2)
    if only_close:
        pm = np.pad(refposes, pad_width=(pad + 1, pad), mode="constant",
                    constant_values=(refposes[0] - 1000, refposes[-1] + 1000))
        pm = (np.diff(pm) == 2).astype(int)
        return sliding_window_view(pm, seq_len), hm
    pm = np.pad(refposes, pad_width=(pad, pad), mode="constant",
                constant_values=(refposes[0] - 1000, refposes[-1] + 1000))
    pm = sliding_window_view(pm, seq_len)
    centre = np.repeat(refposes, seq_len).reshape((-1, seq_len))
    pm = np.absolute(np.subtract(pm, centre))
    return pm, hm


def postprocess(out):
    """np.round(np.clip(out, 0, 1), 6)  (call_mods_freq_bam.py:302)."""
    return np.round(np.clip(np.asarray(out, dtype=np.float32), 0, 1), 6)
