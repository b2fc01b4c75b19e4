"""Host side of the call_mods batch loop (the reference's L2 layer, ccsmeth/call_modifications.py).

Mirrors, with the same names, argument meaning and outputs:
  * ``_batch_feature_list2s``   reference call_modifications.py:73-123   (the dataloader tensor layout)
  * ``_call_mods2s``            reference call_modifications.py:170-227  (the batch loop)
  * ``load_model``              reference call_modifications.py:313-369  (model lifecycle in _call_mods_q)

What changes is *how* the loop runs: instead of 16 synchronous ``FloatTensor`` H2D copies and one
model call per 512 sites, the whole hole-batch (~5 k sites) is stacked once and handed to libccsm in
ONE call (host entry: pipelined H2D / kernels / D2H).  The h0 stream is still drawn chunk-ordered
exactly as the reference draws it -- for each ``batch_size`` slice, ``randn(6, n_c, 256)`` for strand 1
then strand 2 from the CPU default generator (models.py:77-87,125-130) -- so ``--batch_size`` keeps its
reference meaning and results are comparable site by site.
"""
from collections import OrderedDict

import numpy as np
import torch

from .models import ModelAttRNN, ModelAttRNN2, ModelTransEnc
from .utils.process_utils import base2code_dna, str2bool

_CODE_LUT = np.full(256, 4, dtype=np.int64)
for _b, _c in base2code_dna.items():
    _CODE_LUT[ord(_b)] = _c


def _kmer_codes(kmer_seq):
    return _CODE_LUT[np.frombuffer(kmer_seq.encode("ascii"), dtype=np.uint8)]


def _batch_feature_list2s(feature_list):
    """list of 22-field per-site rows (reference extract_features.py:400-405) -> the 18-tuple
    ``(sampleinfo, fkmers, fpasss, fipdms, fipdsds, fpwms, fpwsds, fsns, fmaps, rkmers, ..., labels)``.

    Same field order and values as the reference; the per-site ``np.ndarray`` rows are kept in lists so the
    result also feeds the reference's own ``_call_mods2s`` unchanged.  Unused feature slots are the scalar 0
    (reference :106-110)."""
    sampleinfo = []
    cols = [[] for _ in range(16)]
    labels = []
    for fl in feature_list:
        (chrom, abs_loc, strand, holeid, loc,
         kmer_seq, kmer_pass, kmer_ipdm, kmer_ipds, kmer_pwm, kmer_pws, kmer_sn, kmer_map,
         kmer_seq2, kmer_pass2, kmer_ipdm2, kmer_ipds2, kmer_pwm2, kmer_pws2, kmer_sn2, kmer_map2,
         label) = fl
        sampleinfo.append("\t".join(map(str, (chrom, abs_loc, strand, holeid, loc))))
        for base, (seq, npass, ipdm, ipds, pwm, pws, sn, mp) in (
                (0, (kmer_seq, kmer_pass, kmer_ipdm, kmer_ipds, kmer_pwm, kmer_pws, kmer_sn, kmer_map)),
                (8, (kmer_seq2, kmer_pass2, kmer_ipdm2, kmer_ipds2, kmer_pwm2, kmer_pws2, kmer_sn2, kmer_map2))):
            cols[base + 0].append(_kmer_codes(seq))
            cols[base + 1].append(np.full(len(seq), npass))
            cols[base + 2].append(np.array(ipdm, dtype=float))
            cols[base + 3].append(np.array(ipds, dtype=float) if type(ipds) is not str else 0)
            cols[base + 4].append(np.array(pwm, dtype=float))
            cols[base + 5].append(np.array(pws, dtype=float) if type(pws) is not str else 0)
            cols[base + 6].append(np.array(sn, dtype=float) if type(sn) is not str else 0)
            cols[base + 7].append(np.array(mp, dtype=float) if type(mp) is not str else 0)
        labels.append(label)
    return (sampleinfo, *cols, labels)


def draw_h0_stream(n, batch_size, num_layers, hidden, generator=None, out=None, offset=0):
    """The reference's h0 stream for a hole-batch of n sites processed in ``batch_size`` slices:
    per slice, strand-1 draw then strand-2 draw.  Returns two (2*layers, n, hidden) float32 tensors -- or, with
    ``out`` = a pair of (2*layers, N, hidden) float32 numpy arrays, writes sites offset..offset+n of them (the
    caller assembles several hole-batches without further copies).  Each slice is drawn as one contiguous
    ``torch.randn`` (the reference's call) and placed with a plain memory copy."""
    if out is None:
        dst = (np.empty((2 * num_layers, n, hidden), dtype=np.float32),
               np.empty((2 * num_layers, n, hidden), dtype=np.float32))
        offset = 0
    else:
        dst = out
    for s in range(0, n, batch_size):
        e = min(n, s + batch_size)
        dst[0][:, offset + s:offset + e] = torch.randn(2 * num_layers, e - s, hidden, generator=generator).numpy()
        dst[1][:, offset + s:offset + e] = torch.randn(2 * num_layers, e - s, hidden, generator=generator).numpy()
    if out is None:
        return torch.from_numpy(dst[0]), torch.from_numpy(dst[1])
    return out


def draw_h0_stream_batches(counts, batch_size, num_layers, hidden, generator=None):
    """`draw_h0_stream` for consecutive hole-batches of `counts` sites each (the reference's stream restarts its
    ``batch_size`` slicing at every hole-batch), assembled in place.  -> two (2*layers, sum(counts), hidden) tensors."""
    total = int(sum(int(c) for c in counts))
    out = (np.empty((2 * num_layers, total, hidden), dtype=np.float32),
           np.empty((2 * num_layers, total, hidden), dtype=np.float32))
    off = 0
    for c in counts:
        draw_h0_stream(int(c), batch_size, num_layers, hidden, generator, out=out, offset=off)
        off += int(c)
    return torch.from_numpy(out[0]), torch.from_numpy(out[1])


def _stack(col, n):
    return np.ascontiguousarray(np.asarray(col, dtype=np.float32).reshape(n, -1))


def _call_mods2s(features_batch, model, batch_size, device=0, h0=None):
    """Batch loop replacement (reference call_modifications.py:170-227).

    features_batch: the 18-tuple of ``_batch_feature_list2s``.  Returns ``(pred_info, batch_num)`` with
    ``pred_info = [(holeid, loc, prob_1_norm)]`` where ``prob_1_norm = round(p1 / (p0 + p1), 6)`` and
    ``batch_num`` = number of ``batch_size`` slices the reference would have run.
    ``h0``: optional explicit (h0_strand1, h0_strand2); default = the reference's random stream."""
    (sampleinfo, fkmers, fpasss, fipdms, fipdsds, fpwms, fpwsds, fsns, fmaps,
     rkmers, rpasss, ripdms, ripdsds, rpwms, rpwsds, rsns, rmaps, _) = features_batch
    n = len(sampleinfo)
    if n == 0:
        return [], 0
    L = model.seq_len
    feats = {}
    for sfx, (km, ps, im, isd, pm, psd, sn, mp) in (("", (fkmers, fpasss, fipdms, fipdsds, fpwms, fpwsds, fsns, fmaps)),
                                                    ("2", (rkmers, rpasss, ripdms, ripdsds, rpwms, rpwsds, rsns, rmaps))):
        feats["kmer" + sfx] = _stack(km, n)
        feats["kpass" + sfx] = _stack(ps, n)
        feats["ipd" + sfx] = _stack(im, n)
        feats["pw" + sfx] = _stack(pm, n)
        if model.is_stds:
            feats["ipd_sd" + sfx] = _stack(isd, n)
            feats["pw_sd" + sfx] = _stack(psd, n)
        if model.is_sn:
            feats["sns" + sfx] = _stack(sn, n)
        if model.is_map:
            feats["maps" + sfx] = _stack(mp, n)
    if h0 is None:
        if model._stream_h0():
            model.set_h0_batching([n], batch_size)   # the library draws the reference's stream on the device
        elif model._host_h0():
            h0 = draw_h0_stream(n, batch_size, model.num_layers, model.hidden_size)
    _, probs = model.forward_host(feats, h0=h0)
    p = probs.numpy()
    prob_1_norm = np.round(p[:, 1] / (p[:, 0] + p[:, 1]), 6)  # float32, like round(np.float32, 6) (:223)
    pred_info = []
    for idx in range(n):
        f = sampleinfo[idx].split("\t")
        pred_info.append((f[3], int(f[4]), prob_1_norm[idx]))
    batch_num = (n + batch_size - 1) // batch_size
    return pred_info, batch_num


def load_model(model_path, args, device=0, precision=None):
    """Model lifecycle of the reference's model worker (call_modifications.py:313-369): construct from the
    CLI args, ``torch.load`` the checkpoint on CPU, ``state_dict().update(); load_state_dict``, with the
    ``module.``-prefix fallback, then ``.cuda(device)`` and ``.eval()``."""
    if args.model_type not in {"attbigru2s", "attbilstm2s", "attbigru2s2", "attbilstm2s2", "transencoder2s"}:
        raise ValueError("--model_type not right!")
    if args.model_type == "transencoder2s":  # reference call_modifications.py:333-338
        model = ModelTransEnc(args.seq_len, args.layer_trans, args.class_num, args.dropout_rate, args.d_model, args.nhead,
                              args.dim_ff, is_npass=str2bool(args.is_npass), is_sn=str2bool(args.is_sn),
                              is_map=str2bool(args.is_map), is_stds=str2bool(args.is_stds), model_type=args.model_type,
                              device=device)
        return _load_and_place(model, model_path, device)
    cls = ModelAttRNN2 if args.model_type.endswith("2s2") else ModelAttRNN
    model = cls(args.seq_len, args.layer_rnn, args.class_num, args.dropout_rate, args.hid_rnn,
                        is_sn=str2bool(args.is_sn), is_map=str2bool(args.is_map), is_stds=str2bool(args.is_stds),
                        is_npass=str2bool(args.is_npass), model_type=args.model_type, device=device,
                        precision=precision)
    return _load_and_place(model, model_path, device)


def _load_and_place(model, model_path, device):
    para_dict = torch.load(model_path, map_location=torch.device('cpu'))
    try:
        model_dict = model.state_dict()
        model_dict.update(para_dict)
        model.load_state_dict(model_dict)
    except RuntimeError:
        new = OrderedDict((k[7:], v) for k, v in para_dict.items())
        model.load_state_dict(new)
    model = model.cuda(device)
    model.eval()
    return model
