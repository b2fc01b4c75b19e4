#!/usr/bin/env python
"""Device feature extraction kernels against the HBM roofline.

Synthetic batch: R reads x 15 kb, BAM-style 4-bit sequence + four uint8 kinetics arrays per read (4.5 B/base),
~8 % CpG.  Reports, from the library's CUDA events on the launching stream:
  read_scan      per-read statistics + motif scan + ordered site list: algorithmic bytes = 4.5 B/base read once
                 (+ 12 B/site written)
  window_gather  the 16-tensor feature layout: 2 strands x 21 x (2 kinetics bytes + base) read ~ 105 B/site,
                 8 x 21 x 4 B = 672 B/site written
Prints one JSON line."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from ccsmeth_b200.extract_features import READ_DTYPE, READ_SEQ_4BIT, ReadBatch
from ccsmeth_b200.models import ModelAttRNN

R = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
LEN = 15000
rng = np.random.default_rng(1)
codes = rng.integers(0, 4, (R, LEN), dtype=np.uint8)
cg = rng.random((R, LEN - 1)) < 0.08
idx = np.nonzero(cg)
codes[idx[0], idx[1]] = 1
codes[idx[0], idx[1] + 1] = 2
nib = np.array([1, 2, 4, 8], dtype=np.uint8)[codes]
packed = (nib[:, 0::2] << 4) | nib[:, 1::2]
per_read = LEN // 2 + 4 * LEN
blob = np.empty((R, per_read), dtype=np.uint8)
blob[:, :LEN // 2] = packed
blob[:, LEN // 2:] = rng.integers(0, 256, (R, 4 * LEN), dtype=np.uint8)
descs = np.zeros(R, dtype=READ_DTYPE)
base = np.arange(R, dtype=np.int64) * per_read
descs["seq_off"] = base
for k, name in enumerate(("fi_off", "ri_off", "fp_off", "rp_off")):
    descs[name] = base + LEN // 2 + k * LEN
descs["len"], descs["fn"], descs["rn"], descs["flags"], descs["win_hi"] = LEN, 7, 9, READ_SEQ_4BIT, LEN
batch = ReadBatch(blob.reshape(-1), descs, list(range(R)))

ck = dict(np.load(os.path.join(ROOT, "tests", "golden", "ckpt_att2s_v3.npz")))
m = ModelAttRNN(21, 3, 2, 0, 256, is_npass=True, model_type="attbigru2s", device=0, precision="bf16")
m.load_state_dict({k: torch.from_numpy(v) for k, v in ck.items()})
m = m.cuda(0).eval()
opts = {"mod_loc": 0, "norm": 0, "decode": 1, "motifs": ["CG"]}
n = m.extract_reads(batch, opts)
m.reads_features(0, min(n, 1 << 20))
torch.cuda.synchronize()
m.profile(True)
m.profile_read()
reps = 5
t0 = time.perf_counter()
for _ in range(reps):
    n = m.extract_reads(batch, opts)
host_s = (time.perf_counter() - t0) / reps
for _ in range(reps):
    for s0 in range(0, n, 1 << 20):
        f = m.reads_features(s0, min(1 << 20, n - s0))
torch.cuda.synchronize()
prof = m.profile_read()
m.profile(False)
peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6550.7}
bases = R * LEN
scan_ms = prof["read_scan"][0] / reps
gath_ms = prof["window_gather"][0] / reps
scan_bytes = bases * 4.5 + n * 12
gath_bytes = n * (2 * 21 * 2.5 + 8 * 21 * 4)
out = {"workload": "%d reads x %d bases (%.0f MB of sequence + kinetics), %d CpG sites" % (R, LEN, blob.size / 1e6, n),
       "read_scan": {"ms": scan_ms, "algorithmic_GB": scan_bytes / 1e9, "GBps": scan_bytes / scan_ms / 1e6,
                     "hbm_frac": scan_bytes / scan_ms / 1e6 / peaks["hbm_gbs"], "bases_per_s": bases / scan_ms * 1e3},
       "window_gather": {"ms": gath_ms, "algorithmic_GB": gath_bytes / 1e9, "GBps": gath_bytes / gath_ms / 1e6,
                         "hbm_frac": gath_bytes / gath_ms / 1e6 / peaks["hbm_gbs"], "sites_per_s": n / gath_ms * 1e3},
       "extract_call_host_ms": host_s * 1e3, "hbm_peak_GBps": peaks["hbm_gbs"],
       "note": "extract_call_host_ms includes the pageable H2D copy of the batch and one synchronisation"}
print(json.dumps(out))
