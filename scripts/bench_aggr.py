#!/usr/bin/env python
"""Config 5: aggregate-mode model (attbigru_b11.v2p) on synthetic pileup windows, 1 GPU.  One JSON line."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from ccsmeth_b200.models import AggrAttRNN
from ccsmeth_b200 import _lib
from oracle import torch_port

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
ck = dict(np.load(os.path.join(ROOT, "tests", "golden", "ckpt_aggr_v2p.npz")))
m = AggrAttRNN(11, 1, 1, 0, 32, binsize=20, model_type="attbigru", device=0)
m.load_state_dict({k: torch.from_numpy(v) for k, v in ck.items()})
m = m.cuda(0).eval()
g = torch.Generator(device="cuda").manual_seed(20261017)
histos = torch.rand((n, 11, 20), generator=g, device="cuda")
histos = torch.round(histos / histos.norm(dim=2, keepdim=True) * 1e6) / 1e6
offsets = torch.randint(0, 1200, (n, 11), generator=g, device="cuda").float()
h0 = torch.randn((2, n, 32), generator=g, device="cuda")
for _ in range(2):
    out = m(offsets, histos, h0=h0)
torch.cuda.synchronize()
l0 = _lib.kernel_launches()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
K = 3
for _ in range(K):
    out = m(offsets, histos, h0=h0)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / K
# parity on a slice vs the oracle's torch port, and the port's CPU throughput
port = torch_port.load_numpy_state(torch_port.AggrPort(), ck)
P = 4096
with torch.no_grad():
    ref = port(offsets[:P].cpu(), histos[:P].cpu(), h0[:, :P].cpu().contiguous())
d = float((out[:P].cpu() - ref).abs().max())
t0 = time.perf_counter()
for _ in range(10):
    port(offsets[:1024].cpu(), histos[:1024].cpu())
cpu = 10 * 1024 / (time.perf_counter() - t0)
# ---- config 5 from the pileup: ML bytes per site -> histograms -> in-kernel windows -> model -> frequencies
rng = np.random.default_rng(20261017)
cov = rng.integers(4, 61, size=n)
ptr = np.concatenate(([0], np.cumsum(cov))).astype(np.int64)
mlb = np.floor(256 * rng.beta(0.3, 0.3, size=int(ptr[-1]))).clip(0, 255).astype(np.uint8)
pos = np.cumsum(rng.integers(2, 201, size=n)).astype(np.int64)
h0p = torch.randn(2, n, 32).pin_memory()
m.pileup_begin(pos, ptr, mlb, None, call_mode="aggregate", no_hap=True)
m.pileup_finish((h0p, None, None))
t0 = time.perf_counter()
for _ in range(K):
    m.pileup_begin(pos, ptr, mlb, None, call_mode="aggregate", no_hap=True)
    covd, cntd, freqd = m.pileup_finish((h0p, None, None))
pile_s = (time.perf_counter() - t0) / K
from oracle import pileup_numpy
Q = 3000
refp = pileup_numpy.call_region(pos[:Q], ptr[:Q + 1], mlb, np.zeros(len(mlb), np.uint8), ck, no_hap=True,
                                h0=(h0p[:, :Q].numpy(), None, None))
dp = float(np.abs(freqd[0][:Q - 6] - refp[0][:Q - 6, 2]).max())  # the last 5 sites of the slice see other neighbours
pile = {"workload": "synthetic pileup: %d sites, %d calls (coverage U{4..60}) -> frequencies, host CSR in, host results out"
                    % (n, int(ptr[-1])), "seconds": pile_s, "sites_per_s": n / pile_s,
        "h2d_bytes": int(ptr[-1]) + 16 * n + 256 * n, "max_abs_dfreq_vs_oracle_first_%d" % Q: dp}
print(json.dumps({"pileup_end_to_end": pile, "workload": "call_freqb aggregate attbigru_b11.v2p forward, synthetic (n,11,21) windows, 1xB200, fp32",
                  "sites": n, "ms_per_step": ms, "sites_per_s": n / (ms * 1e-3), "gpu_launches_per_step": (_lib.kernel_launches() - l0) / K,
                  "hbm_roofline_frac_materialised_windows": n * 1184 / (ms * 1e-3) / 6550.7e9,
                  "max_abs_diff_vs_cpu_port": d, "cpu_port_sites_per_s": cpu, "cpu_threads": torch.get_num_threads()}))
