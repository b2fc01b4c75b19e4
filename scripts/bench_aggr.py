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
print(json.dumps({"workload": "call_freqb aggregate attbigru_b11.v2p forward, synthetic (n,11,21) windows, 1xB200, fp32",
                  "sites": n, "ms_per_step": ms, "sites_per_s": n / (ms * 1e-3), "gpu_launches_per_step": (_lib.kernel_launches() - l0) / K,
                  "hbm_roofline_frac_materialised_windows": n * 1184 / (ms * 1e-3) / 6550.7e9,
                  "max_abs_diff_vs_cpu_port": d, "cpu_port_sites_per_s": cpu, "cpu_threads": torch.get_num_threads()}))
