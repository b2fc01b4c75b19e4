#!/usr/bin/env python
"""Host-entry throughput of the forward per h0 mode (device Philox / the reference's torch.randn stream on the device /
the same stream drawn by torch on the host).  One JSON line."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from ccsmeth_b200 import synth
from ccsmeth_b200.models import ModelAttRNN
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 19
prec = sys.argv[2] if len(sys.argv) > 2 else "fp16c8"
ck = dict(np.load(os.path.join(ROOT, "tests", "golden", "ckpt_att2s_v3.npz")))
m = ModelAttRNN(21, 3, 2, 0, 256, is_npass=True, model_type="attbigru2s", device=0, precision=prec)
m.load_state_dict({k: torch.from_numpy(v) for k, v in ck.items()}); m = m.cuda(0).eval()
b = synth.make_batch(n, with_h0=False)
feats = {k: b[k].pin_memory() for k in ("kmer", "kpass", "ipd", "pw", "kmer2", "kpass2", "ipd2", "pw2")}
res = {"sites": n, "precision": prec}
for mode in ("device", "reference") + (("reference_host",) if n <= (1 << 17) else ()):
    m.set_h0_mode(mode, seed=7)
    torch.manual_seed(7)
    m.forward_host(feats)
    t0 = time.perf_counter()
    for _ in range(3):
        m.forward_host(feats)
    res[mode] = 3 * n / (time.perf_counter() - t0)
print(json.dumps(res))
