"""GPU debugging aid: per-layer error of the tensor-core path vs the numpy oracle, all precisions."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import att2s_numpy
from tests.test_tc_gpu import run, layer_out
from tests.test_parity_gpu import FEATS
from tests.conftest import load_npz
from ccsmeth_b200.models import ModelAttRNN

ck = load_npz("ckpt_att2s_v3.npz"); gs = load_npz("att2s_synth.npz")
m = ModelAttRNN(21, 3, 2, 0, 256, is_npass=True, model_type="attbigru2s", device=0, precision="fp32")
m.load_state_dict({k: torch.from_numpy(v) for k, v in ck.items()}); m = m.cuda(0).eval()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
g = {k: (v[:, :n] if k.startswith("h0") else v[:n]) for k, v in gs.items()}
_, _, it = att2s_numpy.forward(ck, *[g[k] for k in FEATS], g["h0_f"], g["h0_r"], return_internals=True)
for prec in ("fp16x3", "bf16x3", "fp16", "bf16"):
    m.set_precision(prec)
    t0 = time.time()
    _, probs = run(m, g)
    print(prec, "forward ok %.2fs" % (time.time() - t0), "max|dprob| %.3e" % np.abs(probs - g["probs"]).max(), flush=True)
    for l in range(3):
        out = layer_out(m, l, n)
        errs = [np.abs(out[:, s] - it["layers%d" % s][l]) for s in range(2)]
        e = np.maximum(errs[0], errs[1])
        print("  layer", l, "max err %.3e" % e.max(), " fwd-half %.3e rev-half %.3e" % (e[..., :256].max(), e[..., 256:].max()),
              " t0 %.2e t20 %.2e" % (e[:, 0].max(), e[:, 20].max()), flush=True)
