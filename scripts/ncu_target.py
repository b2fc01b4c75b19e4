"""Small driver for ncu captures: a few forwards over one library chunk (75,776 sites) in one precision."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from ccsmeth_b200 import synth
from ccsmeth_b200.models import ModelAttRNN
prec = sys.argv[1] if len(sys.argv) > 1 else "bf16"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 75776
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
ck = dict(np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "ckpt_att2s_v3.npz")))
m = ModelAttRNN(21, 3, 2, 0, 256, is_npass=True, model_type="attbigru2s", device=0, precision=prec)
m.load_state_dict({k: torch.from_numpy(v) for k, v in ck.items()}); m = m.cuda(0).eval()
b = synth.make_batch(n, device="cuda:0")
args = synth.to_forward_args(b)
for _ in range(reps):
    m(*args, h0=(b["h0_f"], b["h0_r"]))
torch.cuda.synchronize()
print("done", prec, n)
