"""Materialise the shipped checkpoints (tests/golden/ckpt_*.npz) as torch .ckpt files (the reference's format:
zip-format torch.save of the state_dict OrderedDict).  usage: make_ckpt.py <out_dir>"""
import os, sys
from collections import OrderedDict
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = sys.argv[1] if len(sys.argv) > 1 else "."
os.makedirs(out, exist_ok=True)
for src, dst in (("ckpt_att2s_v3.npz", "model_ccsmeth_5mCpG_call_mods_attbigru2s_b21.v3.ckpt"),
                 ("ckpt_aggr_v2p.npz", "model_ccsmeth_5mCpG_aggregate_attbigru_b11.v2p.ckpt")):
    z = np.load(os.path.join(ROOT, "tests", "golden", src))
    torch.save(OrderedDict((k, torch.from_numpy(z[k])) for k in z.files), os.path.join(out, dst))
    print("wrote", os.path.join(out, dst))
