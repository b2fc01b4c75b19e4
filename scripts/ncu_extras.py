"""ncu target for the non-GRU kernels: one device-extraction call chain on the demo BAM (read scan, site list,
window gather, forward, site tags) and one fused aggregate forward."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from ccsmeth_b200 import call_mods as cm
from ccsmeth_b200.bamio import BamReader
from ccsmeth_b200.extract_features import extract_opts, pack_reads
from ccsmeth_b200.models import AggrAttRNN, ModelAttRNN
ck = dict(np.load(os.path.join(ROOT, "tests", "golden", "ckpt_att2s_v3.npz")))
m = ModelAttRNN(21, 3, 2, 0, 256, is_npass=True, model_type="attbigru2s", device=0, precision="bf16")
m.load_state_dict({k: torch.from_numpy(v) for k, v in ck.items()}); m = m.cuda(0).eval()
demo = os.path.join(ROOT, "tests", "golden", "demo", "hg002.chr20_demo.hifi.bam")
args = cm.build_parser().parse_args(["-i", demo, "-m", "x", "-o", "o"])
recs = list(BamReader(demo)) * 8
batch = pack_reads(recs, args)
m.set_h0_mode("device", 1)
for _ in range(2):
    n = m.extract_reads(batch, extract_opts(args, ["CG"]))
    m.reads_forward(want_probs=False)
ca = dict(np.load(os.path.join(ROOT, "tests", "golden", "ckpt_aggr_v2p.npz")))
a = AggrAttRNN(11, 1, 1, 0, 32, binsize=20, model_type="attbigru", device=0)
a.load_state_dict({k: torch.from_numpy(v) for k, v in ca.items()}); a = a.cuda(0).eval()
N = 1 << 20
g = torch.Generator(device="cuda").manual_seed(1)
h = torch.rand((N, 11, 20), generator=g, device="cuda")
o = torch.randint(0, 1200, (N, 11), generator=g, device="cuda").float()
h0 = torch.randn((2, N, 32), generator=g, device="cuda")
for _ in range(2):
    a(o, h, h0=h0)
# region pileup: 2^18 sites, coverage U{4..60}
rng = np.random.default_rng(3)
ns = 1 << 18
cov = rng.integers(4, 61, size=ns)
ptr = np.concatenate(([0], np.cumsum(cov))).astype(np.int64)
ml = rng.integers(0, 256, int(ptr[-1])).astype(np.uint8)
hap = rng.integers(0, 3, int(ptr[-1])).astype(np.uint8)
pos = np.cumsum(rng.integers(2, 201, size=ns)).astype(np.int64)
for _ in range(2):
    a.pileup_begin(pos, ptr, ml, hap, call_mode="aggregate")
    a.pileup_finish()
torch.cuda.synchronize()
print("done", n)
