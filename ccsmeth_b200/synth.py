"""Synthetic (batch, 21, feat) call_mods feature batches (BASELINE.json configs 2 and 4).

The geometry mirrors what the reference extractor produces for a CpG site
(reference ccsmeth/extract_features.py:343-363): one 22-mer ``s`` with ``s[10:12] == "CG"``;
forward window ``s[0:21]``, reverse window = reverse complement of ``s[1:22]`` -- so both
strands carry C at index 10 and G at index 11.  Kinetics are standardised log-normal
(heavy right tail like the demo's IPD/PW z-scores), npass ~ U{1..30} per strand tiled over
the window (reference call_modifications.py:104,113).  Layout = the 8 live tensors of the
reference's 16-tensor forward, float32, plus the two explicit h0.
"""
import numpy as np
import torch

SEED = 20261017


def make_batch(n, seq_len=21, seed=SEED, device="cpu", with_h0=True, num_layers=3, hidden=256,
               n_frac=0.001):
    """Returns dict with kmer/kpass/ipd/pw (+ '2' suffix for the reverse strand), each (n, seq_len)
    float32 on `device`, and h0_f/h0_r (2*layers, n, hidden) if with_h0."""
    dev = torch.device(device)
    g = torch.Generator(device=dev)
    g.manual_seed(int(seed))
    L = seq_len
    c = L // 2
    s = torch.randint(0, 4, (n, L + 1), generator=g, device=dev, dtype=torch.int64)
    s[:, c] = 1      # 'C'
    s[:, c + 1] = 2  # 'G'
    kf = s[:, 0:L].clone()
    kr = (3 - s[:, 1:L + 1]).flip(1).clone()
    if n_frac > 0:
        kf[torch.rand((n, L), generator=g, device=dev) < n_frac] = 4
        kr[torch.rand((n, L), generator=g, device=dev) < n_frac] = 4

    def kin():
        x = torch.exp(0.6 * torch.randn((n, L), generator=g, device=dev))
        # standardised log-normal: mean exp(s^2/2), var (exp(s^2)-1) exp(s^2)
        mu = float(np.exp(0.18))
        sd = float(np.sqrt((np.exp(0.36) - 1.0) * np.exp(0.36)))
        return torch.round(((x - mu) / sd) * 1e6) / 1e6

    out = {
        "kmer": kf.float(), "kmer2": kr.float(),
        "ipd": kin(), "pw": kin(), "ipd2": kin(), "pw2": kin(),
    }
    npf = torch.randint(1, 31, (n, 1), generator=g, device=dev).float()
    npr = torch.randint(1, 31, (n, 1), generator=g, device=dev).float()
    out["kpass"] = npf.expand(n, L).contiguous()
    out["kpass2"] = npr.expand(n, L).contiguous()
    if with_h0:
        out["h0_f"] = torch.randn((2 * num_layers, n, hidden), generator=g, device=dev)
        out["h0_r"] = torch.randn((2 * num_layers, n, hidden), generator=g, device=dev)
    return out


def to_forward_args(b):
    """The reference's 16 positional forward tensors (dead slots are (n,) zeros,
    reference call_modifications.py:201-208)."""
    n = b["kmer"].shape[0]
    z = torch.zeros((n,), dtype=torch.float32, device=b["kmer"].device)
    return (b["kmer"], b["kpass"], b["ipd"], z, b["pw"], z, z, z,
            b["kmer2"], b["kpass2"], b["ipd2"], z, b["pw2"], z, z, z)


def make_aggr_batch(n_sites, seq_len=11, bins=20, seed=SEED, hidden=32):
    """Synthetic pileup for the aggregate model (BASELINE.json config 5): coverage ~ U{4..60},
    per-read probs = round(ML/256 + 1e-6, 6) with ML = floor(256*Beta(.3,.3))
    (reference call_mods_freq_bam.py:102-107), 20-bin L2-normalised histogram rounded to 6 dp
    (:221-237), positions = cumsum U{2..200}.  Returns positions (n,), histos (n, bins) float64, h0."""
    rng = np.random.default_rng(seed)
    cov = rng.integers(4, 61, size=n_sites)
    pos = np.cumsum(rng.integers(2, 201, size=n_sites)).astype(np.int64)
    histos = np.zeros((n_sites, bins), dtype=np.float64)
    for i in range(n_sites):
        ml = np.floor(256 * rng.beta(0.3, 0.3, size=cov[i])).clip(0, 255)
        p = np.round(ml / 256.0 + 1e-6, 6)
        h, _ = np.histogram(p, bins=bins, range=(0, 1))
        h = h.astype(np.float64)
        nrm = np.linalg.norm(h)
        histos[i] = np.round(h / nrm, 6) if nrm > 0 else h
    g = torch.Generator().manual_seed(int(seed))
    h0 = torch.randn((2, n_sites, hidden), generator=g)
    return pos, histos, h0
