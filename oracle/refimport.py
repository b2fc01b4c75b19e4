"""Import the UNMODIFIED reference -- from the read-only /root/reference in the build container, or from the
byte-identical staged copy under oracle/_ref (oracle/stage_ref.py; git-ignored, travels to the GPU box) -- test and
measurement infrastructure only.

The reference's hot-path modules import ``pysam`` (and transitively
``statsmodels``, ``tabix``, ``pybedtools``) at module top
(reference ccsmeth/utils/process_utils.py:7); none of them is installed and none is
touched by the model forward.  Registering empty stub modules makes
``ccsmeth.models`` / ``ccsmeth.call_modifications`` importable unmodified.

/root/reference does not exist on the GPU box: there only the staged package is importable (checkpoints and the demo
BAM come from tests/golden/).  Used by scripts/gen_golden.py to generate the committed fixtures under tests/golden/
and by oracle/ref_cpu_bench.py (the CPU baseline arm of bench.py).
"""
import os
import sys
import types

REFERENCE_ROOT = "/root/reference"
V3_CKPT = os.path.join(REFERENCE_ROOT, "models", "model_ccsmeth_5mCpG_call_mods_attbigru2s_b21.v3.ckpt")
AGGR_CKPT = os.path.join(REFERENCE_ROOT, "models", "model_ccsmeth_5mCpG_aggregate_attbigru_b11.v2p.ckpt")
DEMO_BAM = os.path.join(REFERENCE_ROOT, "demo", "hg002.chr20_demo.hifi.bam")

STAGED_ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")

_STUBS = ["pysam", "statsmodels", "statsmodels.robust", "tabix", "pybedtools"]


def available():
    """The full reference tree (checkpoints, demo data) is present: build container only."""
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "ccsmeth"))


def package_root():
    """Directory holding an importable, unmodified ``ccsmeth`` package, or None."""
    if available():
        return REFERENCE_ROOT
    if os.path.isfile(os.path.join(STAGED_ROOT, "ccsmeth", "models.py")):
        return STAGED_ROOT
    return None


def import_reference():
    """Returns the reference's ``ccsmeth`` package (models, call_modifications importable)."""
    root = package_root()
    if root is None:
        raise RuntimeError("reference package neither at %s nor staged under %s" % (REFERENCE_ROOT, STAGED_ROOT))
    for m in _STUBS:
        if m not in sys.modules:
            sys.modules[m] = types.ModuleType(m)
    sys.modules["statsmodels"].robust = sys.modules["statsmodels.robust"]
    if root not in sys.path:
        sys.path.insert(0, root)
    import ccsmeth  # noqa: F401  (the reference package, not ours: ours is ccsmeth_b200)
    import ccsmeth.models  # noqa: F401
    return ccsmeth


def load_ref_att2s(ckpt=V3_CKPT):
    """The reference ModelAttRNN(attbigru2s) with the shipped v3 weights, eval mode, CPU
    (construction + load mirror reference call_modifications.py:316-369)."""
    import torch
    ref = import_reference()
    m = ref.models.ModelAttRNN(21, 3, 2, 0, 256, is_npass=True, is_sn=False, is_map=False, is_stds=False,
                               model_type="attbigru2s", device=0)
    sd = torch.load(ckpt, map_location="cpu")
    d = m.state_dict()
    d.update(sd)
    m.load_state_dict(d)
    m.eval()
    return m


def load_ref_aggr(ckpt=AGGR_CKPT):
    """The reference AggrAttRNN with the shipped aggregate weights (call_mods_freq_bam.py:317-342)."""
    import torch
    from collections import OrderedDict
    ref = import_reference()
    m = ref.models.AggrAttRNN(11, 1, 1, 0, 32, binsize=20, model_type="attbigru", device="cpu")
    sd = torch.load(ckpt, map_location="cpu")
    sd = OrderedDict((k[7:] if k.startswith("module.") else k, v) for k, v in sd.items())
    m.load_state_dict(sd)
    m.eval()
    return m


class fixed_h0:
    """Context manager: make the reference's ``init_hidden`` return the given tensors in call order
    (the reference draws h0 with torch.randn per strand per batch, models.py:77-87)."""

    def __init__(self, model, h0_list):
        self.model = model
        self.h0 = list(h0_list)

    def __enter__(self):
        self._orig = self.model.init_hidden
        it = iter(self.h0)
        self.model.init_hidden = lambda *a, **k: next(it)
        return self

    def __exit__(self, *exc):
        self.model.init_hidden = self._orig
        return False
