"""Import the UNMODIFIED reference (read-only /root/reference) -- build-container tool only.

The reference's hot-path modules import ``pysam`` (and transitively
``statsmodels``, ``tabix``, ``pybedtools``) at module top
(reference ccsmeth/utils/process_utils.py:7); none of them is installed and none is
touched by the model forward.  Registering empty stub modules makes
``ccsmeth.models`` / ``ccsmeth.call_modifications`` importable unmodified.

/root/reference does not exist on the GPU box: nothing that runs there may call this
(tests guard on ``available()``).  Used by scripts/gen_golden.py to generate the
committed fixtures under tests/golden/.
"""
import os
import sys
import types

REFERENCE_ROOT = "/root/reference"
V3_CKPT = os.path.join(REFERENCE_ROOT, "models", "model_ccsmeth_5mCpG_call_mods_attbigru2s_b21.v3.ckpt")
AGGR_CKPT = os.path.join(REFERENCE_ROOT, "models", "model_ccsmeth_5mCpG_aggregate_attbigru_b11.v2p.ckpt")
DEMO_BAM = os.path.join(REFERENCE_ROOT, "demo", "hg002.chr20_demo.hifi.bam")

_STUBS = ["pysam", "statsmodels", "statsmodels.robust", "tabix", "pybedtools"]


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "ccsmeth"))


def import_reference():
    """Returns the reference's ``ccsmeth`` package (models, call_modifications importable)."""
    if not available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    for m in _STUBS:
        if m not in sys.modules:
            sys.modules[m] = types.ModuleType(m)
    sys.modules["statsmodels"].robust = sys.modules["statsmodels.robust"]
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
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
