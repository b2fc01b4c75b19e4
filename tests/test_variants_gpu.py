"""Every GRU layer kernel form that ships in libccsm (CCSM_TC_VARIANT, csrc/tc_path.cu: gru_variant) against the same
golden vectors as the default forms.  The variant is read once per process, so each one runs the fast parity tests of
tests/test_tc_gpu.py in a subprocess.  Reference: models.py:109-150 (ModelAttRNN.forward, attbigru2s)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# (layer 0, layers >= 1): d / e = one direction per CTA (4 / 8 epilogue warps), k / l / m = both directions of a row tile
# interleaved (4 / 8 / 16 epilogue warps), h = two CTAs per SM, f = two row tiles per CTA, g = CTA pair (cta_group::2),
# i / j = fp16c8 with on-chip operand conversion, 3 = round 1's CTA-pair kernel (not in fp16c8), n = CTA pair with tensor-map
# loads (the default for layers >= 1), o / p = the same with both directions interleaved (per-recurrence buffers / IL order)
VARIANTS = ["dd", "ee", "kk", "ll", "md", "hf", "gd", "ij", "c3", "ln", "oo", "pn", "nn"]


@pytest.mark.parametrize("variant", VARIANTS)
def test_variant_matches_the_golden_vectors(variant):
    sel = "test_tc_matches_reference_synth or test_tc_multi_tile_and_chunks"
    if "3" in variant or "c" in variant:
        sel = "(%s) and not fp16c8" % sel  # round-1 forms predate the fp16c8 images
    env = dict(os.environ, CCSM_TC_VARIANT=variant)
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_tc_gpu.py"), "-x", "-q", "-m", "gpu",
                        "-k", sel, "-p", "no:cacheprovider"], env=env, cwd=ROOT, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert " passed" in r.stdout
