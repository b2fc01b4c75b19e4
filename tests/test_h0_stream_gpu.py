"""The reference's h0 stream (torch.randn on the CPU generator, models.py:77-87) reproduced on the device."""
import ctypes

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("seed,skip,n", [(1234, 0, 1 << 20), (0, 160, 4096), (20261017, 624 * 16, 6 * 37 * 256),
                                         # long draws are cut into 2^21-word sub-streams by polynomial jump-ahead
                                         (7, 160, 1 << 26), (3, 624 * 16 * 3 + 16, (1 << 22) + 16 * 37),
                                         (11, 0, (1 << 22) + (1 << 21))])
def test_device_randn_is_bit_identical_to_torch(seed, skip, n):
    """MT19937 outputs -> 24-bit uniforms -> ATen's 16-wide Box-Muller with its log / sincos polynomials: every float
    equals torch.randn's on this host, bit for bit."""
    from ccsmeth_b200 import _lib
    lib = _lib.load()
    out = np.empty(n, dtype=np.float32)
    _lib.check(lib.ccsm_debug_torch_randn(0, seed, skip, n, out.ctypes.data_as(ctypes.c_void_p)))
    torch.manual_seed(seed)
    if skip:
        torch.randn(skip)
    ref = torch.randn(n).numpy()
    bad = int((out.view(np.uint32) != ref.view(np.uint32)).sum())
    assert bad == 0, "%d of %d values differ (max |d| %.3e)" % (bad, n, np.abs(out - ref).max())


@pytest.fixture(scope="module")
def model(ckpt_att2s):
    from ccsmeth_b200.models import ModelAttRNN
    m = ModelAttRNN(21, 3, 2, 0, 256, is_npass=True, model_type="attbigru2s", device=0, precision="fp32")
    m.load_state_dict({k: torch.from_numpy(v) for k, v in ckpt_att2s.items()})
    return m.cuda(0).eval()


def test_forward_draws_the_reference_stream_and_advances_torch(model, golden_synth):
    """forward(h0=None) in the default mode == forward with the host-drawn stream, for every slicing the reference's
    batch loop can produce, and torch's generator ends where the reference's would."""
    from tests.test_parity_gpu import args16
    from ccsmeth_b200.call_modifications import draw_h0_stream_batches
    g = golden_synth
    a = [x.cuda() for x in args16(g)]
    n = 256
    for counts, bs in (([256], 512), ([256], 100), ([100, 56, 100], 64), ([1, 255], 512)):
        torch.manual_seed(99)
        h0 = draw_h0_stream_batches(counts, bs, 3, 256)
        after_host = torch.get_rng_state().clone()
        model.set_h0_mode("reference")
        _, p_host = model(*a, h0=h0)
        torch.manual_seed(99)
        model.set_h0_batching(counts, bs)
        _, p_dev = model(*a)
        assert torch.equal(torch.get_rng_state(), after_host), "torch's generator was not advanced like the reference's"
        assert np.array_equal(p_dev.cpu().numpy(), p_host.cpu().numpy()), (counts, bs)
        assert sum(counts) == n


def test_host_entry_and_chunking_use_the_same_stream(model, golden_synth):
    """More sites than one library chunk through the host entry: chunk cuts fall on model-call boundaries and the
    stream continues across them."""
    from tests.test_parity_gpu import FEATS
    from ccsmeth_b200.call_modifications import draw_h0_stream_batches
    g = golden_synth
    rep = 310   # 79,360 sites > 75,776
    feats = {k: np.concatenate([g[k]] * rep) for k in FEATS}
    n = 256 * rep
    counts = [5000] * (n // 5000) + [n % 5000]
    model.set_precision("fp16c8")
    torch.manual_seed(5)
    h0 = draw_h0_stream_batches(counts, 512, 3, 256)
    after_host = torch.get_rng_state().clone()
    _, p_host = model.forward_host(feats, h0=h0)
    torch.manual_seed(5)
    model.set_h0_batching(counts, 512)
    _, p_dev = model.forward_host(feats)
    assert torch.equal(torch.get_rng_state(), after_host)
    assert np.array_equal(p_dev.numpy(), p_host.numpy())
    model.set_precision("fp32")
