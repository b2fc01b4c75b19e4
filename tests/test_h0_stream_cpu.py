"""Host-side pieces of the reference h0 stream (csrc/mtstream.cu, csrc/mtjump.h) that need no GPU: the numpy restatement of
torch's generator against torch itself, the torch generator-state layout the Python mirror relies on, and the GF(2)
jump-ahead polynomials."""
import numpy as np
import pytest
import torch

from ccsmeth_b200 import _lib
from oracle.torch_randn_numpy import MT19937


def test_numpy_mt19937_is_torchs_engine():
    """torch.manual_seed -> the same 624 state words as init_genrand; get_rng_state exposes them as uint64 at byte 24,
    `left` at byte 8 and `next` at byte 16 (the layout ccsmeth_b200.models._borrow_torch_rng reads and writes)."""
    for seed in (0, 77, 20261017):
        torch.manual_seed(seed)
        st = torch.get_rng_state().numpy()
        assert len(st) == 5056
        g = MT19937(seed)
        assert np.array_equal(st[24:24 + 4992].view(np.uint64).astype(np.uint32), g.state)
        assert int.from_bytes(st[8:12].tobytes(), "little", signed=True) == 1       # left: the next draw twists
        torch.randn(160)
        g.raw(160)
        st = torch.get_rng_state().numpy()
        assert np.array_equal(st[24:24 + 4992].view(np.uint64).astype(np.uint32), g.state)
        assert int.from_bytes(st[16:24].tobytes(), "little") == g.pos == 160
        assert int.from_bytes(st[8:12].tobytes(), "little", signed=True) == 625 - 160


def test_set_rng_state_round_trip_continues_the_stream():
    """Writing an advanced state back (what _return_torch_rng does) makes torch continue exactly there."""
    torch.manual_seed(5)
    a = torch.randn(6 * 7 * 256)
    b = torch.randn(1024)
    torch.manual_seed(5)
    st = torch.get_rng_state().numpy().copy()
    g = MT19937(5)
    g.raw(a.numel())
    st[24:24 + 4992] = g.state.astype(np.uint64).view(np.uint8)
    p = g.pos
    st[8:12] = np.frombuffer(int(1 if p == 624 else 625 - p).to_bytes(4, "little", signed=True), dtype=np.uint8)
    st[16:24] = np.frombuffer(int(p).to_bytes(8, "little"), dtype=np.uint8)
    torch.set_rng_state(torch.from_numpy(st))
    assert torch.equal(torch.randn(1024), b)


@pytest.mark.parametrize("k", [0, 1, 3])
def test_jump_polynomials_reproduce_plain_generation(k):
    """x^(2^k 2^21) mod phi applied to a state == the state that many words further on (host self-check in the library)."""
    lib = _lib.load()
    assert lib.ccsm_debug_mt_jump_check(1234 + k, k) == 0
