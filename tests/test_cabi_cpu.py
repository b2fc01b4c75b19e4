"""C-ABI surface checks that need no GPU: the library builds for sm_100a, loads, exports every symbol
include/ccsm.h declares, and fails cleanly (no crash, message set) when no device is present."""
import ctypes
import os
import re

import pytest
import torch

from ccsmeth_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    _lib.build()
    return _lib.load()


def test_header_symbols_are_exported(lib):
    hdr = open(os.path.join(ROOT, "include", "ccsm.h")).read()
    declared = set(re.findall(r"\b(ccsm_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    assert declared == set(_lib.EXPORTS)
    for name in declared:
        assert hasattr(lib, name), name


def test_abi_version_and_error_string(lib):
    assert lib.ccsm_abi_version() == 1
    assert lib.ccsm_last_error() is not None


def test_create_rejects_bad_config(lib):
    h = ctypes.c_void_p()
    cfg = _lib.Config(7, 21, 3, 256, 2, 5, 8, 1, 0, 0)  # unknown kind
    assert lib.ccsm_create(ctypes.byref(h), ctypes.byref(cfg)) == _lib.EINVAL
    assert b"kind" in lib.ccsm_last_error()
    cfg = _lib.Config(0, 22, 3, 256, 2, 5, 8, 1, 0, 0)  # even seq_len (reference call_modifications.py:500-501)
    assert lib.ccsm_create(ctypes.byref(h), ctypes.byref(cfg)) == _lib.EINVAL
    assert lib.ccsm_create(None, None) == _lib.EINVAL


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-device failure mode")
def test_no_device_fails_loudly(lib):
    h = ctypes.c_void_p()
    cfg = _lib.Config(0, 21, 3, 256, 2, 5, 8, 1, 0, 0)
    rc = lib.ccsm_create(ctypes.byref(h), ctypes.byref(cfg))
    assert rc != 0 and not h.value
    assert len(lib.ccsm_last_error()) > 0
    from ccsmeth_b200.models import ModelAttRNN
    m = ModelAttRNN(21, 3, 2, 0, 256)
    z = torch.zeros(2, 21)
    with pytest.raises(RuntimeError):  # no CPU fallback
        m(z, z, z, z, z, z, z, z, z, z, z, z, z, z, z, z)


def test_state_dict_contract(ckpt_att2s, ckpt_aggr):
    from ccsmeth_b200.models import ModelAttRNN, AggrAttRNN
    m = ModelAttRNN(21, 3, 2, 0, 256, is_npass=True, model_type="attbigru2s", device=0)
    sd = {k: torch.from_numpy(v) for k, v in ckpt_att2s.items()}
    d = m.state_dict()
    d.update(sd)
    m.load_state_dict(d)  # the reference's load sequence (call_modifications.py:343-347)
    assert {k: tuple(v.shape) for k, v in m.state_dict().items()} == {k: v.shape for k, v in ckpt_att2s.items()}
    a = AggrAttRNN(11, 1, 1, 0, 32, binsize=20, model_type="attbigru", device="cpu")
    a.load_state_dict({k: torch.from_numpy(v) for k, v in ckpt_aggr.items()})  # "module." prefix stripped
    assert a.get_model_type() == "attbigru" and m.get_model_type() == "attbigru2s"
    with pytest.raises(ValueError):
        ModelAttRNN(model_type="transencoder2s")  # that model type has its own class, like in the reference
    lstm = ModelAttRNN(21, 2, 2, 0, 64, model_type="attbilstm2s")  # reference models.py:48-51: nn.LSTM parameter shapes
    assert lstm.state_dict()["rnn.weight_ih_l0"].shape == (256, 11) and lstm.get_precision() == "fp32"
    h, c = lstm.init_hidden(5, 2, 64)
    assert h.shape == c.shape == (4, 5, 64)


def test_library_is_tied_to_its_sources(tmp_path, monkeypatch):
    """build() records the hash of the sources next to the library; load() refuses to run a binary that was built from
    other sources (it rebuilds; if that fails the error names both hashes)."""
    from ccsmeth_b200 import _lib as L
    assert L.built_from() == L.source_hash()
    # a stale record + no working compiler -> a loud error, not a silently stale binary
    stale = str(tmp_path / "libccsm.so.srchash")
    open(stale, "w").write("0" * 64 + "\n")
    monkeypatch.setattr(L, "HASH_PATH", stale)
    monkeypatch.setattr(L, "_lib", None)

    def no_build(force=False, verbose=False):
        raise RuntimeError("nvcc unavailable (test)")
    monkeypatch.setattr(L, "build", no_build)
    with pytest.raises(RuntimeError, match="built from other sources"):
        L.load()
