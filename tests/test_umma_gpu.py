"""Pins the tcgen05 conventions (smem descriptor LBO/SBO, idesc bits, TMEM ld mapping) with a one-CTA GEMM."""
import ctypes

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _run(N, K, f16, swap):
    from ccsmeth_b200 import _lib
    lib = _lib.load()
    rng = np.random.default_rng(N * 1000 + K)
    A = rng.standard_normal((128, K)).astype(np.float32)
    B = rng.standard_normal((N, K)).astype(np.float32)
    D = np.zeros((128, N), dtype=np.float32)
    vp = ctypes.c_void_p
    _lib.check(lib.ccsm_debug_umma_gemm(0, N, K, int(f16), int(swap), A.ctypes.data_as(vp), B.ctypes.data_as(vp),
                                        D.ctypes.data_as(vp)))
    dt = torch.float16 if f16 else torch.bfloat16
    Ar = torch.from_numpy(A).to(dt).double().numpy()
    Br = torch.from_numpy(B).to(dt).double().numpy()
    return D, Ar @ Br.T


@pytest.mark.parametrize("N,K", [(192, 64), (64, 16), (128, 32), (256, 128), (192, 16)])
@pytest.mark.parametrize("f16", [False, True])
def test_umma_gemm_matches_numpy(N, K, f16):
    D, ref = _run(N, K, f16, swap=False)
    err = np.abs(D - ref).max()
    assert err < 1e-3 * max(1.0, np.abs(ref).max()), err


@pytest.mark.parametrize("N,K", [(192, 64), (64, 16), (256, 128)])
@pytest.mark.parametrize("f16", [False, True])
def test_umma_pair_gemm_matches_numpy(N, K, f16):
    """cta_group::2: M = 256 across a CTA pair, B rows split half/half, multicast commit, remote mbarrier arrive,
    tcgen05.st zeroing."""
    from ccsmeth_b200 import _lib
    lib = _lib.load()
    rng = np.random.default_rng(N * 7 + K)
    A = rng.standard_normal((256, K)).astype(np.float32)
    B = rng.standard_normal((N, K)).astype(np.float32)
    D = np.zeros((256, N), dtype=np.float32)
    Z = np.ones((256, 32), dtype=np.float32)
    vp = ctypes.c_void_p
    _lib.check(lib.ccsm_debug_umma_pair_gemm(0, N, K, int(f16), A.ctypes.data_as(vp), B.ctypes.data_as(vp),
                                             D.ctypes.data_as(vp), Z.ctypes.data_as(vp)))
    dt = torch.float16 if f16 else torch.bfloat16
    ref = torch.from_numpy(A).to(dt).double().numpy() @ torch.from_numpy(B).to(dt).double().numpy().T
    assert np.abs(D - ref).max() < 1e-3 * max(1.0, np.abs(ref).max())
    assert np.all(Z[:, :16] == 0.0)                      # tcgen05.st zeroed columns 16..31
    assert np.abs(Z[:, 16:] - ref[:, :16]).max() < 1e-3 * max(1.0, np.abs(ref).max())  # neighbours untouched


@pytest.mark.parametrize("N,K", [(192, 32), (192, 64), (256, 128), (64, 32)])
def test_umma_mixed_f16_e4m3_accumulate(N, K):
    """kind::f16 and kind::f8f6f4 (e4m3, K = 32 per MMA, 16-element slabs) accumulating into the same TMEM columns:
    the layout and the mixing the fp16c8 precision mode is built on."""
    from ccsmeth_b200 import _lib
    lib = _lib.load()
    rng = np.random.default_rng(N * 31 + K)
    A = rng.standard_normal((128, K)).astype(np.float32)
    B = rng.standard_normal((N, K)).astype(np.float32)
    D = np.zeros((128, N), dtype=np.float32)
    vp = ctypes.c_void_p
    _lib.check(lib.ccsm_debug_umma_mixed_gemm(0, N, K, A.ctypes.data_as(vp), B.ctypes.data_as(vp), D.ctypes.data_as(vp)))
    q16 = lambda x: torch.from_numpy(x).to(torch.float16).double().numpy()
    q8 = lambda x: torch.from_numpy(x).to(torch.float8_e4m3fn).double().numpy()
    ref = q16(A) @ q16(B).T + q8(A) @ q8(B).T
    assert np.abs(D - ref).max() < 1e-3 * max(1.0, np.abs(ref).max())
