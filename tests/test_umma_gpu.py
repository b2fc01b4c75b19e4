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
