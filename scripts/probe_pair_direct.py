import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from ccsmeth_b200 import _lib
lib = _lib.load()
for flag in (0, 4, 5):
    N, K = 192, 64
    rng = np.random.default_rng(1)
    A = rng.standard_normal((256, K)).astype(np.float32); B = rng.standard_normal((N, K)).astype(np.float32)
    D = np.zeros((256, N), dtype=np.float32); Z = np.ones((256, 32), dtype=np.float32)
    vp = ctypes.c_void_p
    rc = lib.ccsm_debug_umma_pair_gemm(0, N, K, flag, A.ctypes.data_as(vp), B.ctypes.data_as(vp), D.ctypes.data_as(vp), Z.ctypes.data_as(vp))
    if rc != 0:
        print("flag", flag, "rc", rc, lib.ccsm_last_error().decode()); continue
    cv = (lambda x: x.half()) if flag & 1 else (lambda x: x.bfloat16())
    ref = cv(torch.from_numpy(A)).double().numpy() @ cv(torch.from_numpy(B)).double().numpy().T
    print("flag", flag, "maxerr %.3e" % np.abs(D - ref).max(), flush=True)
