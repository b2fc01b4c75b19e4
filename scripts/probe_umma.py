import sys; sys.path.insert(0,'.')
import numpy as np
from tests.test_umma_gpu import _run
for swap in (0,1):
    for (N,K) in [(192,64),(64,16),(128,32),(256,128)]:
        for f16 in (0,1):
            try:
                D, ref = _run(N,K,f16,swap)
                print("swap",swap,"N",N,"K",K,"f16",f16,"maxerr %.3e"%np.abs(D-ref).max(), "ref max %.2f"%np.abs(ref).max(), flush=True)
            except Exception as e:
                print("swap",swap,N,K,f16,"EXC",e, flush=True)
