#!/usr/bin/env python
"""Tensor-pipe issue-rate probe (ccsm_debug_umma_rate): cycles per MMA for kind::f16, kind::f8f6f4 (e4m3) and the
mixed patterns of the fp16c8 mode, one CTA, operands resident in shared memory."""
import ctypes
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ccsmeth_b200 import _lib

lib = _lib.load()
out = {}
names = {0: "f16 x4", 1: "e4m3 x4", 2: "f16 f16 e4m3 e4m3", 3: "f16 e4m3 f16 e4m3", 4: "rounds alternate f16x4 / e4m3x4"}
for N in (192, 256, 128, 64):
    for mode in range(5):
        res = []
        for iters in (64, 256):
            c = ctypes.c_int64(0)
            _lib.check(lib.ccsm_debug_umma_rate(0, N, mode, iters, ctypes.byref(c)))
            res.append(c.value)
        per = (res[1] - res[0]) / ((256 - 64) * 4.0)
        out["N%d %s" % (N, names[mode])] = per
        print("N=%3d  %-34s %7.1f cycles/MMA  (fixed %d)" % (N, names[mode], per, res[0] - per * 256))
for N in (256, 192, 128):
    for mode, nm in ((16, "pair f16 x4"), (17, "pair e4m3 x4"), (19, "pair f16 e4m3 f16 e4m3")):
        res = []
        for iters in (64, 256):
            c = ctypes.c_int64(0)
            _lib.check(lib.ccsm_debug_umma_rate(0, N, mode, iters, ctypes.byref(c)))
            res.append(c.value)
        per = (res[1] - res[0]) / ((256 - 64) * 4.0)
        out["N%d %s" % (N, nm)] = per
        print("N=%3d  %-34s %7.1f cycles/MMA  (fixed %d)" % (N, nm, per, res[0] - per * 256))
print(json.dumps(out))
