"""numpy restatement of ``torch.randn`` on the CPU default generator -- TEST ORACLE, not product code.

The reference draws the GRU initial state with ``torch.randn(2*layers, n, hidden)`` (reference ccsmeth/models.py:77-87)
on the process-wide CPU generator seeded by ``torch.manual_seed`` (call_modifications.py:479-481).  ATen implements that
draw as follows (aten/src/ATen/native/cpu/DistributionTemplates.h ``normal_fill`` -- the path taken for contiguous
float32 tensors of >= 16 elements on builds dispatching to the AVX512 or DEFAULT kernels; the AVX2 kernel uses
polynomial approximations of log/sin/cos and differs in the last bits):

  * engine: MT19937 (``at::mt19937``), state initialised from the low 32 bits of the seed exactly like
    ``init_genrand``; one 32-bit output per element;
  * uniform: ``(x & (2^24 - 1)) * 2^-24`` (``at::uniform_real_distribution<float>``);
  * Box-Muller on groups of 16 consecutive elements: for j < 8, ``u1 = 1 - u[j]``, ``u2 = u[j + 8]``,
    ``r = sqrtf(-2 logf(u1))``, ``theta = float(2 pi (double) u2)``, ``out[j] = r cosf(theta)``,
    ``out[j + 8] = r sinf(theta)`` with the C library's float functions.

``randn(seed_state, n)`` reproduces it with the host libm through ctypes (n a multiple of 16).  The library's device
implementation (csrc/mtstream.cu) restates glibc's logf / sinf / cosf in IEEE double operations; scripts/check_glibcf.c
checks that restatement against the host libm on every reachable input (2^24 uniforms).
"""
import ctypes
import ctypes.util

import numpy as np

N, M = 624, 397


class MT19937:
    def __init__(self, seed):
        s = np.zeros(N, dtype=np.uint64)
        s[0] = int(seed) & 0xFFFFFFFF
        for j in range(1, N):
            s[j] = (1812433253 * (int(s[j - 1]) ^ (int(s[j - 1]) >> 30)) + j) & 0xFFFFFFFF
        self.state = s.astype(np.uint32)
        self.pos = N   # next output triggers a twist

    def _twist(self):
        s = self.state.astype(np.uint64)
        out = np.empty(N, dtype=np.uint64)

        def tw(u, v):
            y = (u & 0x80000000) | (v & 0x7FFFFFFF)
            return (y >> 1) ^ np.where(y & 1, 0x9908B0DF, 0).astype(np.uint64)
        out[:N - M] = s[M:] ^ tw(s[:N - M], s[1:N - M + 1])
        for base in (N - M, 2 * (N - M)):          # words that depend on freshly generated ones, 227 at a time
            hi = min(base + (N - M), N - 1)
            out[base:hi] = out[base - (N - M):hi - (N - M)] ^ tw(s[base:hi], s[base + 1:hi + 1])
        out[N - 1] = out[M - 1] ^ tw(s[N - 1:N], out[0:1])[0]
        self.state = out.astype(np.uint32)
        self.pos = 0

    def raw(self, n):
        """next n tempered 32-bit outputs"""
        res = np.empty(n, dtype=np.uint32)
        got = 0
        while got < n:
            if self.pos == N:
                self._twist()
            k = min(n - got, N - self.pos)
            y = self.state[self.pos:self.pos + k].astype(np.uint64)
            y ^= y >> 11
            y ^= (y << 7) & 0x9D2C5680
            y ^= (y << 15) & 0xEFC60000
            y ^= y >> 18
            res[got:got + k] = y.astype(np.uint32)
            self.pos += k
            got += k
        return res


_libm = None


def _vec(fn_name, x):
    global _libm
    if _libm is None:
        _libm = ctypes.CDLL(ctypes.util.find_library("m"))
        for f in ("logf", "sinf", "cosf"):
            getattr(_libm, f).restype = ctypes.c_float
            getattr(_libm, f).argtypes = [ctypes.c_float]
    fn = getattr(_libm, fn_name)
    return np.array([fn(float(v)) for v in x], dtype=np.float32)


def normal_from_raw(raw):
    """ATen's normal_fill on raw MT19937 outputs (length multiple of 16) -> float32 N(0, 1) values."""
    raw = np.asarray(raw, dtype=np.uint32)
    assert raw.size % 16 == 0
    u = ((raw & 0xFFFFFF).astype(np.float64) * 2.0 ** -24).astype(np.float32).reshape(-1, 2, 8)
    u1 = (np.float32(1.0) - u[:, 0, :]).ravel()
    u2 = u[:, 1, :].ravel()
    radius = np.sqrt(np.float32(-2.0) * _vec("logf", u1)).astype(np.float32)
    theta = (6.283185307179586 * u2.astype(np.float64)).astype(np.float32)
    out = np.empty_like(u)
    out[:, 0, :] = (radius * _vec("cosf", theta)).reshape(-1, 8)
    out[:, 1, :] = (radius * _vec("sinf", theta)).reshape(-1, 8)
    return out.ravel()


def randn(gen, n):
    return normal_from_raw(gen.raw(n))
