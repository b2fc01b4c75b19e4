"""The BGZF block codec in libccsm (csrc/hostio.cu, csrc/inflate_fast.h) against Python's zlib: the table decoder on
every DEFLATE block type and compressor strategy, the zlib fallback on blocks it must reject, the run-length
strategy of the writer, and corruption detection.  Host code only."""
import ctypes
import os
import zlib

import numpy as np
import pytest

from ccsmeth_b200 import _lib
from ccsmeth_b200.bamio import _BGZF_EOF, BgzfWriter, _deflate_block
from tests.conftest import GOLDEN

DEMO = os.path.join(GOLDEN, "demo", "hg002.chr20_demo.hifi.bam")
BLOCK = 65280


def _stats(lib):
    a, b = ctypes.c_int64(), ctypes.c_int64()
    lib.ccsm_bgzf_inflate_stats(ctypes.byref(a), ctypes.byref(b))
    return a.value, b.value


def _inflate(lib, blob, threads=3):
    used = ctypes.c_int64()
    size = lib.ccsm_bgzf_inflated_size(blob, len(blob), ctypes.byref(used))
    assert size >= 0
    out = np.empty(max(int(size), 1), dtype=np.uint8)
    got = lib.ccsm_bgzf_inflate(blob, len(blob), out.ctypes.data, int(size), threads, ctypes.byref(used))
    return got, out[:max(int(got), 0)].tobytes(), used.value


def _py_inflate(blob):
    out, p = [], 0
    while p < len(blob):
        bsize = int.from_bytes(blob[p + 16:p + 18], "little") + 1
        out.append(zlib.decompress(blob[p + 18:p + bsize - 8], -15))
        p += bsize
    return b"".join(out)


def _payloads():
    rng = np.random.default_rng(11)
    demo = _py_inflate(open(DEMO, "rb").read())[:6 * BLOCK]
    text = ("\n".join("read%d\t0\tchr1\t%d\t60\t100M\t*\t0\t0\t%s\t%s" % (i, i * 37, "ACGT" * 25, "I" * 100)
                      for i in range(3000))).encode()
    return {
        "hifi_records": demo,
        "text": text,
        "random": rng.integers(0, 256, 3 * BLOCK + 17, dtype=np.uint8).tobytes(),   # stored blocks inside zlib's output
        "zeros": bytes(2 * BLOCK),                                                 # longest matches, distance 1
        "period3": (b"abc" * (BLOCK // 3 + 5))[:BLOCK],                             # overlapping copies, distance < 8
        "skewed": rng.choice(np.arange(256, dtype=np.uint8), size=2 * BLOCK,
                             p=np.r_[0.9, np.full(255, 0.1 / 255)]).tobytes(),      # codes longer than the primary table
        "tiny": b"x",
    }


@pytest.mark.parametrize("strategy,level", [(zlib.Z_DEFAULT_STRATEGY, 6), (zlib.Z_DEFAULT_STRATEGY, 9),
                                            (zlib.Z_DEFAULT_STRATEGY, 1), (zlib.Z_DEFAULT_STRATEGY, 0),
                                            (zlib.Z_HUFFMAN_ONLY, 6), (zlib.Z_RLE, 6), (zlib.Z_FIXED, 6),
                                            (zlib.Z_FILTERED, 6)])
def test_table_decoder_handles_every_block_type(strategy, level):
    lib = _lib.load()
    for name, data in _payloads().items():
        blob = b"".join(_deflate_block(data[i:i + BLOCK], level, strategy) for i in range(0, len(data), BLOCK))
        blob += _BGZF_EOF  # the empty EOF block decodes to nothing
        before = _stats(lib)
        got, out, used = _inflate(lib, blob)
        after = _stats(lib)
        assert got == len(data) and out == data and used == len(blob), name
        assert after[1] == before[1], "%s: a block fell back to zlib" % name
        assert after[0] - before[0] == -(-len(data) // BLOCK), name


def test_demo_bam_decodes_without_fallback_and_like_zlib(monkeypatch):
    lib = _lib.load()
    blob = open(DEMO, "rb").read()
    before = _stats(lib)
    got, fast, _ = _inflate(lib, blob, threads=4)
    after = _stats(lib)
    assert after[1] == before[1] and after[0] > before[0]
    monkeypatch.setenv("CCSM_INFLATE", "zlib")
    got2, slow, _ = _inflate(lib, blob, threads=4)
    mid = _stats(lib)
    assert mid[0] == after[0] and mid[1] > after[1]
    assert got == got2 and fast == slow == _py_inflate(blob)


def test_corruption_is_detected():
    lib = _lib.load()
    data = _payloads()["hifi_records"][:2 * BLOCK]
    blob = bytearray(b"".join(_deflate_block(data[i:i + BLOCK], 6) for i in range(0, len(data), BLOCK)))
    for pos in (40, 1000, len(blob) - 6):  # payload bytes of the first block, CRC of the last block
        bad = bytearray(blob)
        bad[pos] ^= 0x5a
        got, _, _ = _inflate(lib, bytes(bad))
        assert got == _lib.EINVAL and b"corrupt" in lib.ccsm_last_error()


def test_truncated_tail_is_left_for_the_next_call():
    lib = _lib.load()
    data = _payloads()["text"]
    blocks = [_deflate_block(data[i:i + BLOCK], 6) for i in range(0, len(data), BLOCK)]
    blob = b"".join(blocks)
    cut = blob[:len(blob) - 10]  # the last block is incomplete
    got, out, used = _inflate(lib, cut)
    assert used == len(blob) - len(blocks[-1]) and out == data[:BLOCK * (len(blocks) - 1)]


@pytest.mark.parametrize("strategy", ["rle", "zlib"])
@pytest.mark.parametrize("threads", [1, 4])
def test_writer_round_trips_and_rle_is_not_larger_on_hifi_records(tmp_path, strategy, threads):
    lib = _lib.load()
    data = _payloads()["hifi_records"]
    path = str(tmp_path / "x.bgzf")
    w = BgzfWriter(path, threads=threads, strategy=strategy)
    w.write(data[:100])                                       # small write: buffered
    w.write(np.frombuffer(data[100:], dtype=np.uint8))        # large write: straight to the thread team when native
    w.close()
    blob = open(path, "rb").read()
    assert blob.endswith(_BGZF_EOF)
    got, out, _ = _inflate(lib, blob)
    assert out == data == _py_inflate(blob)
    if strategy == "rle":
        ref = sum(len(_deflate_block(data[i:i + BLOCK], 6)) for i in range(0, len(data), BLOCK))
        assert len(blob) <= 1.03 * ref  # run-length + Huffman is within 3 % of zlib's default strategy on these records


def test_deflate_rejects_bad_arguments():
    lib = _lib.load()
    src = np.zeros(10, dtype=np.uint8)
    dst = np.zeros(int(lib.ccsm_bgzf_deflate_bound(10)), dtype=np.uint8)
    assert lib.ccsm_bgzf_deflate(src.ctypes.data, 10, dst.ctypes.data, 5, 6, 1) == _lib.EINVAL      # capacity
    assert lib.ccsm_bgzf_deflate(src.ctypes.data, 10, dst.ctypes.data, len(dst), 12, 1) == _lib.EINVAL  # level
    assert lib.ccsm_bgzf_deflate(src.ctypes.data, 10, dst.ctypes.data, len(dst), 6 | _lib.BGZF_RLE, 1) > 0


def test_random_corruptions_never_pass_silently():
    """Garbage in the deflate payloads: every call either reports corruption or (when the damaged bytes decode to the
    same content) returns the original bytes -- and the process survives bounds-wise."""
    lib = _lib.load()
    rng = np.random.default_rng(99)
    payloads = _payloads()
    for name in ("hifi_records", "text", "skewed", "period3"):
        data = payloads[name][:2 * BLOCK]
        blob = b"".join(_deflate_block(data[i:i + BLOCK], 6) for i in range(0, len(data), BLOCK))
        first = int.from_bytes(blob[16:18], "little") + 1
        for _ in range(60):
            bad = bytearray(blob)
            for pos in rng.integers(18, first - 8, size=int(rng.integers(1, 4))):  # inside the first block's payload
                bad[pos] = int(rng.integers(0, 256))
            got, out, _ = _inflate(lib, bytes(bad), threads=2)
            assert got == _lib.EINVAL or out == data, name


def test_own_rle_encoder_matches_zlib_rle(monkeypatch):
    """CCSM_BGZF_RLE runs csrc/deflate_rle.h: any inflater must read its blocks, and its sizes must be those of zlib's
    Z_RLE strategy (same token stream, same 32 K-token sub-blocks) -- incompressible input is stored, never expanded."""
    lib = _lib.load()
    rng = np.random.default_rng(3)
    payloads = dict(_payloads())
    payloads["runs"] = b"".join(bytes([int(rng.integers(0, 256))]) * int(rng.integers(1, 600)) for _ in range(400))
    payloads["empty"] = b""
    geo = np.minimum(rng.geometric(0.5, size=BLOCK), 60).astype(np.uint8).tobytes()   # deep Huffman trees
    payloads["geometric"] = geo
    for name, data in payloads.items():
        src = np.frombuffer(data, dtype=np.uint8)
        cap = int(lib.ccsm_bgzf_deflate_bound(len(data)))
        sizes = {}
        for impl in ("own", "zlib"):
            if impl == "zlib":
                monkeypatch.setenv("CCSM_DEFLATE", "zlib")
            else:
                monkeypatch.delenv("CCSM_DEFLATE", raising=False)
            dst = np.empty(cap, dtype=np.uint8)
            got = lib.ccsm_bgzf_deflate(src.ctypes.data if len(data) else None, len(data), dst.ctypes.data, cap,
                                        6 | _lib.BGZF_RLE, 3)
            assert got >= 0, (name, impl)
            blob = dst[:got].tobytes()
            assert _py_inflate(blob) == data, (name, impl)            # Python's zlib reads it
            if len(data):
                g2, out, _ = _inflate(lib, blob)
                assert g2 == len(data) and out == data, (name, impl)   # and so does the table decoder
            sizes[impl] = got
        assert sizes["own"] <= sizes["zlib"] * 1.002 + 16, (name, sizes)
        assert sizes["own"] <= len(data) + 40 * (len(data) // BLOCK + 1), name   # stored fallback: 26 B BGZF frame + 5 B per stored sub-block
