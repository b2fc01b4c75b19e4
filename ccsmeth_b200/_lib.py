"""ctypes binding of libccsm.so (C ABI declared in include/ccsm.h).

The shared library is built in-tree (``ccsmeth_b200/libccsm.so``) by ``build()`` with
``nvcc -gencode arch=compute_100a,code=sm_100a``.  There is deliberately no CPU fallback: if the
library is missing or fails to load, every product entry point raises.
"""
import ctypes
import glob
import os
import shutil
import subprocess
import threading

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_PATH = os.path.join(PKG_DIR, "libccsm.so")
HASH_PATH = LIB_PATH + ".srchash"   # sources the library was built from (git-ignored like the library, travels with it)
INCLUDE = os.path.join(os.path.dirname(PKG_DIR), "include")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]
LINK_LIBS = ["-lz", "-lpthread"]

# error codes / enums (mirror include/ccsm.h)
OK, EINVAL, ESTATE, ECUDA, ENOMEM, EUNSUPPORTED, EKEY = 0, -1, -2, -3, -4, -5, -6
KIND_ATT2S, KIND_AGGR = 0, 1
PREC = {"fp32": 0, "bf16x3": 1, "bf16": 2, "fp16x3": 3, "fp16": 4, "fp16c8": 5}
FEAT_NPASS, FEAT_STDS, FEAT_SN, FEAT_MAP, CELL_LSTM, MODEL_2S2, MODEL_TRANSENC = 1, 2, 4, 8, 16, 32, 64
AGGR_LSTM = 0x100
BGZF_RLE = 0x100

EXPORTS = ["ccsm_abi_version", "ccsm_last_error", "ccsm_kernel_launches", "ccsm_create", "ccsm_destroy",
           "ccsm_set_weight", "ccsm_finalize", "ccsm_set_precision", "ccsm_forward_att2s",
           "ccsm_forward_att2s_host", "ccsm_forward_att2s_lstm", "ccsm_forward_aggr", "ccsm_forward_aggr_lstm", "ccsm_debug_last_rnn_out", "ccsm_debug_umma_gemm",
           "ccsm_debug_tc_layer_out", "ccsm_profile_enable", "ccsm_profile_read", "ccsm_set_h0_mode",
           "ccsm_debug_umma_pair_gemm", "ccsm_debug_umma_mixed_gemm", "ccsm_debug_umma_rate", "ccsm_debug_torch_randn", "ccsm_set_h0_batching",
           "ccsm_h0_stream_set_state", "ccsm_h0_stream_get_state", "ccsm_bam_scan_records", "ccsm_forward_aggr_sites", "ccsm_debug_mt_jump_check", "ccsm_reads_extract_host", "ccsm_reads_sites", "ccsm_reads_features",
           "ccsm_reads_forward_host", "ccsm_bgzf_inflated_size", "ccsm_bgzf_inflate", "ccsm_bgzf_inflate_stats", "ccsm_bgzf_deflate_bound",
           "ccsm_bgzf_deflate", "ccsm_bam_index", "ccsm_bam_tag_records", "ccsm_bam_modcalls", "ccsm_pileup_luts",
           "ccsm_pileup_begin_host", "ccsm_pileup_finish_host", "ccsm_pileup_finish_lstm_host"]


class CcsmError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("libccsm error %d: %s" % (code, msg))
        self.code = code


class Config(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int32) for n in
                ("kind", "seq_len", "num_layers", "hidden", "num_classes", "n_vocab", "n_embed",
                 "feat_flags", "precision", "device")]


class Strand(ctypes.Structure):
    _fields_ = [(n, ctypes.c_void_p) for n in
                ("kmer", "kpass", "ipd_means", "ipd_stds", "pw_means", "pw_stds", "sns", "maps")]


class ExtractOpts(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int32) for n in ("mod_loc", "norm", "decode", "n_motifs", "motif_len")] + \
               [("motifs", ctypes.c_char * 64)]


class PileupOpts(ctypes.Structure):
    _fields_ = [("call_mode", ctypes.c_int32), ("cov_cf", ctypes.c_int32), ("prob_cf", ctypes.c_double),
                ("no_amb_cov", ctypes.c_int32), ("no_hap", ctypes.c_int32), ("discrete", ctypes.c_int32),
                ("only_close", ctypes.c_int32)]


class ModcallOpts(ctypes.Structure):
    _fields_ = [("mapq", ctypes.c_int32), ("no_supplementary", ctypes.c_int32), ("base_clip", ctypes.c_int32),
                ("hap_tag", ctypes.c_char * 4), ("identity", ctypes.c_double), ("refsites_all", ctypes.c_int32),
                ("n_refs", ctypes.c_int32), ("ref_off", ctypes.c_void_p), ("sites_fwd", ctypes.c_void_p),
                ("sites_rev", ctypes.c_void_p)]


class BamFilter(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int32) for n in ("mode_align", "mapq", "no_supplementary", "skip_unmapped", "want_sn", "pad_")] + \
               [("identity", ctypes.c_double)]

    def __init__(self, mode_align=0, mapq=0, no_supplementary=0, skip_unmapped=0, want_sn=0, identity=0.0):
        super().__init__(mode_align, mapq, no_supplementary, skip_unmapped, want_sn, 0, float(identity))


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def source_hash():
    """sha256 over every source the library is built from (names + contents, sorted): ties a loaded .so to its sources
    (the hash is printed by __graft_entry__.build() and recorded in profiles/sass_summary.md)."""
    import hashlib
    h = hashlib.sha256()
    deps = sources() + sorted(glob.glob(os.path.join(CSRC, "*.h"))) + sorted(glob.glob(os.path.join(CSRC, "*.cuh"))) + \
        sorted(glob.glob(os.path.join(INCLUDE, "*.h")))
    for p in deps:
        h.update(os.path.basename(p).encode())
        h.update(open(p, "rb").read())
    return h.hexdigest()


def _stale():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = sources() + glob.glob(os.path.join(CSRC, "*.h")) + glob.glob(os.path.join(CSRC, "*.cuh")) + \
        glob.glob(os.path.join(INCLUDE, "*.h"))
    return any(os.path.getmtime(p) > t for p in deps)


def build(force=False, verbose=False):
    """Compile csrc/*.cu into ccsmeth_b200/libccsm.so for sm_100a (cross-compiles without a GPU).
    One nvcc -c per source file (in parallel, objects cached under csrc/build/), then one link."""
    if not force and not _stale():
        if not os.path.exists(HASH_PATH):  # a library from before the record existed, newer than every source
            with open(HASH_PATH, "w") as f:
                f.write(source_hash() + "\n")
        return LIB_PATH
    from concurrent.futures import ThreadPoolExecutor
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    objdir = os.path.join(CSRC, "build")
    os.makedirs(objdir, exist_ok=True)
    hdrs = glob.glob(os.path.join(CSRC, "*.h")) + glob.glob(os.path.join(CSRC, "*.cuh")) + \
        glob.glob(os.path.join(INCLUDE, "*.h"))
    hdr_t = max(os.path.getmtime(p) for p in hdrs)
    flags = [f for f in NVCC_FLAGS if f != "-shared"]

    def compile_one(src):
        obj = os.path.join(objdir, os.path.basename(src)[:-3] + ".o")
        if not force and os.path.exists(obj) and os.path.getmtime(obj) > max(os.path.getmtime(src), hdr_t):
            return obj, ""
        cmd = [nvcc] + flags + (["-Xptxas", "-v"] if verbose else []) + ["-I", INCLUDE, "-c", "-o", obj, src]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("nvcc failed:\n%s\n%s" % (" ".join(cmd), res.stderr[-4000:]))
        return obj, res.stderr

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        done = list(ex.map(compile_one, sources()))
    tmp = LIB_PATH + ".tmp.%d" % os.getpid()
    cmd = [nvcc] + NVCC_FLAGS + ["-o", tmp] + [o for o, _ in done] + LINK_LIBS
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc link failed:\n%s\n%s" % (" ".join(cmd), res.stderr[-4000:]))
    os.replace(tmp, LIB_PATH)
    with open(HASH_PATH, "w") as f:  # which sources this binary was built from (checked by load())
        f.write(source_hash() + "\n")
    if verbose:
        print("".join(e for _, e in done))
    return LIB_PATH


def built_from():
    """Source hash recorded next to the library when it was built, or None (a library from before the record existed)."""
    try:
        return open(HASH_PATH).read().strip() or None
    except OSError:
        return None


_lib = None
_lock = threading.Lock()


def load():
    """Returns the loaded library; raises if it is not built (no fallback path exists)."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("%s is not built; run `python -c 'import __graft_entry__ as g; g.build()'` "
                               "(nvcc, sm_100a). ccsmeth_b200 has no CPU fallback." % LIB_PATH)
        rec = built_from()
        if rec is not None and rec != source_hash():
            # the binary rode along from an older tree (it is git-ignored, mtimes do not survive a copy): never run it
            try:
                build(force=True)
            except Exception as e:
                raise RuntimeError("%s was built from other sources (recorded %s..., tree %s...) and rebuilding failed: %s"
                                   % (LIB_PATH, rec[:16], source_hash()[:16], e))
        lib = ctypes.CDLL(LIB_PATH)
        vp, i32, i64 = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64
        lib.ccsm_abi_version.restype = ctypes.c_int
        lib.ccsm_last_error.restype = ctypes.c_char_p
        lib.ccsm_kernel_launches.restype = i64
        lib.ccsm_create.argtypes = [ctypes.POINTER(vp), ctypes.POINTER(Config)]
        lib.ccsm_destroy.argtypes = [vp]
        lib.ccsm_destroy.restype = None
        lib.ccsm_set_weight.argtypes = [vp, ctypes.c_char_p, vp, ctypes.POINTER(i64), i32]
        lib.ccsm_finalize.argtypes = [vp]
        lib.ccsm_set_precision.argtypes = [vp, i32]
        lib.ccsm_forward_att2s.argtypes = [vp, i64, ctypes.POINTER(Strand), ctypes.POINTER(Strand), vp, vp, vp, vp, vp]
        lib.ccsm_forward_att2s_host.argtypes = [vp, i64, ctypes.POINTER(Strand), ctypes.POINTER(Strand), vp, vp, vp, vp]
        lib.ccsm_forward_att2s_lstm.argtypes = [vp, i64, ctypes.POINTER(Strand), ctypes.POINTER(Strand), vp, vp, vp, vp, vp, vp, vp]
        lib.ccsm_forward_att2s_lstm.restype = ctypes.c_int
        lib.ccsm_forward_aggr.argtypes = [vp, i64, vp, vp, vp, vp, vp]
        lib.ccsm_forward_aggr_lstm.argtypes = [vp, i64, vp, vp, vp, vp, vp, vp]
        lib.ccsm_forward_aggr_lstm.restype = ctypes.c_int
        lib.ccsm_debug_last_rnn_out.argtypes = [vp, vp, i64]
        lib.ccsm_debug_last_rnn_out.restype = i64
        lib.ccsm_debug_tc_layer_out.argtypes = [vp, i32, vp, i64]
        lib.ccsm_debug_tc_layer_out.restype = i64
        lib.ccsm_debug_umma_pair_gemm.argtypes = [i32, i32, i32, i32, vp, vp, vp, vp]
        lib.ccsm_debug_umma_pair_gemm.restype = ctypes.c_int
        lib.ccsm_debug_umma_mixed_gemm.argtypes = [i32, i32, i32, vp, vp, vp]
        lib.ccsm_debug_umma_mixed_gemm.restype = ctypes.c_int
        lib.ccsm_debug_umma_rate.argtypes = [i32, i32, i32, i32, vp]
        lib.ccsm_debug_umma_rate.restype = ctypes.c_int
        lib.ccsm_debug_torch_randn.argtypes = [i32, ctypes.c_uint64, i64, i64, vp]
        lib.ccsm_debug_torch_randn.restype = ctypes.c_int
        lib.ccsm_set_h0_batching.argtypes = [vp, vp, i64, i32]
        lib.ccsm_set_h0_batching.restype = ctypes.c_int
        lib.ccsm_h0_stream_set_state.argtypes = [vp, vp, i32]
        lib.ccsm_h0_stream_set_state.restype = ctypes.c_int
        lib.ccsm_h0_stream_get_state.argtypes = [vp, vp, vp]
        lib.ccsm_h0_stream_get_state.restype = ctypes.c_int
        lib.ccsm_bam_scan_records.argtypes = [vp, i64, i64, vp, vp, vp, vp, vp, vp, vp, vp]
        lib.ccsm_bam_scan_records.restype = ctypes.c_int64
        lib.ccsm_forward_aggr_sites.argtypes = [vp, i64, vp, vp, i32, vp, vp, vp]
        lib.ccsm_forward_aggr_sites.restype = ctypes.c_int
        lib.ccsm_debug_mt_jump_check.argtypes = [ctypes.c_uint32, i32]
        lib.ccsm_debug_mt_jump_check.restype = ctypes.c_int
        lib.ccsm_set_h0_mode.argtypes = [vp, i32, ctypes.c_uint64]
        lib.ccsm_set_h0_mode.restype = ctypes.c_int
        lib.ccsm_profile_enable.argtypes = [vp, i32]
        lib.ccsm_profile_enable.restype = ctypes.c_int
        lib.ccsm_profile_read.argtypes = [vp, vp, vp, vp, i32]
        lib.ccsm_profile_read.restype = ctypes.c_int
        lib.ccsm_debug_umma_gemm.argtypes = [i32, i32, i32, i32, i32, vp, vp, vp]
        lib.ccsm_debug_umma_gemm.restype = ctypes.c_int
        lib.ccsm_reads_extract_host.argtypes = [vp, ctypes.POINTER(ExtractOpts), vp, i64, vp, i32, ctypes.POINTER(i64)]
        lib.ccsm_reads_sites.argtypes = [vp, vp, vp]
        lib.ccsm_reads_features.argtypes = [vp, i64, i64, ctypes.POINTER(Strand), ctypes.POINTER(Strand), vp]
        lib.ccsm_reads_forward_host.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp]
        lib.ccsm_bgzf_inflated_size.argtypes = [vp, i64, ctypes.POINTER(i64)]
        lib.ccsm_bgzf_inflate.argtypes = [vp, i64, vp, i64, i32, ctypes.POINTER(i64)]
        lib.ccsm_bgzf_inflate_stats.argtypes = [ctypes.POINTER(i64), ctypes.POINTER(i64)]
        lib.ccsm_bgzf_inflate_stats.restype = None
        lib.ccsm_bgzf_deflate_bound.argtypes = [i64]
        lib.ccsm_bgzf_deflate.argtypes = [vp, i64, vp, i64, i32, i32]
        for fn in ("ccsm_bgzf_inflated_size", "ccsm_bgzf_inflate", "ccsm_bgzf_inflate_stats", "ccsm_bgzf_deflate_bound", "ccsm_bgzf_deflate"):
            getattr(lib, fn).restype = i64
        lib.ccsm_pileup_luts.argtypes = [ctypes.POINTER(PileupOpts), i32, vp, vp]
        lib.ccsm_pileup_begin_host.argtypes = [vp, ctypes.POINTER(PileupOpts), i64, vp, vp, vp, vp, vp]
        lib.ccsm_pileup_finish_host.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp]
        lib.ccsm_pileup_finish_lstm_host.argtypes = [vp, ctypes.POINTER(vp), ctypes.POINTER(vp), vp, vp, vp, vp]
        for fn in ("ccsm_pileup_luts", "ccsm_pileup_begin_host", "ccsm_pileup_finish_host", "ccsm_pileup_finish_lstm_host"):
            getattr(lib, fn).restype = ctypes.c_int
        lib.ccsm_bam_modcalls.argtypes = [vp, vp, i32, ctypes.POINTER(ModcallOpts), vp, vp, vp, vp, vp, i64, ctypes.POINTER(i32)]
        lib.ccsm_bam_modcalls.restype = i64
        lib.ccsm_bam_index.argtypes = [vp, i64, ctypes.POINTER(BamFilter), vp, i32, vp, ctypes.POINTER(i32),
                                       ctypes.POINTER(i32), ctypes.POINTER(i64)]
        lib.ccsm_bam_index.restype = ctypes.c_int
        lib.ccsm_bam_tag_records.argtypes = [vp, vp, i32, i32, vp, vp, vp, vp, i64, ctypes.POINTER(i32)]
        lib.ccsm_bam_tag_records.restype = i64
        for fn in ("ccsm_reads_extract_host", "ccsm_reads_sites", "ccsm_reads_features", "ccsm_reads_forward_host"):
            getattr(lib, fn).restype = ctypes.c_int
        for fn in ("ccsm_create", "ccsm_set_weight", "ccsm_finalize", "ccsm_set_precision", "ccsm_forward_att2s",
                   "ccsm_forward_att2s_host", "ccsm_forward_aggr"):
            getattr(lib, fn).restype = ctypes.c_int
        _lib = lib
    return _lib


def check(rc):
    if rc != OK:
        raise CcsmError(rc, load().ccsm_last_error().decode("utf-8", "replace"))


def kernel_launches():
    return int(load().ccsm_kernel_launches())
