"""Host-side mirror of the reference model interface for the call_mods inference path.

``ModelAttRNN`` / ``AggrAttRNN`` keep the reference's constructor arguments, ``state_dict`` keys and
shapes, ``.cuda(device)`` / ``.eval()`` / ``get_model_type()`` and the 16-tensor ``forward`` ->
``(logits, probs)`` contract (reference ccsmeth/models.py:17-150, 625-694), so the reference's call
sites (call_modifications.py:316-369, 201-214; call_mods_freq_bam.py:317-342, 301) work unchanged.
The arithmetic is NOT torch: ``forward`` hands raw device pointers to libccsm.so (include/ccsm.h),
which launches the sm_100a kernels.  torch is used for parameter storage, device memory and streams.

Extension over the reference: ``forward(..., h0=(h0_strand1, h0_strand2))``.  The reference draws the
GRU initial state with ``torch.randn`` on every call (models.py:77-87), so parity is only defined
with h0 as a shared explicit input.  ``h0=None`` reproduces the reference exactly: two
``torch.randn(2*layers, n, hidden)`` draws from the CPU default generator, strand 1 then strand 2.
"""
import ctypes
import os

import numpy as np
import torch
import torch.nn as nn

from . import _lib
from .utils.process_utils import N_VOCAB, NEMBED_BASE

DEFAULT_PRECISION = os.environ.get("CCSMETH_B200_PRECISION", "fp16c8")


class Attention(nn.Module):
    """Parameter container with the reference's names (utils/attention.py:39-46); all bias-free."""

    def __init__(self, query_size, key_size, hidden_size=128):
        super().__init__()
        self.hidden_size = hidden_size
        self.Wa = nn.Linear(query_size, hidden_size, bias=False)
        self.Ua = nn.Linear(key_size, hidden_size, bias=False)
        self.va = nn.Linear(hidden_size, 1, bias=False)


class _NativeModule(nn.Module):
    """Owns the libccsm handle; re-packs weights whenever parameters or the device change."""

    _kind = None

    def _native_init(self, precision):
        if precision not in _lib.PREC:
            raise ValueError("precision must be one of %s" % sorted(_lib.PREC))
        self._precision = precision
        self._handle = None
        self._handle_key = None
        self._dirty = True
        self._h0_mode = "reference"
        self._h0_seed = 1234
        self._h0_mode_dirty = False

    # -- parameter / device changes invalidate the packed weights
    def load_state_dict(self, state_dict, strict=True, **kw):
        # accept the DDP/DataParallel "module." prefix the reference strips on RuntimeError
        # (call_modifications.py:350-358)
        if state_dict and all(k.startswith("module.") for k in state_dict.keys()):
            state_dict = type(state_dict)((k[7:], v) for k, v in state_dict.items())
        r = super().load_state_dict(state_dict, strict=strict, **kw)
        self._dirty = True
        return r

    def _apply(self, fn, *a, **kw):
        r = super()._apply(fn, *a, **kw)
        self._dirty = True
        return r

    def set_precision(self, precision):
        if precision not in _lib.PREC:
            raise ValueError("precision must be one of %s" % sorted(_lib.PREC))
        if precision != self._precision:
            self._precision = precision
            if self._handle is not None and not self._dirty:
                _lib.check(_lib.load().ccsm_set_precision(self._handle, _lib.PREC[precision]))
        return self

    def get_precision(self):
        return self._precision

    def set_h0_mode(self, mode, seed=1234):
        """How ``forward(..., h0=None)`` obtains the GRU initial state:
        "reference" (default) -- the reference's stream (models.py:77-87: torch.randn on the process-wide CPU
                                 generator, per model call, strand 1 then strand 2), drawn ON THE DEVICE bit for bit
                                 (include/ccsm.h CCSM_H0_TORCH_STREAM): the call borrows torch's generator state, the
                                 library draws from it, and the advanced state is handed back to torch -- so
                                 ``torch.manual_seed(s); model(...)`` behaves like the reference and nothing crosses
                                 PCIe.  GRU models; LSTM models fall back to "reference_host";
        "reference_host"      -- the same stream drawn by torch.randn on the host and copied over (12 KB/site);
        "device"              -- N(0,1) drawn inside the feature-packing kernel (Philox, include/ccsm.h
                                 CCSM_H0_DEVICE_RANDOM): same distribution, not the reference's stream;
        "zeros"               -- zero state."""
        if mode not in ("reference", "reference_host", "device", "zeros"):
            raise ValueError("h0 mode must be reference, reference_host, device or zeros")
        self._h0_mode = mode
        self._h0_seed = int(seed)
        self._h0_mode_dirty = True
        return self

    def _stream_h0(self):
        """True when a call with h0=None draws the reference's stream inside the library."""
        return self._h0_mode == "reference" and getattr(self, "rnn_cell", None) == "gru"

    def _host_h0(self):
        """True when a call with h0=None draws the reference's stream with torch.randn on the host."""
        return self._h0_mode == "reference_host" or (self._h0_mode == "reference" and not self._stream_h0())

    def _apply_h0_mode(self, handle):
        if getattr(self, "_h0_mode_dirty", False):
            mode = 1 if self._h0_mode == "device" else (2 if self._stream_h0() else 0)
            _lib.check(_lib.load().ccsm_set_h0_mode(handle, mode, ctypes.c_uint64(self._h0_seed)))
            self._h0_mode_dirty = False

    def set_h0_batching(self, holebatch_sites, batch_size=512):
        """"reference" mode: the next call's sites are these consecutive hole-batches, each cut into model calls of at
        most ``batch_size`` sites like the reference's batch loop (call_modifications.py:177-181).  One-shot."""
        self._h0_batching = (np.ascontiguousarray(holebatch_sites, dtype=np.int64), int(batch_size))

    def _borrow_torch_rng(self, handle, n):
        """Hands torch's CPU generator (MT19937 state + position) to the library's stream and announces the batching."""
        lib = _lib.load()
        st = torch.get_rng_state().numpy()
        left = int.from_bytes(st[8:12].tobytes(), "little", signed=True)
        nxt = int.from_bytes(st[16:24].tobytes(), "little")
        words = np.ascontiguousarray(st[24:24 + 4992].view(np.uint64).astype(np.uint32))
        pos = 624 if left == 1 else nxt
        _lib.check(lib.ccsm_h0_stream_set_state(handle, words.ctypes.data, pos))
        counts, bs = getattr(self, "_h0_batching", None) or (np.array([n], dtype=np.int64), 512)
        self._h0_batching = None
        if int(counts.sum()) != n:
            raise ValueError("set_h0_batching announced %d sites, the call has %d" % (int(counts.sum()), n))
        _lib.check(lib.ccsm_set_h0_batching(handle, counts.ctypes.data, len(counts), bs))
        return st

    def _return_torch_rng(self, handle, st):
        """Reads the advanced generator back and installs it as torch's CPU generator state."""
        words = np.empty(624, dtype=np.uint32)
        pos = ctypes.c_int32(0)
        _lib.check(_lib.load().ccsm_h0_stream_get_state(handle, words.ctypes.data, ctypes.byref(pos)))
        st = st.copy()
        st[24:24 + 4992] = words.astype(np.uint64).view(np.uint8)
        p = pos.value
        st[8:12] = np.frombuffer(int(1 if p == 624 else 625 - p).to_bytes(4, "little", signed=True), dtype=np.uint8)
        st[16:24] = np.frombuffer(int(p).to_bytes(8, "little"), dtype=np.uint8)
        torch.set_rng_state(torch.from_numpy(st))

    def _device_index(self):
        p = next(self.parameters())
        if p.is_cuda:
            return p.device.index if p.device.index is not None else torch.cuda.current_device()
        if not torch.cuda.is_available():
            raise RuntimeError("ccsmeth_b200 needs a CUDA device (sm_100a): there is no CPU fallback path")
        d = self.device if isinstance(self.device, int) else 0
        return d

    def _config(self, dev):
        raise NotImplementedError

    def _ensure_handle(self):
        lib = _lib.load()
        dev = self._device_index()
        key = (dev,)
        if self._handle is not None and (key != self._handle_key):
            lib.ccsm_destroy(self._handle)
            self._handle = None
        if self._handle is None:
            cfg = self._config(dev)
            h = ctypes.c_void_p()
            _lib.check(lib.ccsm_create(ctypes.byref(h), ctypes.byref(cfg)))
            self._handle, self._handle_key, self._dirty = h, key, True
            self._h0_mode_dirty = True
        if self._dirty:
            _lib.check(lib.ccsm_set_precision(self._handle, _lib.PREC[self._precision]))
            for k, v in self.state_dict().items():
                if k.endswith("num_batches_tracked"):
                    continue  # BatchNorm bookkeeping, not a weight
                a = np.ascontiguousarray(v.detach().to("cpu", torch.float32).numpy())
                shp = (ctypes.c_int64 * a.ndim)(*a.shape)
                _lib.check(lib.ccsm_set_weight(self._handle, k.encode(), a.ctypes.data_as(ctypes.c_void_p), shp, a.ndim))
            _lib.check(lib.ccsm_finalize(self._handle))
            self._dirty = False
        return self._handle, dev

    def __del__(self):
        try:
            if getattr(self, "_handle", None) is not None:
                _lib.load().ccsm_destroy(self._handle)
                self._handle = None
        except Exception:
            pass

    def get_model_type(self):
        return self.model_type


def _dev_f32(t, device, shape=None):
    t = torch.as_tensor(t)
    if t.dtype != torch.float32 or t.device != device:
        t = t.to(device=device, dtype=torch.float32)
    if shape is not None:
        t = t.reshape(shape)
    return t.contiguous()


class _ReadsMixin:
    """reads in, calls out: the device feature extractor + forward + MM/ML values (include/ccsm.h ccsm_reads_*),
    shared by every two-strand model class."""

    # ---- reads in, calls out: device feature extraction (include/ccsm.h ccsm_reads_*)
    def extract_reads(self, batch, opts):
        """Uploads a ``extract_features.ReadBatch`` and builds its site list on the device
        (replaces extract_features_from_double_strand_read, reference extract_features.py:261-406).
        ``opts`` = ``extract_features.extract_opts(args, motifs)``.  Returns the number of candidate sites."""
        handle, _ = self._ensure_handle()
        motifs = opts["motifs"]
        o = _lib.ExtractOpts(opts["mod_loc"], opts["norm"], opts["decode"], len(motifs), len(motifs[0]),
                             "".join(motifs).encode("ascii"))
        blob = np.ascontiguousarray(batch.blob, dtype=np.uint8)
        descs = np.ascontiguousarray(batch.descs)
        n_sites = ctypes.c_int64(0)
        _lib.check(_lib.load().ccsm_reads_extract_host(handle, ctypes.byref(o), blob.ctypes.data, blob.size,
                                                       descs.ctypes.data, len(descs), ctypes.byref(n_sites)))
        self._n_sites = int(n_sites.value)
        return self._n_sites

    def reads_sites(self):
        """(site_read, site_loc) int32 arrays of the resident batch: index into the batch's reads, and the 0-based
        position of the called base in the forward read."""
        handle, _ = self._ensure_handle()
        sr = np.empty(self._n_sites, dtype=np.int32)
        sl = np.empty(self._n_sites, dtype=np.int32)
        _lib.check(_lib.load().ccsm_reads_sites(handle, sr.ctypes.data, sl.ctypes.data))
        return sr, sl

    def reads_features(self, s0=0, cn=None):
        """The reference's feature tensors of sites [s0, s0+cn) of the resident batch, as device tensors
        {kmer, kpass, ipd, pw, kmer2, kpass2, ipd2, pw2 (, sns, sns2)} -- what _batch_feature_list2s + the
        FloatTensor stacking would hand to forward (reference call_modifications.py:73-123,201-208)."""
        handle, dev = self._ensure_handle()
        device = torch.device("cuda", dev)
        cn = self._n_sites - s0 if cn is None else cn
        out, strands = {}, []
        for sfx in ("", "2"):
            s = _lib.Strand()
            for name, key, w in (("kmer", "kmer", self.seq_len), ("kpass", "kpass", self.seq_len),
                                 ("ipd_means", "ipd", self.seq_len), ("pw_means", "pw", self.seq_len), ("sns", "sns", 4)):
                if (key == "kpass" and not self.is_npass) or (key == "sns" and not self.is_sn):
                    continue
                t = torch.empty((cn, w), dtype=torch.float32, device=device)
                out[key + sfx] = t
                setattr(s, name, t.data_ptr())
            strands.append(s)
        stream = torch.cuda.current_stream(device).cuda_stream
        _lib.check(_lib.load().ccsm_reads_features(handle, s0, cn, ctypes.byref(strands[0]), ctypes.byref(strands[1]),
                                                   ctypes.c_void_p(stream)))
        return out

    def reads_forward(self, h0=None, want_probs=True):
        """Features -> forward -> per-site outputs for the resident batch (include/ccsm.h ccsm_reads_forward_host).
        Returns a dict of numpy arrays: probs (n, 2) float32 [if want_probs], prob1 (n,) float32 =
        round(p1/(p0+p1), 6), mm (n,) int32 MM-tag deltas, ml (n,) uint8 ML bytes."""
        handle, _ = self._ensure_handle()
        n = self._n_sites
        self._apply_h0_mode(handle)
        cell = getattr(self, "rnn_cell", None)
        if cell is None:
            h0 = None  # no recurrent state (transformer encoder)
        elif cell == "lstm":
            if h0 is not None or self._h0_mode in ("reference", "reference_host"):
                raise ValueError("LSTM models in the reads pipeline take their initial state from the library: use "
                                 "h0 mode 'device' or 'zeros' (the cell state is zero)")
        elif h0 is None and self._host_h0():
            h0 = (self.init_hidden(n, self.num_layers, self.hidden_size),
                  self.init_hidden(n, self.num_layers, self.hidden_size))
        borrowed = self._borrow_torch_rng(handle, n) if (h0 is None and cell == "gru" and self._stream_h0() and n > 0) else None
        pa = pb = None
        if h0 is not None:
            hshape = (2 * self.num_layers, n, self.hidden_size)
            h0a, h0b = _dev_f32(h0[0], torch.device("cpu"), hshape), _dev_f32(h0[1], torch.device("cpu"), hshape)
            pa, pb = h0a.data_ptr(), h0b.data_ptr()
        res = {"prob1": np.empty(n, dtype=np.float32), "mm": np.empty(n, dtype=np.int32),
               "ml": np.empty(n, dtype=np.uint8)}
        probs = np.empty((n, self.num_classes), dtype=np.float32) if want_probs else None
        if n > 0:
            _lib.check(_lib.load().ccsm_reads_forward_host(handle, pa, pb, None,
                                                           probs.ctypes.data if want_probs else None,
                                                           res["prob1"].ctypes.data, res["mm"].ctypes.data,
                                                           res["ml"].ctypes.data))
        if borrowed is not None:
            self._return_torch_rng(handle, borrowed)
        if want_probs:
            res["probs"] = probs
        return res


class ModelAttRNN(_ReadsMixin, _NativeModule):
    """Drop-in for the reference ``ModelAttRNN`` with ``model_type="attbigru2s"`` (models.py:17-150)."""

    def __init__(self, seq_len=21, num_layers=3, num_classes=2, dropout_rate=0.5, hidden_size=256,
                 is_npass=True, is_sn=False, is_map=False, is_stds=False, model_type="attbigru2s", device=0,
                 precision=None):
        super().__init__()
        if model_type not in ("attbigru2s", "attbilstm2s"):
            raise ValueError("--model_type not set right! (ccsmeth_b200 implements attbigru2s and attbilstm2s)")
        self.model_type = model_type
        self.device = device
        self.seq_len, self.num_layers, self.num_classes, self.hidden_size = seq_len, num_layers, num_classes, hidden_size
        self.n_embed = NEMBED_BASE
        self.is_stds, self.is_npass, self.is_sn, self.is_map = is_stds, is_npass, is_sn, is_map
        self.feas_ccs = 2 + (2 if is_stds else 0) + (1 if is_npass else 0) + (4 if is_sn else 0) + (1 if is_map else 0)
        # the LSTM variant (reference models.py:48-51) has no shipped checkpoint; it runs on the fp32 kernels
        self.rnn_cell = "lstm" if model_type == "attbilstm2s" else "gru"
        # parameter containers, never called: same names/shapes as the reference state_dict
        self.embed = nn.Embedding(N_VOCAB, self.n_embed)
        rnn_cls = nn.LSTM if self.rnn_cell == "lstm" else nn.GRU
        self.rnn = rnn_cls(self.n_embed + self.feas_ccs, hidden_size, num_layers, dropout=dropout_rate,
                           batch_first=True, bidirectional=True)
        self._att3 = Attention(hidden_size * 2, hidden_size * 2, hidden_size)
        self.dropout1 = nn.Dropout(p=dropout_rate)  # identity at inference; kept for interface parity
        self.fc1 = nn.Linear(hidden_size * 2 * 2, num_classes)
        self.init_weights()
        self.requires_grad_(False)
        self._native_init("fp32" if self.rnn_cell == "lstm" else (precision or DEFAULT_PRECISION))

    def init_weights(self):  # reference models.py:71-75
        nn.init.uniform_(self.embed.weight, -0.1, 0.1)
        nn.init.zeros_(self.fc1.bias)
        nn.init.uniform_(self.fc1.weight, -0.1, 0.1)

    def init_hidden(self, batch_size, num_layers, hidden_size):
        """Same draw as the reference (models.py:77-87): CPU default generator, then moved to the device.
        LSTM: (h0, c0), h0 drawn first."""
        h0 = torch.randn(num_layers * 2, batch_size, hidden_size)
        if self.rnn_cell == "lstm":
            return h0, torch.randn(num_layers * 2, batch_size, hidden_size)
        return h0

    def _config(self, dev):
        flags = (_lib.FEAT_NPASS if self.is_npass else 0) | (_lib.FEAT_STDS if self.is_stds else 0) | \
                (_lib.FEAT_SN if self.is_sn else 0) | (_lib.FEAT_MAP if self.is_map else 0) | \
                (_lib.CELL_LSTM if self.rnn_cell == "lstm" else 0)
        return _lib.Config(_lib.KIND_ATT2S, self.seq_len, self.num_layers, self.hidden_size, self.num_classes,
                           N_VOCAB, self.n_embed, flags, _lib.PREC[self._precision], dev)

    def _strand(self, device, n, kmer, kpass, ipd_means, ipd_stds, pw_means, pw_stds, sns, maps, keep):
        L = self.seq_len
        s = _lib.Strand()

        def put(name, t, shape):
            t = _dev_f32(t, device, shape)
            keep.append(t)
            setattr(s, name, t.data_ptr())

        put("kmer", kmer, (n, L))
        put("ipd_means", ipd_means, (n, L))
        put("pw_means", pw_means, (n, L))
        if self.is_npass:
            put("kpass", kpass, (n, L))
        if self.is_stds:
            put("ipd_stds", ipd_stds, (n, L))
            put("pw_stds", pw_stds, (n, L))
        if self.is_sn:
            put("sns", sns, (n, 4))
        if self.is_map:
            put("maps", maps, (n, L))
        return s

    def forward(self, kmer, kpass, ipd_means, ipd_stds, pw_means, pw_stds, sns, maps,
                kmer2, kpass2, ipd_means2, ipd_stds2, pw_means2, pw_stds2, sns2, maps2, h0=None):
        handle, dev = self._ensure_handle()
        device = torch.device("cuda", dev)
        n = int(torch.as_tensor(kmer).reshape(-1, self.seq_len).shape[0])
        self._apply_h0_mode(handle)
        if h0 is None and self._host_h0():
            h0 = (self.init_hidden(n, self.num_layers, self.hidden_size),
                  self.init_hidden(n, self.num_layers, self.hidden_size))
        keep = []
        fwd = self._strand(device, n, kmer, kpass, ipd_means, ipd_stds, pw_means, pw_stds, sns, maps, keep)
        rev = self._strand(device, n, kmer2, kpass2, ipd_means2, ipd_stds2, pw_means2, pw_stds2, sns2, maps2, keep)
        hshape = (2 * self.num_layers, n, self.hidden_size)
        logits = torch.empty((n, self.num_classes), dtype=torch.float32, device=device)
        probs = torch.empty((n, self.num_classes), dtype=torch.float32, device=device)
        if self.rnn_cell == "lstm":
            # h0 = ((h0, c0) of strand 1, (h0, c0) of strand 2), as init_hidden returns them
            ptrs = [None] * 4
            if h0 is not None:
                ts = [_dev_f32(t, device, hshape) for pair in h0 for t in pair]
                keep += ts
                ptrs = [t.data_ptr() for t in ts]
            if n > 0:
                stream = torch.cuda.current_stream(device).cuda_stream
                _lib.check(_lib.load().ccsm_forward_att2s_lstm(handle, n, ctypes.byref(fwd), ctypes.byref(rev), *ptrs,
                                                               logits.data_ptr(), probs.data_ptr(), ctypes.c_void_p(stream)))
            return logits, probs
        if h0 is not None:
            h0a, h0b = _dev_f32(h0[0], device, hshape), _dev_f32(h0[1], device, hshape)
            pa, pb = h0a.data_ptr(), h0b.data_ptr()
        else:
            pa = pb = None  # library draws (device mode) or uses zeros
        if n > 0:
            stream = torch.cuda.current_stream(device).cuda_stream
            borrowed = self._borrow_torch_rng(handle, n) if (h0 is None and self._stream_h0()) else None
            _lib.check(_lib.load().ccsm_forward_att2s(handle, n, ctypes.byref(fwd), ctypes.byref(rev),
                                                      pa, pb, logits.data_ptr(),
                                                      probs.data_ptr(), ctypes.c_void_p(stream)))
            if borrowed is not None:
                self._return_torch_rng(handle, borrowed)
        return logits, probs

    def profile(self, on=True):
        """Enable/disable per-kernel CUDA-event timing inside the library (include/ccsm.h ccsm_profile_*)."""
        handle, _ = self._ensure_handle()
        _lib.check(_lib.load().ccsm_profile_enable(handle, 1 if on else 0))

    def profile_read(self):
        """Returns {class: (ms, sites, launches)} accumulated since the last read."""
        handle, _ = self._ensure_handle()
        ms = (ctypes.c_double * 6)()
        units = (ctypes.c_double * 6)()
        launches = (ctypes.c_int64 * 6)()
        _lib.check(_lib.load().ccsm_profile_read(handle, ms, units, launches, 6))
        names = ("prep", "gru_l0", "gru_ln", "att_head", "read_scan", "window_gather")
        return {n: (ms[i], units[i], int(launches[i])) for i, n in enumerate(names)}

    def forward_host(self, feats, h0=None):
        """Host-buffer entry (include/ccsm.h ccsm_forward_att2s_host): `feats` maps the live tensor names
        (kmer, kpass, ipd, pw and the same with a '2' suffix) to float32 CPU tensors / numpy arrays of shape
        (n, seq_len).  Copies, forward and result read-back are pipelined inside the library.  Returns
        CPU tensors (logits, probs)."""
        if self.rnn_cell == "lstm":
            raise NotImplementedError("the host-buffer entry implements the GRU model; use forward() for attbilstm2s")
        handle, dev = self._ensure_handle()
        L = self.seq_len
        cpu = torch.device("cpu")
        n = int(torch.as_tensor(feats["kmer"]).reshape(-1, L).shape[0])
        keep = []
        strands = []
        for sfx in ("", "2"):
            s = _lib.Strand()
            for name, key in (("kmer", "kmer"), ("kpass", "kpass"), ("ipd_means", "ipd"), ("pw_means", "pw"),
                              ("ipd_stds", "ipd_sd"), ("pw_stds", "pw_sd"), ("sns", "sns"), ("maps", "maps")):
                if key + sfx in feats and feats[key + sfx] is not None:
                    t = _dev_f32(feats[key + sfx], cpu)
                    keep.append(t)
                    setattr(s, name, t.data_ptr())
            strands.append(s)
        self._apply_h0_mode(handle)
        if h0 is None and self._host_h0():
            h0 = (self.init_hidden(n, self.num_layers, self.hidden_size),
                  self.init_hidden(n, self.num_layers, self.hidden_size))
        hshape = (2 * self.num_layers, n, self.hidden_size)
        if h0 is not None:
            h0a, h0b = _dev_f32(h0[0], cpu, hshape), _dev_f32(h0[1], cpu, hshape)
            pa, pb = h0a.data_ptr(), h0b.data_ptr()
        else:
            pa = pb = None
        logits = torch.empty((n, self.num_classes), dtype=torch.float32)
        probs = torch.empty((n, self.num_classes), dtype=torch.float32)
        if n > 0:
            borrowed = self._borrow_torch_rng(handle, n) if (h0 is None and self._stream_h0()) else None
            _lib.check(_lib.load().ccsm_forward_att2s_host(handle, n, ctypes.byref(strands[0]), ctypes.byref(strands[1]),
                                                           pa, pb, logits.data_ptr(), probs.data_ptr()))
            if borrowed is not None:
                self._return_torch_rng(handle, borrowed)
        return logits, probs



class ModelAttRNN2(ModelAttRNN):
    """Drop-in for the reference ``ModelAttRNN2`` (``model_type`` "attbigru2s2" / "attbilstm2s2", models.py:221-382):
    bases, kinetics (as integers: use ``--norm none``) and pass counts are embedded, the head is a two-layer
    classifier.  No checkpoint ships for it; fp32 kernels; same 16-tensor ``forward`` as ``ModelAttRNN``."""

    def __init__(self, seq_len=21, num_layers=3, num_classes=2, dropout_rate=0.5, hidden_size=256,
                 is_npass=True, is_sn=False, is_map=False, is_stds=False, model_type="attbigru2s2", device=0,
                 precision=None):
        nn.Module.__init__(self)
        if model_type not in ("attbigru2s2", "attbilstm2s2"):
            raise ValueError("--model_type not set right!")
        if is_sn or is_map or is_stds:
            raise ValueError("ccsmeth_b200 implements ModelAttRNN2 with the kinetics and pass-count features only")
        self.model_type = model_type
        self.device = device
        self.seq_len, self.num_layers, self.num_classes, self.hidden_size = seq_len, num_layers, num_classes, hidden_size
        self.n_embed = NEMBED_BASE
        self.is_stds, self.is_npass, self.is_sn, self.is_map = is_stds, is_npass, is_sn, is_map
        self.rnn_cell = "lstm" if model_type == "attbilstm2s2" else "gru"
        self.feas_ccs = 2 + (1 if is_npass else 0)
        self.nembed_all = NEMBED_BASE + 2 * 8 + (4 if is_npass else 0)  # NEMBED_KINETICS = 8, NEMBED_PASSES = 4
        # parameter containers in the reference's construction order (same state_dict keys, same generator draws)
        self.seq_embed = nn.Embedding(N_VOCAB, NEMBED_BASE)
        self.ipd_embed = nn.Embedding(952 + 1, 8)   # MAX_KINETICS + 1
        self.pw_embed = nn.Embedding(952 + 1, 8)
        if is_npass:
            self.npass_embed = nn.Embedding(30 + 1, 4)  # MAX_PASSES + 1
        rnn_cls = nn.LSTM if self.rnn_cell == "lstm" else nn.GRU
        self.rnn = rnn_cls(self.nembed_all, hidden_size, num_layers, dropout=dropout_rate, batch_first=True,
                           bidirectional=True)
        self._att3 = Attention(hidden_size * 2, hidden_size * 2, hidden_size)
        self.classifier = nn.Sequential(nn.Linear(hidden_size * 4, hidden_size * 4), nn.ReLU(), nn.Dropout(p=dropout_rate),
                                        nn.Linear(hidden_size * 4, num_classes))
        self.init_weights()
        self.requires_grad_(False)
        self._native_init("fp32")

    def init_weights(self):  # reference models.py:285-303
        for emb in (self.seq_embed, self.ipd_embed, self.pw_embed):
            nn.init.uniform_(emb.weight, -0.1, 0.1)
        if self.is_npass:
            nn.init.uniform_(self.npass_embed.weight, -0.1, 0.1)
        for m in self.classifier.modules():
            if isinstance(m, nn.Linear):
                nn.init.uniform_(m.weight, -0.1, 0.1)
                nn.init.zeros_(m.bias)

    def _config(self, dev):
        flags = _lib.MODEL_2S2 | (_lib.FEAT_NPASS if self.is_npass else 0) | (_lib.CELL_LSTM if self.rnn_cell == "lstm" else 0)
        return _lib.Config(_lib.KIND_ATT2S, self.seq_len, self.num_layers, self.hidden_size, self.num_classes,
                           N_VOCAB, self.n_embed, flags, _lib.PREC["fp32"], dev)

    def set_precision(self, precision):  # fp32 kernels only
        return self


class _EmbedBlockPlus(nn.Module):  # parameter container, reference models.py:153-170
    def __init__(self, d_model):
        super().__init__()
        self.conv_embed = nn.Sequential(nn.Conv1d(d_model, d_model, 3, 1, 1, bias=False), nn.BatchNorm1d(d_model),
                                        nn.ReLU(inplace=True), nn.MaxPool1d(3, 1, 1))


class _SrcEmbed(nn.Module):  # parameter container, reference models.py:173-218
    def __init__(self, input_dim, d_model, block_plus=1):
        super().__init__()
        self.conv_embed = nn.Sequential(nn.Conv1d(input_dim, d_model // 2, 3, 1, 1, bias=False), nn.BatchNorm1d(d_model // 2),
                                        nn.ReLU(inplace=True), nn.MaxPool1d(3, 1, 1),
                                        nn.Conv1d(d_model // 2, d_model, 3, 1, 1, bias=False), nn.BatchNorm1d(d_model),
                                        nn.ReLU(inplace=True), nn.MaxPool1d(3, 1, 1))
        self.conv_embed_plus = nn.Sequential(*[_EmbedBlockPlus(d_model) for _ in range(block_plus)])


class _PositionalEmbedding(nn.Module):  # parameter container, reference models.py:437-448
    def __init__(self, seq_len, d_model):
        super().__init__()
        self.pos_embed = nn.Embedding(seq_len, d_model)


class ModelTransEnc(_ReadsMixin, _NativeModule):
    """Drop-in for the reference ``ModelTransEnc`` (``model_type="transencoder2s"``, models.py:451-620): integer
    embeddings, SrcEmbed conv stack, learned positions, post-norm transformer encoder, mean pooling, classifier.
    No checkpoint ships for it; fp32 kernels; same 16-tensor ``forward`` (there is no recurrent state, so no h0)."""

    def __init__(self, seq_len=21, num_layers=6, num_classes=2, dropout_rate=0.5, d_model=256, nhead=4, dim_ff=512,
                 is_npass=True, is_sn=False, is_map=False, is_stds=False, model_type="transencoder2s", device=0):
        super().__init__()
        if model_type != "transencoder2s":
            raise ValueError("--model_type not set right!")
        if is_sn or is_map or is_stds:
            raise ValueError("ccsmeth_b200 implements ModelTransEnc with the kinetics and pass-count features only")
        self.model_type, self.device = model_type, device
        self.seq_len, self.num_layers, self.num_classes, self.d_model = seq_len, num_layers, num_classes, d_model
        self.nhead, self.dim_ff = nhead, dim_ff
        self.hidden_size = d_model
        self.rnn_cell = None  # no recurrent state
        self.n_embed = NEMBED_BASE
        self.is_stds, self.is_npass, self.is_sn, self.is_map = is_stds, is_npass, is_sn, is_map
        self.nembed_all = NEMBED_BASE + 2 * 8 + (4 if is_npass else 0)
        self.seq_embed = nn.Embedding(N_VOCAB, NEMBED_BASE)
        self.ipd_embed = nn.Embedding(952 + 1, 8)
        self.pw_embed = nn.Embedding(952 + 1, 8)
        if is_npass:
            self.npass_embed = nn.Embedding(30 + 1, 4)
        self.trans_input = _SrcEmbed(self.nembed_all, d_model, block_plus=1)
        self.pos_encoder = _PositionalEmbedding(seq_len, d_model)
        layer = nn.TransformerEncoderLayer(d_model=d_model, nhead=nhead, dim_feedforward=dim_ff, dropout=dropout_rate,
                                           batch_first=True)
        self.transformer_encoder = nn.TransformerEncoder(layer, num_layers)
        self.classifier = nn.Sequential(nn.Linear(d_model * 2, d_model * 2), nn.ReLU(), nn.Dropout(p=dropout_rate),
                                        nn.Linear(d_model * 2, num_classes))
        self.requires_grad_(False)
        self._native_init("fp32")

    def _config(self, dev):
        flags = _lib.MODEL_TRANSENC | (_lib.FEAT_NPASS if self.is_npass else 0) | (self.nhead << 8)
        return _lib.Config(_lib.KIND_ATT2S, self.seq_len, self.num_layers, self.d_model, self.num_classes,
                           N_VOCAB, self.n_embed, flags, _lib.PREC["fp32"], dev)

    def set_precision(self, precision):  # fp32 kernels only
        return self

    def forward(self, kmer, kpass, ipd_means, ipd_stds, pw_means, pw_stds, sns, maps,
                kmer2, kpass2, ipd_means2, ipd_stds2, pw_means2, pw_stds2, sns2, maps2, has_mask=False):
        handle, dev = self._ensure_handle()
        device = torch.device("cuda", dev)
        n = int(torch.as_tensor(kmer).reshape(-1, self.seq_len).shape[0])
        keep = []
        fwd = ModelAttRNN._strand(self, device, n, kmer, kpass, ipd_means, ipd_stds, pw_means, pw_stds, sns, maps, keep)
        rev = ModelAttRNN._strand(self, device, n, kmer2, kpass2, ipd_means2, ipd_stds2, pw_means2, pw_stds2, sns2, maps2, keep)
        logits = torch.empty((n, self.num_classes), dtype=torch.float32, device=device)
        probs = torch.empty((n, self.num_classes), dtype=torch.float32, device=device)
        if n > 0:
            stream = torch.cuda.current_stream(device).cuda_stream
            _lib.check(_lib.load().ccsm_forward_att2s(handle, n, ctypes.byref(fwd), ctypes.byref(rev), None, None,
                                                      logits.data_ptr(), probs.data_ptr(), ctypes.c_void_p(stream)))
        return logits, probs


class AggrAttRNN(_NativeModule):
    """Drop-in for the reference ``AggrAttRNN`` (``model_type`` "attbigru" / "attbilstm", models.py:625-694):
    regression over 11 neighbouring CpG sites, raw fc1 output (no softmax).  The shipped configuration (GRU, hidden 32,
    20 bins, one layer) runs one fused kernel; other shapes and the LSTM cell run the layer-by-layer fp32 kernels."""

    def __init__(self, seq_len=11, num_layers=1, num_classes=1, dropout_rate=0.5, hidden_size=32, binsize=20,
                 model_type="attbigru", device=0, precision=None):
        super().__init__()
        if model_type not in ("attbigru", "attbilstm"):
            raise ValueError("--model_type not set right!")
        self.model_type = model_type
        self.device = device
        self.seq_len, self.num_layers, self.num_classes, self.hidden_size = seq_len, num_layers, num_classes, hidden_size
        self.binsize = binsize
        self.feas_ccs = binsize + 1
        self.rnn_cell = "lstm" if model_type == "attbilstm" else "gru"
        rnn = nn.LSTM if self.rnn_cell == "lstm" else nn.GRU
        self.rnn = rnn(self.feas_ccs, hidden_size, num_layers, dropout=0, batch_first=True, bidirectional=True)
        self._att3 = Attention(hidden_size * 2, hidden_size * 2, hidden_size)
        self.dropout1 = nn.Dropout(p=dropout_rate)
        self.fc1 = nn.Linear(hidden_size * 2, num_classes)
        self.requires_grad_(False)
        self._native_init("fp32")  # K=21/32 contractions are not tensor-core shaped

    def init_hidden(self, batch_size, num_layers, hidden_size):  # reference models.py:661-671 (h0, then c0)
        h0 = torch.randn(num_layers * 2, batch_size, hidden_size)
        if self.rnn_cell == "lstm":
            return h0, torch.randn(num_layers * 2, batch_size, hidden_size)
        return h0

    def _config(self, dev):
        flags = self.binsize | (_lib.AGGR_LSTM if self.rnn_cell == "lstm" else 0)
        return _lib.Config(_lib.KIND_AGGR, self.seq_len, self.num_layers, self.hidden_size, self.num_classes,
                           0, 0, flags, _lib.PREC["fp32"], dev)

    def forward(self, offsets, histos, h0=None):
        """h0: the initial state -- a (2*layers, n, hidden) tensor, for the LSTM cell an (h0, c0) pair; None draws it
        like the reference's init_hidden."""
        handle, dev = self._ensure_handle()
        device = torch.device("cuda", dev)
        L = self.seq_len
        offsets = _dev_f32(offsets, device, (-1, L))
        n = int(offsets.shape[0])
        histos = _dev_f32(histos, device, (n, L, self.binsize))
        if h0 is None:
            h0 = self.init_hidden(n, self.num_layers, self.hidden_size)
        c0 = None
        if self.rnn_cell == "lstm":
            if not isinstance(h0, (tuple, list)) or len(h0) != 2:
                raise ValueError("attbilstm takes the initial state as an (h0, c0) pair")
            h0, c0 = h0
            c0 = _dev_f32(c0, device, (2 * self.num_layers, n, self.hidden_size))
        h0 = _dev_f32(h0, device, (2 * self.num_layers, n, self.hidden_size))
        out = torch.empty((n, self.num_classes), dtype=torch.float32, device=device)
        if n > 0:
            stream = torch.cuda.current_stream(device).cuda_stream
            if c0 is not None:
                _lib.check(_lib.load().ccsm_forward_aggr_lstm(handle, n, offsets.data_ptr(), histos.data_ptr(), h0.data_ptr(),
                                                              c0.data_ptr(), out.data_ptr(), ctypes.c_void_p(stream)))
            else:
                _lib.check(_lib.load().ccsm_forward_aggr(handle, n, offsets.data_ptr(), histos.data_ptr(), h0.data_ptr(),
                                                         out.data_ptr(), ctypes.c_void_p(stream)))
        return out

    def fused(self):
        """True for the configuration the one-launch kernel implements (include/ccsm.h ccsm_forward_aggr_sites)."""
        return self.rnn_cell == "gru" and self.hidden_size == 32 and self.num_layers == 1 and self.binsize == 20 \
            and self.num_classes == 1 and self.seq_len <= 32

    def forward_sites(self, positions, site_histos, h0=None, only_close=False):
        """The forward from per-site rows (include/ccsm.h ccsm_forward_aggr_sites): positions (n,) and histograms
        (n, bins) of the region's sites in order; the kernel forms every site's window of ``seq_len`` neighbours itself
        (what reference call_mods_freq_bam.py:265-293 materialises on the host)."""
        handle, dev = self._ensure_handle()
        device = torch.device("cuda", dev)
        pos = torch.as_tensor(np.ascontiguousarray(positions, dtype=np.int64)).to(device)
        n = int(pos.shape[0])
        histos = _dev_f32(site_histos, device, (n, self.binsize))
        if h0 is None:
            h0 = self.init_hidden(n, self.num_layers, self.hidden_size)
        h0 = _dev_f32(h0, device, (2 * self.num_layers, n, self.hidden_size))
        out = torch.empty((n, self.num_classes), dtype=torch.float32, device=device)
        if n > 0:
            stream = torch.cuda.current_stream(device).cuda_stream
            _lib.check(_lib.load().ccsm_forward_aggr_sites(handle, n, pos.data_ptr(), histos.data_ptr(), 1 if only_close else 0,
                                                           h0.data_ptr(), out.data_ptr(), ctypes.c_void_p(stream)))
        return out

    # ---- call_freqb on the device: one region's pileup -> per-site frequencies (include/ccsm.h ccsm_pileup_*)
    def pileup_begin(self, refpos, ptr, ml, hap=None, call_mode="aggregate", cov_cf=4, prob_cf=0.0, no_amb_cov=False,
                     no_hap=False, only_close=False):
        """Uploads a region's pileup in CSR form (site i at reference position refpos[i] is covered by entries
        ptr[i]:ptr[i+1] of `ml` / `hap`) and returns n_high = (all, hp1, hp2): how many sites of each read group
        go through the aggregate model -- the n of the h0 tensors ``pileup_finish`` takes."""
        handle, _ = self._ensure_handle()
        self._pu = (np.ascontiguousarray(refpos, dtype=np.int64), np.ascontiguousarray(ptr, dtype=np.int64),
                    np.ascontiguousarray(ml, dtype=np.uint8),
                    None if hap is None else np.ascontiguousarray(hap, dtype=np.uint8))
        pos, ptr, ml, hap = self._pu
        o = _lib.PileupOpts({"count": 0, "aggregate": 1}[call_mode], int(cov_cf), float(prob_cf), int(bool(no_amb_cov)),
                            int(bool(no_hap)), 0, int(bool(only_close)))
        n_high = (ctypes.c_int64 * 3)()
        _lib.check(_lib.load().ccsm_pileup_begin_host(handle, ctypes.byref(o), len(pos), pos.ctypes.data, ptr.ctypes.data,
                                                      ml.ctypes.data, hap.ctypes.data if hap is not None else None, n_high))
        self._pu_n = len(pos)
        return tuple(int(v) for v in n_high)

    def pileup_finish(self, h0=(None, None, None), with_kind=False):
        """h0: per read group (all, hp1, hp2) the initial state for the sites ``pileup_begin`` counted, or None (zeros);
        for the LSTM cell each entry is an (h0, c0) pair.
        -> (cov (3, n) int32, cnt_mod (3, n) float64, freq (3, n) float64[, kind (3, n) uint8]), rows = all reads /
        haplotype 1 / 2; cov == -1 marks "no call of this group at this site" (the reference's None); kind: see
        include/ccsm.h (which Python / NumPy value types the reference would hold)."""
        handle, _ = self._ensure_handle()
        n = self._pu_n
        cov = np.full((3, n), -1, dtype=np.int32)
        cnt = np.zeros((3, n), dtype=np.float64)
        freq = np.zeros((3, n), dtype=np.float64)
        kind = np.zeros((3, n), dtype=np.uint8)
        cpu = torch.device("cpu")
        if self.rnn_cell == "lstm":
            hs = [None if h is None else _dev_f32(h[0], cpu) for h in h0]
            cs = [None if h is None else _dev_f32(h[1], cpu) for h in h0]
            ptrs = lambda ts: (ctypes.c_void_p * 3)(*[None if t is None else t.data_ptr() for t in ts])
            _lib.check(_lib.load().ccsm_pileup_finish_lstm_host(handle, ptrs(hs), ptrs(cs), cov.ctypes.data, cnt.ctypes.data,
                                                                freq.ctypes.data, kind.ctypes.data))
        else:
            hs = [None if h is None else _dev_f32(h, cpu) for h in h0]
            _lib.check(_lib.load().ccsm_pileup_finish_host(handle, *[None if h is None else h.data_ptr() for h in hs],
                                                           cov.ctypes.data, cnt.ctypes.data, freq.ctypes.data,
                                                           kind.ctypes.data))
        return (cov, cnt, freq, kind) if with_kind else (cov, cnt, freq)
