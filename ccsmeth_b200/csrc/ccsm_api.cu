// C-ABI layer of libccsm.so (declared in include/ccsm.h).
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "ccsm_internal.h"

namespace ccsm {

static thread_local char g_err[1024] = "";
std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int DevBuf::reserve(size_t bytes) {
  if (bytes <= cap) return CCSM_OK;
  if (p) {
    cudaError_t e = cudaFree(p);
    p = nullptr;
    cap = 0;
    if (e != cudaSuccess) {
      set_error("cudaFree failed: %s", cudaGetErrorString(e));
      return CCSM_ECUDA;
    }
  }
  cudaError_t e = cudaMalloc(&p, bytes);
  if (e != cudaSuccess) {
    p = nullptr;
    set_error("cudaMalloc(%zu bytes) failed: %s", bytes, cudaGetErrorString(e));
    (void)cudaGetLastError();
    return e == cudaErrorMemoryAllocation ? CCSM_ENOMEM : CCSM_ECUDA;
  }
  cap = bytes;
  return CCSM_OK;
}

void DevBuf::release() {
  if (p) cudaFree(p);
  p = nullptr;
  cap = 0;
}

static bool is_tc(int prec) { return prec != CCSM_PREC_FP32; }

}  // namespace ccsm

using namespace ccsm;

extern "C" {

int ccsm_abi_version(void) { return CCSM_ABI_VERSION; }
const char* ccsm_last_error(void) { return g_err; }
int64_t ccsm_kernel_launches(void) { return g_launches.load(); }

int ccsm_create(ccsm_model** out, const ccsm_config* cfg) {
  if (!out || !cfg) {
    set_error("ccsm_create: null argument");
    return CCSM_EINVAL;
  }
  *out = nullptr;
  if (cfg->kind != CCSM_KIND_ATT2S && cfg->kind != CCSM_KIND_AGGR) {
    set_error("ccsm_create: unknown kind %d", cfg->kind);
    return CCSM_EINVAL;
  }
  if (cfg->seq_len < 1 || cfg->seq_len > 32 || (cfg->kind == CCSM_KIND_ATT2S && cfg->seq_len % 2 == 0)) {
    // the reference requires an odd --seq_len (call_modifications.py:500-501)
    set_error("ccsm_create: seq_len %d unsupported (odd, <= 32)", cfg->seq_len);
    return CCSM_EINVAL;
  }
  if (cfg->num_layers < 1 || cfg->hidden < 1 || cfg->hidden % 16 != 0 || cfg->num_classes < 1 || cfg->num_classes > 4) {
    set_error("ccsm_create: layers=%d hidden=%d classes=%d unsupported", cfg->num_layers, cfg->hidden, cfg->num_classes);
    return CCSM_EINVAL;
  }
  if (cfg->precision < CCSM_PREC_FP32 || cfg->precision > CCSM_PREC_FP16C8) {
    set_error("ccsm_create: unknown precision %d", cfg->precision);
    return CCSM_EINVAL;
  }
  int ndev = 0;
  CCSM_CUDA(cudaGetDeviceCount(&ndev));
  if (cfg->device < 0 || cfg->device >= ndev) {
    set_error("ccsm_create: device %d not present (%d visible)", cfg->device, ndev);
    return CCSM_EINVAL;
  }
  cudaDeviceProp prop;
  CCSM_CUDA(cudaGetDeviceProperties(&prop, cfg->device));
  if (prop.major != 10) {
    set_error("ccsm_create: device %d is sm_%d%d; this library is built for sm_100a (B200) only", cfg->device,
              prop.major, prop.minor);
    return CCSM_EUNSUPPORTED;
  }
  ccsm_model* m = new (std::nothrow) ccsm_model();
  if (!m) return CCSM_ENOMEM;
  m->cfg = *cfg;
  if (cfg->kind == CCSM_KIND_ATT2S) {
    m->strands = 2;
    int feas = 2;  // ipd, pw  (reference models.py:39-47)
    if (cfg->feat_flags & CCSM_FEAT_STDS) feas += 2;
    if (cfg->feat_flags & CCSM_FEAT_NPASS) feas += 1;
    if (cfg->feat_flags & CCSM_FEAT_SN) feas += 4;
    if (cfg->feat_flags & CCSM_FEAT_MAP) feas += 1;
    m->in_feat = cfg->n_embed + feas;
    if (cfg->feat_flags & CCSM_CELL_LSTM) {
      m->gates = 4;
      m->cfg.precision = CCSM_PREC_FP32;  // the tensor-core kernels implement the GRU cell only
    }
    if (cfg->feat_flags & CCSM_MODEL_TRANSENC) {
      const int nhead = (cfg->feat_flags >> 8) & 255;
      if ((cfg->feat_flags & (CCSM_FEAT_STDS | CCSM_FEAT_SN | CCSM_FEAT_MAP)) || nhead < 1 || cfg->hidden % nhead != 0 ||
          cfg->hidden % 32 != 0) {
        set_error("ccsm_create: ModelTransEnc needs d_model %% 32 == 0, d_model %% nhead == 0 and only the kinetics / npass features");
        delete m;
        return CCSM_EUNSUPPORTED;
      }
      m->is_trans = true;
      m->nhead = nhead;
      m->in_feat = cfg->n_embed + 2 * 8 + ((cfg->feat_flags & CCSM_FEAT_NPASS) ? 4 : 0);
      m->cfg.precision = CCSM_PREC_FP32;
    } else if (cfg->feat_flags & CCSM_MODEL_2S2) {
      if (cfg->feat_flags & (CCSM_FEAT_STDS | CCSM_FEAT_SN | CCSM_FEAT_MAP)) {
        set_error("ccsm_create: ModelAttRNN2 with --is_stds / --is_sn / --is_map is not implemented");
        delete m;
        return CCSM_EUNSUPPORTED;
      }
      m->is_2s2 = true;
      m->in_feat = cfg->n_embed + 2 * 8 + ((cfg->feat_flags & CCSM_FEAT_NPASS) ? 4 : 0);  // process_utils.py:68-70
      m->cfg.precision = CCSM_PREC_FP32;
    }
  } else {
    m->strands = 1;
    if (cfg->feat_flags & CCSM_AGGR_LSTM) m->gates = 4;  // AggrAttRNN(model_type="attbilstm") (models.py:640-643)
    m->cfg.feat_flags = cfg->feat_flags & 255;  // downstream code reads the bin count here
    if (m->cfg.feat_flags < 1) {
      set_error("ccsm_create: aggregate model needs a bin count in feat_flags (got %d)", m->cfg.feat_flags);
      delete m;
      return CCSM_EINVAL;
    }
    m->in_feat = m->cfg.feat_flags + 1;  // bins + offset (reference models.py:639)
    if (is_tc(cfg->precision)) {
      // the aggregate model's K=21/32 contractions are not tensor-core shaped: fp32 FFMA only
      m->cfg.precision = CCSM_PREC_FP32;
    }
  }
  *out = m;
  return CCSM_OK;
}

void ccsm_destroy(ccsm_model* m) {
  if (!m) return;
  cudaSetDevice(m->cfg.device);
  tc_release(m);
  ex_release(m);
  pu_release(m);
  mt_destroy(m);
  trans_release(m);
  m->aggr_packed.release();
  m->aggr_packed_tiled.release();
  m->aggr_scratch.release();
  for (auto& l : m->fp32.layers) {
    l.w_ih.release(); l.b_ih.release(); l.w_hh.release(); l.b_hh.release();
  }
  m->fp32.ipd_embed.release(); m->fp32.pw_embed.release(); m->fp32.npass_embed.release();
  m->fp32.cls0_w.release(); m->fp32.cls0_b.release();
  m->fp32.embed.release(); m->fp32.Wa.release(); m->fp32.Ua.release(); m->fp32.va.release();
  m->fp32.fc_w.release(); m->fp32.fc_b.release();
  Fp32Workspace& ws = m->ws32;
  ws.x0.release(); ws.gi.release(); ws.gh.release(); ws.h.release(); ws.c.release(); ws.outA.release(); ws.outB.release(); ws.qa.release();
  for (int i = 0; i < 2; ++i) {
    m->stage_in[i].release();
    m->stage_out[i].release();
    if (m->streams[i]) cudaStreamDestroy(m->streams[i]);
    if (m->events[i]) cudaEventDestroy(m->events[i]);
    if (m->done[i]) cudaEventDestroy(m->done[i]);
  }
  delete m;
}

// expected shape of every state_dict key (reference models.py:33,52-64,644-654)
static bool expected_shape(const ccsm_model* m, const std::string& key, std::vector<int64_t>& shp) {
  const int64_t H = m->cfg.hidden, C = m->cfg.num_classes;
  if (m->is_trans) {
    const int64_t d = H, L = m->cfg.seq_len, ff = m->dim_ff;
    if (key == "seq_embed.weight") { shp = {m->cfg.n_vocab, m->cfg.n_embed}; return true; }
    if (key == "ipd_embed.weight" || key == "pw_embed.weight") { shp = {953, 8}; return true; }
    if (key == "npass_embed.weight" && (m->cfg.feat_flags & CCSM_FEAT_NPASS)) { shp = {31, 4}; return true; }
    if (key == "pos_encoder.pos_embed.weight") { shp = {L, d}; return true; }
    if (key == "classifier.0.weight") { shp = {2 * d, 2 * d}; return true; }
    if (key == "classifier.0.bias") { shp = {2 * d}; return true; }
    if (key == "classifier.3.weight") { shp = {C, 2 * d}; return true; }
    if (key == "classifier.3.bias") { shp = {C}; return true; }
    static const char* convs[3] = {"trans_input.conv_embed.0", "trans_input.conv_embed.4", "trans_input.conv_embed_plus.0.conv_embed.0"};
    static const char* bns[3] = {"trans_input.conv_embed.1", "trans_input.conv_embed.5", "trans_input.conv_embed_plus.0.conv_embed.1"};
    const int64_t cin[3] = {m->in_feat, d / 2, d}, cout[3] = {d / 2, d, d};
    for (int i = 0; i < 3; ++i) {
      if (key == std::string(convs[i]) + ".weight") { shp = {cout[i], cin[i], 3}; return true; }
      for (const char* t : {".weight", ".bias", ".running_mean", ".running_var"})
        if (key == std::string(bns[i]) + t) { shp = {cout[i]}; return true; }
    }
    for (int l = 0; l < m->cfg.num_layers; ++l) {
      const std::string pre = "transformer_encoder.layers." + std::to_string(l) + ".";
      if (key == pre + "self_attn.in_proj_weight") { shp = {3 * d, d}; return true; }
      if (key == pre + "self_attn.in_proj_bias") { shp = {3 * d}; return true; }
      if (key == pre + "self_attn.out_proj.weight") { shp = {d, d}; return true; }
      if (key == pre + "self_attn.out_proj.bias") { shp = {d}; return true; }
      if (key == pre + "linear1.weight") { shp = {ff, d}; return true; }   // ff = 0 until known: see ccsm_set_weight
      if (key == pre + "linear1.bias") { shp = {ff}; return true; }
      if (key == pre + "linear2.weight") { shp = {d, ff}; return true; }
      if (key == pre + "linear2.bias") { shp = {d}; return true; }
      for (const char* t : {"norm1.weight", "norm1.bias", "norm2.weight", "norm2.bias"})
        if (key == pre + t) { shp = {d}; return true; }
    }
    return false;
  }
  if (m->is_2s2) {
    if (key == "seq_embed.weight") { shp = {m->cfg.n_vocab, m->cfg.n_embed}; return true; }
    if (key == "ipd_embed.weight" || key == "pw_embed.weight") { shp = {953, 8}; return true; }  // MAX_KINETICS + 1
    if (key == "npass_embed.weight" && (m->cfg.feat_flags & CCSM_FEAT_NPASS)) { shp = {31, 4}; return true; }
    if (key == "classifier.0.weight") { shp = {4 * H, 4 * H}; return true; }
    if (key == "classifier.0.bias") { shp = {4 * H}; return true; }
    if (key == "classifier.3.weight") { shp = {C, 4 * H}; return true; }
    if (key == "classifier.3.bias") { shp = {C}; return true; }
  } else {
    if (key == "embed.weight" && m->cfg.kind == CCSM_KIND_ATT2S) { shp = {m->cfg.n_vocab, m->cfg.n_embed}; return true; }
    if (key == "fc1.weight") { shp = {C, 2 * H * m->strands}; return true; }
    if (key == "fc1.bias") { shp = {C}; return true; }
  }
  if (key == "_att3.Wa.weight" || key == "_att3.Ua.weight") { shp = {H, 2 * H}; return true; }
  if (key == "_att3.va.weight") { shp = {1, H}; return true; }
  for (int l = 0; l < m->cfg.num_layers; ++l)
    for (int d = 0; d < 2; ++d) {
      std::string sfx = "_l" + std::to_string(l) + (d ? "_reverse" : "");
      int64_t K = l == 0 ? m->in_feat : 2 * H;
      const int64_t G = m->gates;
      if (key == "rnn.weight_ih" + sfx) { shp = {G * H, K}; return true; }
      if (key == "rnn.weight_hh" + sfx) { shp = {G * H, H}; return true; }
      if (key == "rnn.bias_ih" + sfx || key == "rnn.bias_hh" + sfx) { shp = {G * H}; return true; }
    }
  return false;
}

int ccsm_set_weight(ccsm_model* m, const char* key_c, const float* host, const int64_t* shape, int32_t ndim) {
  if (!m || !key_c || !host || !shape || ndim < 1 || ndim > 4) {
    set_error("ccsm_set_weight: bad argument");
    return CCSM_EINVAL;
  }
  std::string key(key_c);
  if (key.rfind("module.", 0) == 0) key = key.substr(7);  // DDP prefix (reference call_modifications.py:350-358)
  if (m->is_trans && m->dim_ff == 0) {
    // dim_feedforward is not part of ccsm_config: take it from the first feed-forward tensor that arrives
    const size_t n = key.size();
    if (n > 14 && key.compare(n - 14, 14, "linear1.weight") == 0 && ndim == 2) m->dim_ff = (int)shape[0];
    else if (n > 12 && key.compare(n - 12, 12, "linear1.bias") == 0 && ndim == 1) m->dim_ff = (int)shape[0];
    else if (n > 14 && key.compare(n - 14, 14, "linear2.weight") == 0 && ndim == 2) m->dim_ff = (int)shape[1];
    if (m->dim_ff % 16 != 0) {
      set_error("ccsm_set_weight: dim_feedforward %d must be a multiple of 16", m->dim_ff);
      m->dim_ff = 0;
      return CCSM_EKEY;
    }
  }
  std::vector<int64_t> want;
  if (!expected_shape(m, key, want)) {
    set_error("ccsm_set_weight: unexpected key '%s' for this model", key.c_str());
    return CCSM_EKEY;
  }
  std::vector<int64_t> got(shape, shape + ndim);
  if (got != want) {
    std::string a, b;
    for (auto v : got) a += std::to_string(v) + ",";
    for (auto v : want) b += std::to_string(v) + ",";
    set_error("ccsm_set_weight: size mismatch for %s: got (%s) expected (%s)", key.c_str(), a.c_str(), b.c_str());
    return CCSM_EKEY;
  }
  int64_t numel = 1;
  for (auto v : got) numel *= v;
  HostTensor t;
  t.shape = got;
  t.data.assign(host, host + numel);
  m->w[key] = std::move(t);
  m->finalized = false;
  return CCSM_OK;
}

static int check_complete(ccsm_model* m) {
  if (m->is_trans) {
    std::vector<std::string> tk = {"seq_embed.weight", "ipd_embed.weight", "pw_embed.weight", "pos_encoder.pos_embed.weight",
                                   "classifier.0.weight", "classifier.0.bias", "classifier.3.weight", "classifier.3.bias",
                                   "trans_input.conv_embed.0.weight", "trans_input.conv_embed.4.weight",
                                   "trans_input.conv_embed_plus.0.conv_embed.0.weight"};
    if (m->cfg.feat_flags & CCSM_FEAT_NPASS) tk.push_back("npass_embed.weight");
    for (const char* bn : {"trans_input.conv_embed.1", "trans_input.conv_embed.5", "trans_input.conv_embed_plus.0.conv_embed.1"})
      for (const char* t : {".weight", ".bias", ".running_mean", ".running_var"}) tk.push_back(std::string(bn) + t);
    for (int l = 0; l < m->cfg.num_layers; ++l)
      for (const char* t : {"self_attn.in_proj_weight", "self_attn.in_proj_bias", "self_attn.out_proj.weight",
                            "self_attn.out_proj.bias", "linear1.weight", "linear1.bias", "linear2.weight", "linear2.bias",
                            "norm1.weight", "norm1.bias", "norm2.weight", "norm2.bias"})
        tk.push_back("transformer_encoder.layers." + std::to_string(l) + "." + t);
    for (auto& k : tk)
      if (!m->w.count(k)) {
        set_error("ccsm_finalize: missing key '%s'", k.c_str());
        return CCSM_EKEY;
      }
    return CCSM_OK;
  }
  std::vector<std::string> keys = {"_att3.Wa.weight", "_att3.Ua.weight", "_att3.va.weight"};
  if (m->is_2s2) {
    for (const char* k : {"seq_embed.weight", "ipd_embed.weight", "pw_embed.weight", "classifier.0.weight", "classifier.0.bias",
                          "classifier.3.weight", "classifier.3.bias"})
      keys.push_back(k);
    if (m->cfg.feat_flags & CCSM_FEAT_NPASS) keys.push_back("npass_embed.weight");
  } else {
    keys.push_back("fc1.weight");
    keys.push_back("fc1.bias");
    if (m->cfg.kind == CCSM_KIND_ATT2S) keys.push_back("embed.weight");
  }
  for (int l = 0; l < m->cfg.num_layers; ++l)
    for (int d = 0; d < 2; ++d) {
      std::string sfx = "_l" + std::to_string(l) + (d ? "_reverse" : "");
      for (const char* p : {"rnn.weight_ih", "rnn.weight_hh", "rnn.bias_ih", "rnn.bias_hh"}) keys.push_back(p + sfx);
    }
  for (auto& k : keys)
    if (!m->w.count(k)) {
      set_error("ccsm_finalize: missing key '%s'", k.c_str());
      return CCSM_EKEY;
    }
  return CCSM_OK;
}

int ccsm_finalize(ccsm_model* m) {
  if (!m) {
    set_error("ccsm_finalize: null model");
    return CCSM_EINVAL;
  }
  CCSM_CUDA(cudaSetDevice(m->cfg.device));
  CCSM_TRY(check_complete(m));
  if (m->is_trans) {
    CCSM_TRY(trans_upload_weights(m));
    m->finalized = true;
    return CCSM_OK;
  }
  CCSM_TRY(fp32_upload_weights(m));  // always kept: cross-check path + attention/head fallback
  if (is_tc(m->cfg.precision)) CCSM_TRY(tc_upload_weights(m));
  if (m->cfg.kind == CCSM_KIND_AGGR) CCSM_TRY(aggr_fused_upload(m));
  m->finalized = true;
  return CCSM_OK;
}

int ccsm_set_h0_mode(ccsm_model* m, int32_t mode, uint64_t seed) {
  if (!m || (mode != CCSM_H0_ZEROS && mode != CCSM_H0_DEVICE_RANDOM && mode != CCSM_H0_TORCH_STREAM)) {
    set_error("ccsm_set_h0_mode: bad argument");
    return CCSM_EINVAL;
  }
  if (mode == CCSM_H0_TORCH_STREAM) {
    if (m->cfg.kind != CCSM_KIND_ATT2S || m->gates != 3 || m->is_trans) {
      set_error("ccsm_set_h0_mode: the torch.randn stream is implemented for the GRU att2s models");
      return CCSM_EUNSUPPORTED;
    }
    CCSM_CUDA(cudaSetDevice(m->cfg.device));
    CCSM_TRY(mt_seed(m, seed));
  }
  m->h0_mode = mode;
  m->h0_seed = seed;
  m->h0_calls = 0;
  return CCSM_OK;
}

int ccsm_set_h0_batching(ccsm_model* m, const int64_t* holebatch_sites, int64_t n_holebatches, int32_t batch_size) {
  if (!m || (n_holebatches > 0 && !holebatch_sites) || n_holebatches < 0) {
    set_error("ccsm_set_h0_batching: bad argument");
    return CCSM_EINVAL;
  }
  CCSM_CUDA(cudaSetDevice(m->cfg.device));
  return mt_set_batching(m, holebatch_sites, n_holebatches, batch_size);
}

int ccsm_h0_stream_set_state(ccsm_model* m, const uint32_t* words624, int32_t pos) {
  if (!m || !words624) {
    set_error("ccsm_h0_stream_set_state: bad argument");
    return CCSM_EINVAL;
  }
  CCSM_CUDA(cudaSetDevice(m->cfg.device));
  return mt_set_state(m, words624, pos);
}

int ccsm_h0_stream_get_state(ccsm_model* m, uint32_t* words624, int32_t* pos) {
  if (!m || !words624 || !pos) {
    set_error("ccsm_h0_stream_get_state: bad argument");
    return CCSM_EINVAL;
  }
  CCSM_CUDA(cudaSetDevice(m->cfg.device));
  return mt_get_state(m, words624, pos);
}

int ccsm_set_precision(ccsm_model* m, int32_t precision) {
  if (!m || precision < CCSM_PREC_FP32 || precision > CCSM_PREC_FP16C8) {
    set_error("ccsm_set_precision: bad argument");
    return CCSM_EINVAL;
  }
  if (m->cfg.kind == CCSM_KIND_AGGR || m->gates == 4 || m->is_2s2 || m->is_trans) precision = CCSM_PREC_FP32;
  if (precision == m->cfg.precision) return CCSM_OK;
  m->cfg.precision = precision;
  if (m->finalized && is_tc(precision)) {
    CCSM_CUDA(cudaSetDevice(m->cfg.device));
    CCSM_TRY(tc_upload_weights(m));
  }
  return CCSM_OK;
}

}  // extern "C"

namespace ccsm {

void seg_chunks(const std::vector<int64_t>& segs, int64_t max_sites, std::vector<SegChunk>& out) {
  out.clear();
  int64_t site = 0;
  for (int k = 0; k < (int)segs.size();) {
    SegChunk c{k, 0, site, 0};
    while (k < (int)segs.size() && (c.nseg == 0 || c.sites + segs[k] <= max_sites)) {
      c.sites += segs[k];
      ++c.nseg;
      ++k;
    }
    site += c.sites;
    out.push_back(c);
  }
}

static ccsm_strand strand_at(const ccsm_strand* s, int64_t off, int L) {
  ccsm_strand r = *s;
  auto adv = [&](const float* p, int64_t per) { return p ? p + off * per : nullptr; };
  r.kmer = adv(s->kmer, L);
  r.kpass = adv(s->kpass, L);
  r.ipd_means = adv(s->ipd_means, L);
  r.ipd_stds = adv(s->ipd_stds, L);
  r.pw_means = adv(s->pw_means, L);
  r.pw_stds = adv(s->pw_stds, L);
  r.sns = adv(s->sns, 4);
  r.maps = adv(s->maps, L);
  return r;
}

int forward_att2s_dev(ccsm_model* m, int64_t n, const ccsm_strand* fwd, const ccsm_strand* rev, const float* h0_fwd,
                      const float* h0_rev, float* logits, float* probs, cudaStream_t st, const int64_t* segs, int nseg) {
  if (m->is_trans) return trans_forward(m, n, fwd, rev, logits, probs, st);  // no recurrent state: h0 is ignored
  auto inner = [&](int64_t cn, const ccsm_strand* f, const ccsm_strand* r, const float* ha, const float* hb, float* lg,
                   float* pr) {
    const int rc = is_tc(m->cfg.precision) ? tc_forward_att2s(m, cn, f, r, ha, hb, lg, pr, st)
                                           : fp32_forward_att2s(m, cn, f, r, ha, hb, lg, pr, st);
    m->h0_calls += 1;
    return rc;
  };
  if (m->h0_mode != CCSM_H0_TORCH_STREAM || h0_fwd || h0_rev || m->gates != 3)
    return inner(n, fwd, rev, h0_fwd, h0_rev, logits, probs);
  // the reference's torch.randn stream, drawn on the device: windows of whole model calls
  std::vector<int64_t> own;
  if (!segs) {
    CCSM_TRY(mt_take_segments(m, n, own));
    segs = own.data();
    nseg = (int)own.size();
  }
  std::vector<int64_t> sv(segs, segs + nseg);
  std::vector<SegChunk> chunks;
  seg_chunks(sv, 75776, chunks);
  const int L = m->cfg.seq_len, C = m->cfg.num_classes;
  for (const SegChunk& c : chunks) {
    const float *ha = nullptr, *hb = nullptr;
    int buf = 0;
    CCSM_TRY(mt_fill(m, segs + c.seg0, c.nseg, st, &ha, &hb, &buf));
    const ccsm_strand f = strand_at(fwd, c.site0, L), r = strand_at(rev, c.site0, L);
    const int rc = inner(c.sites, &f, &r, ha, hb, logits ? logits + c.site0 * C : nullptr,
                         probs ? probs + c.site0 * C : nullptr);
    CCSM_TRY(mt_release(m, buf, st));
    CCSM_TRY(rc);
  }
  return CCSM_OK;
}

}  // namespace ccsm

extern "C" {

static int check_strand(const ccsm_model* m, const ccsm_strand* s, const char* which) {
  const int f = m->cfg.feat_flags;
  if (!s || !s->kmer || !s->ipd_means || !s->pw_means || ((f & CCSM_FEAT_NPASS) && !s->kpass) ||
      ((f & CCSM_FEAT_STDS) && (!s->ipd_stds || !s->pw_stds)) || ((f & CCSM_FEAT_SN) && !s->sns) ||
      ((f & CCSM_FEAT_MAP) && !s->maps)) {
    set_error("forward: %s strand is missing a tensor required by feat_flags=%d", which, f);
    return CCSM_EINVAL;
  }
  return CCSM_OK;
}

int ccsm_forward_att2s(ccsm_model* m, int64_t n, const ccsm_strand* fwd, const ccsm_strand* rev, const float* h0_fwd,
                       const float* h0_rev, float* logits, float* probs, void* stream) {
  if (!m || m->cfg.kind != CCSM_KIND_ATT2S) {
    set_error("ccsm_forward_att2s: not an att2s model");
    return CCSM_EINVAL;
  }
  if (!m->finalized) {
    set_error("ccsm_forward_att2s: model not finalized");
    return CCSM_ESTATE;
  }
  if (n < 0) {
    set_error("ccsm_forward_att2s: n < 0");
    return CCSM_EINVAL;
  }
  if (n == 0) return CCSM_OK;  // empty batch: nothing to do (the reference skips it, call_modifications.py:200)
  CCSM_TRY(check_strand(m, fwd, "forward"));
  CCSM_TRY(check_strand(m, rev, "reverse"));
  CCSM_CUDA(cudaSetDevice(m->cfg.device));
  return forward_att2s_dev(m, n, fwd, rev, h0_fwd, h0_rev, logits, probs, reinterpret_cast<cudaStream_t>(stream), nullptr, 0);
}

int ccsm_forward_att2s_lstm(ccsm_model* m, int64_t n, const ccsm_strand* fwd, const ccsm_strand* rev, const float* h0_fwd,
                            const float* c0_fwd, const float* h0_rev, const float* c0_rev, float* logits, float* probs,
                            void* stream) {
  if (!m || m->cfg.kind != CCSM_KIND_ATT2S || m->gates != 4) {
    set_error("ccsm_forward_att2s_lstm: not an LSTM (CCSM_CELL_LSTM) att2s model");
    return CCSM_EINVAL;
  }
  if (!m->finalized) {
    set_error("ccsm_forward_att2s_lstm: model not finalized");
    return CCSM_ESTATE;
  }
  if (n < 0) {
    set_error("ccsm_forward_att2s_lstm: n < 0");
    return CCSM_EINVAL;
  }
  if (n == 0) return CCSM_OK;
  CCSM_TRY(check_strand(m, fwd, "forward"));
  CCSM_TRY(check_strand(m, rev, "reverse"));
  CCSM_CUDA(cudaSetDevice(m->cfg.device));
  const int rc = fp32_forward_att2s(m, n, fwd, rev, h0_fwd, h0_rev, logits, probs, reinterpret_cast<cudaStream_t>(stream),
                                    c0_fwd, c0_rev);
  m->h0_calls += 1;
  return rc;
}

static int forward_aggr_impl(ccsm_model* m, int64_t n, const float* offsets, const float* histos, const float* h0,
                             const float* c0, float* out, void* stream, const char* who) {
  if (!m || m->cfg.kind != CCSM_KIND_AGGR) {
    set_error("%s: not an aggregate model", who);
    return CCSM_EINVAL;
  }
  if (!m->finalized) {
    set_error("%s: model not finalized", who);
    return CCSM_ESTATE;
  }
  if (n < 0 || (n > 0 && (!offsets || !histos || !out))) {
    set_error("%s: bad argument", who);
    return CCSM_EINVAL;
  }
  if (n == 0) return CCSM_OK;
  CCSM_CUDA(cudaSetDevice(m->cfg.device));
  // one fused kernel for the shipped configuration (GRU, H = 32, 21 inputs, one layer); CCSM_AGGR_UNFUSED=1 keeps the
  // layer-by-layer fp32 kernels (the path other shapes and the LSTM cell take) for cross-checks
  const char* env = getenv("CCSM_AGGR_UNFUSED");
  const bool unfused = env && atoi(env) != 0;
  if (!unfused && aggr_fused_supported(m))
    return aggr_fused_forward(m, n, offsets, histos, h0, out, reinterpret_cast<cudaStream_t>(stream));
  return fp32_forward_aggr(m, n, offsets, histos, h0, out, reinterpret_cast<cudaStream_t>(stream), c0);
}

int ccsm_forward_aggr(ccsm_model* m, int64_t n, const float* offsets, const float* histos, const float* h0, float* out,
                      void* stream) {
  return forward_aggr_impl(m, n, offsets, histos, h0, nullptr, out, stream, "ccsm_forward_aggr");
}

int ccsm_forward_aggr_sites(ccsm_model* m, int64_t n, const int64_t* site_pos, const float* site_histo,
                            int32_t only_close, const float* h0, float* out, void* stream) {
  if (!m || m->cfg.kind != CCSM_KIND_AGGR) {
    set_error("ccsm_forward_aggr_sites: not an aggregate model");
    return CCSM_EINVAL;
  }
  if (!m->finalized) {
    set_error("ccsm_forward_aggr_sites: model not finalized");
    return CCSM_ESTATE;
  }
  if (n < 0 || (n > 0 && (!site_pos || !site_histo || !out))) {
    set_error("ccsm_forward_aggr_sites: bad argument");
    return CCSM_EINVAL;
  }
  if (n == 0) return CCSM_OK;
  if (!aggr_fused_supported(m)) {
    set_error("ccsm_forward_aggr_sites: in-kernel windows need the fused configuration (GRU, hidden 32, one layer)");
    return CCSM_EUNSUPPORTED;
  }
  CCSM_CUDA(cudaSetDevice(m->cfg.device));
  return aggr_fused_forward_sites(m, n, reinterpret_cast<const long long*>(site_pos), site_histo, only_close ? 1 : 0, h0,
                                  out, reinterpret_cast<cudaStream_t>(stream));
}

int ccsm_forward_aggr_lstm(ccsm_model* m, int64_t n, const float* offsets, const float* histos, const float* h0,
                           const float* c0, float* out, void* stream) {
  if (m && m->gates != 4) {
    set_error("ccsm_forward_aggr_lstm: not an LSTM aggregate model (CCSM_AGGR_LSTM)");
    return CCSM_EINVAL;
  }
  return forward_aggr_impl(m, n, offsets, histos, h0, c0, out, stream, "ccsm_forward_aggr_lstm");
}

// Host-buffer entry: H2D staging -> forward -> D2H, pipelined over chunks with two staging buffers.
// streams[0] copies in, streams[1] computes and copies the (tiny) results out.
int ccsm_forward_att2s_host(ccsm_model* m, int64_t n, const ccsm_strand* fwd, const ccsm_strand* rev,
                            const float* h0_fwd, const float* h0_rev, float* logits, float* probs) {
  if (!m || m->cfg.kind != CCSM_KIND_ATT2S) {
    set_error("ccsm_forward_att2s_host: not an att2s model");
    return CCSM_EINVAL;
  }
  if (!m->finalized) {
    set_error("ccsm_forward_att2s_host: model not finalized");
    return CCSM_ESTATE;
  }
  if (n < 0) {
    set_error("ccsm_forward_att2s_host: n < 0");
    return CCSM_EINVAL;
  }
  if (n == 0) return CCSM_OK;
  CCSM_TRY(check_strand(m, fwd, "forward"));
  CCSM_TRY(check_strand(m, rev, "reverse"));
  CCSM_CUDA(cudaSetDevice(m->cfg.device));
  for (int i = 0; i < 2; ++i) {
    if (!m->streams[i]) CCSM_CUDA(cudaStreamCreateWithFlags(&m->streams[i], cudaStreamNonBlocking));
    if (!m->events[i]) CCSM_CUDA(cudaEventCreateWithFlags(&m->events[i], cudaEventDisableTiming));
  }
  const int L = m->cfg.seq_len, H = m->cfg.hidden, NL = m->cfg.num_layers, C = m->cfg.num_classes;
  // one tensor-core library chunk (8 row-tile items per SM) per staging buffer; in the torch-stream h0 mode the chunks
  // are cut at the reference's model-call boundaries so that every chunk draws whole randn calls
  const int64_t kHostChunk = 75776;
  const bool torch_h0 = m->h0_mode == CCSM_H0_TORCH_STREAM && !h0_fwd && !h0_rev && m->gates == 3 && !m->is_trans;
  std::vector<int64_t> segs;
  std::vector<SegChunk> chunks;
  if (torch_h0) {
    CCSM_TRY(mt_take_segments(m, n, segs));
    seg_chunks(segs, kHostChunk, chunks);
  } else {
    for (int64_t s0 = 0; s0 < n; s0 += kHostChunk) chunks.push_back(SegChunk{0, 0, s0, (n - s0) < kHostChunk ? (n - s0) : kHostChunk});
  }
  int64_t chunk = 0;
  for (const SegChunk& c : chunks) chunk = c.sites > chunk ? c.sites : chunk;
  // staging layout per buffer (floats): 2 strands x [kmer,kpass,ipd,ipd_sd,pw,pw_sd,maps](L each) + sns(4) + h0 x2
  const int64_t per_strand = (int64_t)7 * L + 4;
  const int64_t h0_floats = (int64_t)2 * NL * H;
  const size_t in_bytes = (size_t)chunk * (2 * per_strand + ((h0_fwd || h0_rev) ? 2 * h0_floats : 0)) * sizeof(float);
  const size_t out_bytes = (size_t)chunk * 2 * C * sizeof(float);
  for (int i = 0; i < 2; ++i) {
    CCSM_TRY(m->stage_in[i].reserve(in_bytes));
    CCSM_TRY(m->stage_out[i].reserve(out_bytes));
    if (!m->done[i]) CCSM_CUDA(cudaEventCreateWithFlags(&m->done[i], cudaEventDisableTiming));
  }
  cudaStream_t s_in = m->streams[0], s_cmp = m->streams[1];
  int rc = CCSM_OK;
  cudaError_t ce = cudaSuccess;  // first failing runtime call of the pipelined loop
  auto ok = [&](cudaError_t e) {
    if (e != cudaSuccess && ce == cudaSuccess) ce = e;
    return e == cudaSuccess;
  };
  int64_t ci = 0;
  for (const SegChunk& ck : chunks) {
    if (rc != CCSM_OK || ce != cudaSuccess) break;
    const int b = (int)(ci & 1);
    const int64_t s0 = ck.site0, cn = ck.sites;
    float* base = m->stage_in[b].as<float>();
    if (ci >= 2) ok(cudaStreamWaitEvent(s_in, m->done[b], 0));  // buffer b is free once chunk ci-2 finished
    ++ci;
    ccsm_strand dev[2];
    const ccsm_strand* src[2] = {fwd, rev};
    float* cur = base;
    auto stage = [&](const float* hp, int64_t per) -> const float* {
      if (!hp) return nullptr;
      float* d = cur;
      cur += cn * per;
      ok(cudaMemcpyAsync(d, hp + s0 * per, (size_t)cn * per * sizeof(float), cudaMemcpyHostToDevice, s_in));
      return d;
    };
    for (int s = 0; s < 2; ++s) {
      dev[s].kmer = stage(src[s]->kmer, L);
      dev[s].kpass = stage(src[s]->kpass, L);
      dev[s].ipd_means = stage(src[s]->ipd_means, L);
      dev[s].ipd_stds = (m->cfg.feat_flags & CCSM_FEAT_STDS) ? stage(src[s]->ipd_stds, L) : nullptr;
      dev[s].pw_means = stage(src[s]->pw_means, L);
      dev[s].pw_stds = (m->cfg.feat_flags & CCSM_FEAT_STDS) ? stage(src[s]->pw_stds, L) : nullptr;
      dev[s].sns = (m->cfg.feat_flags & CCSM_FEAT_SN) ? stage(src[s]->sns, 4) : nullptr;
      dev[s].maps = (m->cfg.feat_flags & CCSM_FEAT_MAP) ? stage(src[s]->maps, L) : nullptr;
    }
    const float* h0s[2] = {h0_fwd, h0_rev};
    const float* dh0[2] = {nullptr, nullptr};
    for (int s = 0; s < 2; ++s) {
      if (!h0s[s]) continue;
      float* d = cur;
      cur += cn * h0_floats;
      // (2*layers, n, H) -> (2*layers, cn, H): one strided copy
      ok(cudaMemcpy2DAsync(d, (size_t)cn * H * sizeof(float), h0s[s] + s0 * H, (size_t)n * H * sizeof(float),
                           (size_t)cn * H * sizeof(float), 2 * NL, cudaMemcpyHostToDevice, s_in));
      dh0[s] = d;
    }
    ok(cudaEventRecord(m->events[b], s_in));
    ok(cudaStreamWaitEvent(s_cmp, m->events[b], 0));
    if (ce != cudaSuccess) break;
    float* dl = m->stage_out[b].as<float>();
    float* dp = dl + cn * C;
    rc = forward_att2s_dev(m, cn, &dev[0], &dev[1], dh0[0], dh0[1], dl, dp, s_cmp, torch_h0 ? segs.data() + ck.seg0 : nullptr,
                           torch_h0 ? ck.nseg : 0);
    if (rc != CCSM_OK) break;
    if (logits) ok(cudaMemcpyAsync(logits + s0 * C, dl, (size_t)cn * C * sizeof(float), cudaMemcpyDeviceToHost, s_cmp));
    if (probs) ok(cudaMemcpyAsync(probs + s0 * C, dp, (size_t)cn * C * sizeof(float), cudaMemcpyDeviceToHost, s_cmp));
    ok(cudaEventRecord(m->done[b], s_cmp));
  }
  cudaError_t e1 = cudaStreamSynchronize(s_in);
  cudaError_t e2 = cudaStreamSynchronize(s_cmp);
  if (rc != CCSM_OK) return rc;
  if (ce != cudaSuccess || e1 != cudaSuccess || e2 != cudaSuccess) {
    set_error("ccsm_forward_att2s_host: %s", cudaGetErrorString(ce != cudaSuccess ? ce : (e1 != cudaSuccess ? e1 : e2)));
    return CCSM_ECUDA;
  }
  return CCSM_OK;
}

int64_t ccsm_debug_last_rnn_out(ccsm_model* m, float* host, int64_t cap) {
  if (!m || !host || !m->dbg_rnn_out) {
    set_error("ccsm_debug_last_rnn_out: nothing recorded");
    return CCSM_ESTATE;
  }
  int64_t nfl = m->dbg_rnn_out_floats < cap ? m->dbg_rnn_out_floats : cap;
  CCSM_CUDA(cudaSetDevice(m->cfg.device));
  CCSM_CUDA(cudaDeviceSynchronize());
  CCSM_CUDA(cudaMemcpy(host, m->dbg_rnn_out, (size_t)nfl * sizeof(float), cudaMemcpyDeviceToHost));
  return nfl;
}

int ccsm_profile_enable(ccsm_model* m, int32_t on) {
  if (!m) {
    set_error("ccsm_profile_enable: null model");
    return CCSM_EINVAL;
  }
  m->prof.on = on != 0;
  return CCSM_OK;
}

int ccsm_profile_read(ccsm_model* m, double* ms, double* units, int64_t* launches, int32_t nclass) {
  if (!m || !ms || !units || !launches || nclass < 4) {
    set_error("ccsm_profile_read: bad argument");
    return CCSM_EINVAL;
  }
  CCSM_CUDA(cudaSetDevice(m->cfg.device));
  for (int i = 0; i < nclass; ++i) {
    ms[i] = 0;
    units[i] = 0;
    launches[i] = 0;
  }
  for (auto& r : m->prof.recs) {
    CCSM_CUDA(cudaEventSynchronize(r.b));
    float t = 0.f;
    CCSM_CUDA(cudaEventElapsedTime(&t, r.a, r.b));
    if (r.cls < nclass) {
      ms[r.cls] += t;
      units[r.cls] += r.units;
      launches[r.cls] += 1;
    }
    cudaEventDestroy(r.a);
    cudaEventDestroy(r.b);
  }
  m->prof.recs.clear();
  return CCSM_OK;
}

int64_t ccsm_debug_tc_layer_out(ccsm_model* m, int32_t layer, float* host, int64_t cap) {
  if (!m || !host || layer < 0 || layer >= m->cfg.num_layers) {
    set_error("ccsm_debug_tc_layer_out: bad argument");
    return CCSM_EINVAL;
  }
  CCSM_CUDA(cudaSetDevice(m->cfg.device));
  int64_t written = 0;
  int rc = tc_debug_layer_out(m, layer, host, cap, &written);
  return rc != CCSM_OK ? rc : written;
}

}  // extern "C"
