// Internal declarations shared by the C-ABI layer and the kernel files of libccsm.so.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <map>
#include <string>
#include <vector>

#include "../../include/ccsm.h"

namespace ccsm {

void set_error(const char* fmt, ...);
extern std::atomic<int64_t> g_launches;
inline void count_launch(int n = 1) { g_launches.fetch_add(n, std::memory_order_relaxed); }

#define CCSM_CUDA(expr)                                                                        \
  do {                                                                                         \
    cudaError_t _e = (expr);                                                                   \
    if (_e != cudaSuccess) {                                                                   \
      ccsm::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return CCSM_ECUDA;                                                                       \
    }                                                                                          \
  } while (0)

#define CCSM_TRY(expr)          \
  do {                          \
    int _r = (expr);            \
    if (_r != CCSM_OK) return _r; \
  } while (0)

struct HostTensor {
  std::vector<float> data;
  std::vector<int64_t> shape;
};

// A device allocation that grows on demand (workspace sized lazily on larger n).
struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  int reserve(size_t bytes);
  void release();
  template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

inline int round_up(int x, int m) { return (x + m - 1) / m * m; }

// fp32 (FFMA) path: weights in PyTorch (N, K) row-major layout, K zero-padded to a multiple of 16,
// both directions of a layer stacked along N for the input projection.
struct Fp32Layer {
  int K = 0, Kpad = 0;
  DevBuf w_ih;  // (2*3H, Kpad)
  DevBuf b_ih;  // (2*3H)
  DevBuf w_hh;  // (2, 3H, H)
  DevBuf b_hh;  // (2, 3H)
};

struct Fp32Weights {
  std::vector<Fp32Layer> layers;
  DevBuf embed;        // (n_vocab, n_embed)
  DevBuf ipd_embed, pw_embed, npass_embed;  // ModelAttRNN2: (953, 8), (953, 8), (31, 4)
  DevBuf cls0_w, cls0_b;                    // ModelAttRNN2: classifier.0 (4H, 4H), (4H); classifier.3 lives in fc_w / fc_b
  DevBuf Wa, Ua, va;   // (H, 2H), (H, 2H), (H)
  DevBuf fc_w, fc_b;   // (classes, strands*2H), (classes)
  bool ready = false;
};

// ModelTransEnc weights (fp32_path.cu, transformer section)
struct TrConv { DevBuf w, scale, shift; int cin = 0, cout = 0, kpad = 0; };
struct TrLayer { DevBuf in_w, in_b, out_w, out_b, l1_w, l1_b, l2_w, l2_b, n1_w, n1_b, n2_w, n2_b; };
struct TrWeights {
  TrConv conv[3];
  DevBuf pos;
  std::vector<TrLayer> layers;
  DevBuf x0, col, a, b, qkv, ffh, ctx, hid;  // workspace
  int64_t tokens_cap = 0;
};

struct Fp32Workspace {
  int64_t rows_cap = 0;
  DevBuf x0, gi, gh, h, c, outA, outB, qa;
};

struct TcState;  // tensor-core (tcgen05) path state, defined in tc_path.cu
struct PuState;  // resident region pileup (pileup.cu)
struct ExState;  // device feature extraction state (resident read batch), defined in extract.cu
struct MtStream; // device reproduction of the reference's torch.randn h0 stream (mtstream.cu)

// Per-kernel-class device timing (CUDA events recorded on the launching stream around each launch).
enum ProfClass { PROF_PREP = 0, PROF_GRU_L0 = 1, PROF_GRU_LN = 2, PROF_ATT = 3, PROF_EX_SCAN = 4, PROF_EX_GATHER = 5,
                 PROF_NCLASS = 6 };
struct ProfRec {
  cudaEvent_t a, b;
  int cls;
  double units;  // sites processed by this launch
};
struct Profiler {
  bool on = false;
  std::vector<ProfRec> recs;
  int begin(int cls, double units, cudaStream_t st) {
    if (!on) return -1;
    ProfRec r;
    r.cls = cls;
    r.units = units;
    if (cudaEventCreate(&r.a) != cudaSuccess || cudaEventCreate(&r.b) != cudaSuccess) return -1;
    cudaEventRecord(r.a, st);
    recs.push_back(r);
    return (int)recs.size() - 1;
  }
  void end(int id, cudaStream_t st) {
    if (id >= 0) cudaEventRecord(recs[id].b, st);
  }
};

}  // namespace ccsm

struct ccsm_model {
  ccsm_config cfg{};
  int strands = 2;    // 2 for att2s, 1 for aggr
  int in_feat = 0;    // GRU layer-0 input width (att2s: n_embed + feas_ccs; aggr: bins + 1)
  int gates = 3;      // 3 GRU (r, z, n) / 4 LSTM (i, f, g, o): gate row blocks of the rnn weights
  bool is_2s2 = false;  // ModelAttRNN2: integer kinetics embeddings + two-layer classifier (CCSM_MODEL_2S2)
  bool is_trans = false;  // ModelTransEnc (CCSM_MODEL_TRANSENC)
  int nhead = 0, dim_ff = 0;
  ccsm::TrWeights tr;
  std::map<std::string, ccsm::HostTensor> w;
  bool finalized = false;
  ccsm::Fp32Weights fp32;
  ccsm::Fp32Workspace ws32;
  ccsm::TcState* tc = nullptr;
  ccsm::ExState* ex = nullptr;
  ccsm::PuState* pu = nullptr;
  ccsm::MtStream* mts = nullptr;
  ccsm::DevBuf aggr_packed, aggr_packed_tiled, aggr_scratch;  // fused aggregate kernels (aggr_fused.cu)
  ccsm::Profiler prof;
  int h0_mode = 0;            // CCSM_H0_*
  uint64_t h0_seed = 0;
  uint64_t h0_calls = 0;      // forward calls so far (Philox offset = 256 * call)
  // host-entry staging
  cudaStream_t streams[2] = {nullptr, nullptr};
  cudaEvent_t events[2] = {nullptr, nullptr};
  cudaEvent_t done[2] = {nullptr, nullptr};
  ccsm::DevBuf stage_in[2], stage_out[2];
  // debug: where the last layer-stack output of the most recent fp32 chunk lives
  const float* dbg_rnn_out = nullptr;
  int64_t dbg_rnn_out_floats = 0;
};

namespace ccsm {

// ---- fp32 path (fp32_path.cu)
int fp32_upload_weights(ccsm_model* m);
int fp32_forward_att2s(ccsm_model* m, int64_t n, const ccsm_strand* fwd, const ccsm_strand* rev,
                       const float* h0_f, const float* h0_r, float* logits, float* probs, cudaStream_t st,
                       const float* c0_f = nullptr, const float* c0_r = nullptr);
int fp32_forward_aggr(ccsm_model* m, int64_t n, const float* offsets, const float* histos, const float* h0,
                      float* out, cudaStream_t st, const float* c0 = nullptr);
int trans_upload_weights(ccsm_model* m);
int trans_forward(ccsm_model* m, int64_t n, const ccsm_strand* fwd, const ccsm_strand* rev, float* logits, float* probs,
                  cudaStream_t st);
void trans_release(ccsm_model* m);

// ---- tensor-core path (tc_path.cu)
int tc_upload_weights(ccsm_model* m);
void tc_release(ccsm_model* m);
int tc_forward_att2s(ccsm_model* m, int64_t n, const ccsm_strand* fwd, const ccsm_strand* rev,
                     const float* h0_f, const float* h0_r, float* logits, float* probs, cudaStream_t st);
int tc_debug_layer_out(ccsm_model* m, int layer, float* host, int64_t cap, int64_t* written);

// ---- fused aggregate model (aggr_fused.cu)
bool aggr_fused_supported(const ccsm_model* m);
int aggr_fused_upload(ccsm_model* m);
int aggr_fused_forward(ccsm_model* m, int64_t n, const float* offsets, const float* histos, const float* h0, float* out,
                       cudaStream_t st);
int aggr_fused_forward_sites(ccsm_model* m, int64_t n, const long long* site_pos, const float* site_histo, int only_close,
                             const float* h0, float* out, cudaStream_t st);
void pu_release(ccsm_model* m);

// ---- the reference's h0 stream on the device (mtstream.cu)
int mt_seed(ccsm_model* m, uint64_t seed);
int mt_set_state(ccsm_model* m, const uint32_t* words, int32_t pos);
int mt_get_state(ccsm_model* m, uint32_t* words, int32_t* pos);
int mt_set_batching(ccsm_model* m, const int64_t* counts, int64_t n_counts, int32_t batch_size);
int mt_take_segments(ccsm_model* m, int64_t n, std::vector<int64_t>& segs);
int mt_fill(ccsm_model* m, const int64_t* segs, int nseg, cudaStream_t user, const float** h0a, const float** h0b, int* buf);
int mt_release(ccsm_model* m, int buf, cudaStream_t user);
void mt_destroy(ccsm_model* m);
// att2s forward on device buffers; segs (nseg model-call site counts summing to n) only matters in CCSM_H0_TORCH_STREAM
// mode with no explicit h0: nullptr = take the announced hole-batches (ccsm_set_h0_batching) or one hole-batch of n
int forward_att2s_dev(ccsm_model* m, int64_t n, const ccsm_strand* fwd, const ccsm_strand* rev, const float* h0_fwd,
                      const float* h0_rev, float* logits, float* probs, cudaStream_t st, const int64_t* segs, int nseg);
// cuts consecutive model calls into groups of at most max_sites sites: (first segment, segment count, first site, sites)
struct SegChunk { int seg0, nseg; int64_t site0, sites; };
void seg_chunks(const std::vector<int64_t>& segs, int64_t max_sites, std::vector<SegChunk>& out);

// ---- device feature extraction (extract.cu)
void ex_release(ccsm_model* m);

}  // namespace ccsm
