// call_freqb on the device: one region's pileup (per reference position, the ML bytes and haplotypes of the reads
// covering it) -> per-site (coverage, modified count, modification frequency) for all reads / haplotype 1 /
// haplotype 2, in count mode or aggregate mode.
//
// Replaces (reference ccsmeth/call_mods_freq_bam.py):
//   :102-107  _cal_mod_prob                      ML byte -> probability (a 256-entry table here)
//   :200-217  _cal_modfreq_in_count_mode
//   :221-237  _get_normalized_histo              20-bin histogram, L2-normalised, rounded to 6 decimals
//   :265-305  _cal_modfreq_in_aggregate_mode     11-site windows + |position offsets| -> AggrAttRNN -> clip/round
//   :308-442  _call_modfreq_of_one_region(_aggregate_mode)   low/high coverage split, three read groups
// HBM-bound integer/byte kernels around the fused aggregate model (aggr_fused.cu), which gathers its windows from
// the compact per-site rows written here (84 B per site instead of 924 B of materialised windows).
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "ccsm_internal.h"

namespace ccsm {

struct PileupLut {
  double prob[256];   // _cal_mod_prob(ml)
  uint8_t bin[256];   // np.histogram(prob, bins=B, range=[0,1]) bin of that probability
  uint8_t amb[256];   // abs(p - (1 - p)) < prob_cf
  uint8_t mod[256];   // p > 0.5
};

struct PuState {
  DevBuf pos, ptr, ml, hap, lut;
  DevBuf cov, filt, nmod, flag, cidx, blocksum;          // per (group, site)
  DevBuf c_pos, c_histo, c_site, c_out;                  // compact high-coverage lists of the three groups
  DevBuf r_cov, r_cnt, r_freq, r_kind, h0, c0;
  DevBuf w_off, w_histo;                                 // materialised windows (models outside the fused kernel)
  ccsm_pileup_opts opts{};
  int64_t n = -1;
  int64_t n_high[3] = {0, 0, 0};
  int bins = 20;
};

// Python's round(x, 6) on a float: correctly rounded decimal, ties to even on the exact binary value -- glibc's printf
// performs the same conversion.
static double py_round6(double x) {
  char buf[64];
  snprintf(buf, sizeof(buf), "%.6f", x);
  return strtod(buf, nullptr);
}

void pileup_build_lut(const ccsm_pileup_opts& o, int bins, PileupLut& L) {
  // np.linspace(0, 1, bins + 1): arange * step + start, last edge forced to stop
  std::vector<double> edge((size_t)bins + 1);
  const double step = 1.0 / bins;
  for (int i = 0; i <= bins; ++i) edge[(size_t)i] = i * step;
  edge[(size_t)bins] = 1.0;
  for (int v = 0; v < 256; ++v) {
    const double p = v > 0 ? py_round6(v / 256.0 + 0.000001) : 0.0;   // :102-107
    L.prob[v] = p;
    // numpy's uniform-bin fast path (numpy/lib/_histograms_impl.py): scale, truncate, then fix up against the edges
    double f = ((p - 0.0) / (1.0 - 0.0)) * bins;
    long idx = (long)f;
    if (idx == bins) idx -= 1;
    if (p < edge[(size_t)idx]) idx -= 1;
    else if (p >= edge[(size_t)idx + 1] && idx != bins - 1) idx += 1;
    L.bin[v] = (uint8_t)idx;
    L.amb[v] = fabs(p - (1 - p)) < o.prob_cf ? 1 : 0;               // :203
    L.mod[v] = p > 0.5 ? 1 : 0;                                     // :206
  }
}

__device__ __forceinline__ bool in_group(int g, int hap) { return g == 0 || hap == g; }

// ---- kernel 1: per (group, site) coverage / filtered count / modified count; count-mode results; high-coverage flag
__global__ void __launch_bounds__(256) pileup_count_kernel(int64_t n, const long long* __restrict__ ptr,
                                                           const uint8_t* __restrict__ ml, const uint8_t* __restrict__ hap,
                                                           const PileupLut* __restrict__ lut, ccsm_pileup_opts o,
                                                           int* __restrict__ flag, int* __restrict__ r_cov,
                                                           double* __restrict__ r_cnt, double* __restrict__ r_freq,
                                                           uint8_t* __restrict__ r_kind) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= 3 * n) return;
  const int g = (int)(idx / n);
  const int64_t i = idx - (int64_t)g * n;
  int cov = 0, filt = 0, nmod = 0;
  if (g == 0 || !o.no_hap) {
    for (long long k = ptr[i]; k < ptr[i + 1]; ++k) {
      if (!in_group(g, hap ? hap[k] : 0)) continue;
      const int v = ml[k];
      ++cov;
      if (lut->amb[v]) continue;
      ++filt;
      nmod += lut->mod[v];
    }
  }
  const bool high = o.call_mode == 1 && cov >= o.cov_cf && cov > 0;
  flag[idx] = high ? 1 : 0;
  // which of the reference's value types the result has: 0 None, 1 count path with an integer count, 2 count path
  // with np.round(len * freq, 2), 3 model path (float32 values)
  r_kind[idx] = cov == 0 ? 0 : high ? 3 : (o.no_amb_cov || filt == cov) ? 1 : 2;
  if (cov == 0) {
    r_cov[idx] = -1;  // "None" for this group (a coverage of 0 is a legitimate --no_amb_cov result)
    r_cnt[idx] = 0.0;
    r_freq[idx] = 0.0;
  } else if (!high) {
    // _cal_modfreq_in_count_mode (:200-217)
    const double freq = filt > 0 ? __ddiv_rn((double)nmod, (double)filt) : 0.0;
    if (o.no_amb_cov) {
      r_cov[idx] = filt;
      r_cnt[idx] = (double)nmod;
    } else {
      r_cov[idx] = cov;
      // np.round(len * modfreq, 2) when some calls were dropped as ambiguous
      r_cnt[idx] = filt != cov ? __ddiv_rn(rint(__dmul_rn(__dmul_rn((double)cov, freq), 100.0)), 100.0) : (double)nmod;
    }
    r_freq[idx] = freq;
  } else {
    r_cov[idx] = cov;
  }
}

// ---- exclusive scan of int flags (three phases)
constexpr int SCAN_BLOCK = 1024;
__global__ void __launch_bounds__(SCAN_BLOCK) scan_block_kernel(const int* __restrict__ in, long long* __restrict__ out,
                                                                long long* __restrict__ block_sum, int64_t n) {
  __shared__ long long s_warp[32];
  const int64_t i = (int64_t)blockIdx.x * SCAN_BLOCK + threadIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long v = i < n ? in[i] : 0;
  long long x = v;
  for (int o = 1; o < 32; o <<= 1) {
    const long long y = __shfl_up_sync(0xffffffffu, x, o);
    if (lane >= o) x += y;
  }
  if (lane == 31) s_warp[warp] = x;
  __syncthreads();
  if (warp == 0) {
    long long w = s_warp[lane];
    for (int o = 1; o < 32; o <<= 1) {
      const long long y = __shfl_up_sync(0xffffffffu, w, o);
      if (lane >= o) w += y;
    }
    s_warp[lane] = w;
  }
  __syncthreads();
  const long long incl = x + (warp ? s_warp[warp - 1] : 0);
  if (i < n) out[i] = incl - v;
  if (threadIdx.x == SCAN_BLOCK - 1) block_sum[blockIdx.x] = incl;
}
__global__ void __launch_bounds__(SCAN_BLOCK) scan_sums_kernel(long long* __restrict__ block_sum, int nb,
                                                               long long* __restrict__ total) {
  // one CTA: in-place exclusive scan of the block sums (nb is n / 1024: small)
  __shared__ long long s_carry;
  __shared__ long long s_warp[32];
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int base = 0; base < nb; base += SCAN_BLOCK) {
    const int i = base + threadIdx.x;
    const long long v = i < nb ? block_sum[i] : 0;
    long long x = v;
    for (int o = 1; o < 32; o <<= 1) {
      const long long y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += y;
    }
    if (lane == 31) s_warp[warp] = x;
    __syncthreads();
    if (warp == 0) {
      long long w = s_warp[lane];
      for (int o = 1; o < 32; o <<= 1) {
        const long long y = __shfl_up_sync(0xffffffffu, w, o);
        if (lane >= o) w += y;
      }
      s_warp[lane] = w;
    }
    __syncthreads();
    const long long excl = s_carry + (warp ? s_warp[warp - 1] : 0) + x - v;
    if (i < nb) block_sum[i] = excl;
    __syncthreads();
    if (threadIdx.x == SCAN_BLOCK - 1) s_carry = excl + v;
    __syncthreads();
  }
  if (threadIdx.x == 0) *total = s_carry;
}
__global__ void __launch_bounds__(SCAN_BLOCK) scan_add_kernel(long long* __restrict__ out, const long long* __restrict__ block_sum,
                                                              int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * SCAN_BLOCK + threadIdx.x;
  if (i < n) out[i] += block_sum[blockIdx.x];
}

// ---- kernel 2: high-coverage sites -> compact rows: position, source site, normalised histogram (:221-237)
__global__ void __launch_bounds__(128) pileup_histo_kernel(int64_t n, int g, int bins, const long long* __restrict__ pos,
                                                           const long long* __restrict__ ptr, const uint8_t* __restrict__ ml,
                                                           const uint8_t* __restrict__ hap, const PileupLut* __restrict__ lut,
                                                           const int* __restrict__ flag, const long long* __restrict__ cidx,
                                                           long long* __restrict__ c_pos, int* __restrict__ c_site,
                                                           float* __restrict__ c_histo) {
  __shared__ int s_h[32][128];  // [bin][thread]: conflict-free per-thread histograms
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || !flag[(int64_t)g * n + i]) return;
  for (int b = 0; b < bins; ++b) s_h[b][threadIdx.x] = 0;
  for (long long k = ptr[i]; k < ptr[i + 1]; ++k) {
    if (!in_group(g, hap ? hap[k] : 0)) continue;
    s_h[lut->bin[ml[k]]][threadIdx.x] += 1;
  }
  long long ss = 0;
  for (int b = 0; b < bins; ++b) ss += (long long)s_h[b][threadIdx.x] * s_h[b][threadIdx.x];
  const double norm = __dsqrt_rn((double)ss);  // np.linalg.norm of the integer histogram
  const long long c = cidx[(int64_t)g * n + i];
  c_pos[c] = pos[i];
  c_site[c] = (int)i;
  for (int b = 0; b < bins; ++b) {
    const double q = __ddiv_rn((double)s_h[b][threadIdx.x], norm);
    c_histo[c * bins + b] = (float)__ddiv_rn(rint(__dmul_rn(q, 1e6)), 1e6);  // np.round(hist / norm, 6)
  }
}

// ---- kernel 4: model output -> (cov, cnt_mod, freq) of the site (:386-395)
__global__ void __launch_bounds__(256) pileup_finish_kernel(int64_t nc, int64_t n, int g, const float* __restrict__ out,
                                                            const int* __restrict__ c_site, const int* __restrict__ r_cov,
                                                            double* __restrict__ r_cnt, double* __restrict__ r_freq) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= nc) return;
  const int64_t idx = (int64_t)g * n + c_site[c];
  // np.round(np.clip(out, 0, 1), 6) in float32 (:302)
  const float p = __fdiv_rn(rintf(__fmul_rn(fminf(fmaxf(out[c], 0.f), 1.f), 1e6f)), 1e6f);
  // cnt_mod = round(cov * modprob, 2): int * np.float32 -> float32, rounded in float32
  const float cm = __fdiv_rn(rintf(__fmul_rn(__fmul_rn((float)r_cov[idx], p), 100.f)), 100.f);
  r_cnt[idx] = (double)cm;
  r_freq[idx] = (double)p;
}

// Materialised windows for the models the fused kernel does not cover (other hidden sizes / layer counts, the LSTM
// cell): offsets (nc, L) and histos (nc, L, bins) exactly as call_mods_freq_bam.py:272-290 pads and slides them --
// neighbour j = i + t - L/2, zero histogram and a position 1000 bp beyond the region's ends outside it.
__global__ void pileup_windows_kernel(int64_t nc, int L, int bins, int only_close, const long long* __restrict__ pos,
                                      const float* __restrict__ histo, float* __restrict__ offsets,
                                      float* __restrict__ windows) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // over (site, t)
  if (idx >= nc * L) return;
  const int t = (int)(idx % L);
  const int64_t i = idx / L, j = i + t - L / 2;
  const long long pj = j < 0 ? pos[0] - 1000 : j >= nc ? pos[nc - 1] + 1000 : pos[j];
  float off;
  if (only_close) {
    const long long pjm = j - 1 < 0 ? pos[0] - 1000 : j - 1 >= nc ? pos[nc - 1] + 1000 : pos[j - 1];
    off = (pj - pjm == 2) ? 1.f : 0.f;
  } else {
    off = (float)llabs(pj - pos[i]);
  }
  offsets[idx] = off;
  float* w = windows + idx * bins;
  const bool inside = j >= 0 && j < nc;
  for (int b = 0; b < bins; ++b) w[b] = inside ? histo[j * bins + b] : 0.f;
}

void pu_release(ccsm_model* m) {
  PuState* s = m->pu;
  if (!s) return;
  for (DevBuf* b : {&s->pos, &s->ptr, &s->ml, &s->hap, &s->lut, &s->cov, &s->filt, &s->nmod, &s->flag, &s->cidx,
                    &s->blocksum, &s->c_pos, &s->c_histo, &s->c_site, &s->c_out, &s->r_cov, &s->r_cnt, &s->r_freq, &s->r_kind, &s->h0, &s->c0, &s->w_off, &s->w_histo})
    b->release();
  delete s;
  m->pu = nullptr;
}

}  // namespace ccsm

using namespace ccsm;

extern "C" {

int ccsm_pileup_luts(const ccsm_pileup_opts* o, int32_t bins, double* prob, int32_t* bin) {
  if (!o || !prob || !bin || bins < 1 || bins > 32) {
    set_error("ccsm_pileup_luts: bad argument");
    return CCSM_EINVAL;
  }
  PileupLut L;
  pileup_build_lut(*o, bins, L);
  for (int v = 0; v < 256; ++v) {
    prob[v] = L.prob[v];
    bin[v] = L.bin[v];
  }
  return CCSM_OK;
}

int ccsm_pileup_begin_host(ccsm_model* m, const ccsm_pileup_opts* o, int64_t n, const int64_t* refpos, const int64_t* ptr,
                           const uint8_t* ml, const uint8_t* hap, int64_t* n_high) {
  if (!m || m->cfg.kind != CCSM_KIND_AGGR) {
    set_error("ccsm_pileup_begin_host: needs an aggregate (CCSM_KIND_AGGR) model handle");
    return CCSM_EINVAL;
  }
  if (!o || n < 0 || !n_high || (n > 0 && (!refpos || !ptr || !ml))) {
    set_error("ccsm_pileup_begin_host: bad argument");
    return CCSM_EINVAL;
  }
  if (o->call_mode != 0 && o->call_mode != 1) {
    set_error("ccsm_pileup_begin_host: call_mode must be 0 (count) or 1 (aggregate)");
    return CCSM_EINVAL;
  }
  if (o->call_mode == 1 && (!m->finalized || m->cfg.num_classes != 1 || m->cfg.feat_flags > 32)) {
    set_error("ccsm_pileup_begin_host: aggregate mode needs a finalized aggregate model (one output, <= 32 bins)");
    return CCSM_ESTATE;
  }
  CCSM_CUDA(cudaSetDevice(m->cfg.device));
  if (!m->pu) m->pu = new (std::nothrow) PuState();
  PuState* s = m->pu;
  if (!s) return CCSM_ENOMEM;
  s->opts = *o;
  s->n = n;
  s->bins = m->cfg.feat_flags;
  for (int g = 0; g < 3; ++g) s->n_high[g] = n_high[g] = 0;
  if (n == 0) return CCSM_OK;
  const int64_t total = ptr[n];
  if (ptr[0] != 0 || total < 0) {
    set_error("ccsm_pileup_begin_host: ptr must start at 0 and be non-decreasing");
    return CCSM_EINVAL;
  }
  cudaStream_t st = nullptr;
  PileupLut L;
  pileup_build_lut(*o, s->bins, L);
  CCSM_TRY(s->lut.reserve(sizeof(PileupLut)));
  CCSM_TRY(s->pos.reserve((size_t)n * 8));
  CCSM_TRY(s->ptr.reserve((size_t)(n + 1) * 8));
  CCSM_TRY(s->ml.reserve((size_t)total + 16));
  if (hap) CCSM_TRY(s->hap.reserve((size_t)total + 16));
  CCSM_TRY(s->flag.reserve((size_t)3 * n * 4));
  CCSM_TRY(s->cidx.reserve((size_t)3 * n * 8));
  CCSM_TRY(s->r_cov.reserve((size_t)3 * n * 4));
  CCSM_TRY(s->r_cnt.reserve((size_t)3 * n * 8));
  CCSM_TRY(s->r_freq.reserve((size_t)3 * n * 8));
  CCSM_TRY(s->r_kind.reserve((size_t)3 * n + 16));
  const int nb = (int)((n + SCAN_BLOCK - 1) / SCAN_BLOCK);
  CCSM_TRY(s->blocksum.reserve((size_t)(nb + 4) * 8));
  CCSM_CUDA(cudaMemcpyAsync(s->lut.p, &L, sizeof(L), cudaMemcpyHostToDevice, st));
  CCSM_CUDA(cudaMemcpyAsync(s->pos.p, refpos, (size_t)n * 8, cudaMemcpyHostToDevice, st));
  CCSM_CUDA(cudaMemcpyAsync(s->ptr.p, ptr, (size_t)(n + 1) * 8, cudaMemcpyHostToDevice, st));
  CCSM_CUDA(cudaMemcpyAsync(s->ml.p, ml, (size_t)total, cudaMemcpyHostToDevice, st));
  if (hap) CCSM_CUDA(cudaMemcpyAsync(s->hap.p, hap, (size_t)total, cudaMemcpyHostToDevice, st));
  pileup_count_kernel<<<(unsigned)((3 * n + 255) / 256), 256, 0, st>>>(
      n, s->ptr.as<long long>(), s->ml.as<uint8_t>(), hap ? s->hap.as<uint8_t>() : nullptr, s->lut.as<PileupLut>(), *o,
      s->flag.as<int>(), s->r_cov.as<int>(), s->r_cnt.as<double>(), s->r_freq.as<double>(), s->r_kind.as<uint8_t>());
  count_launch();
  long long totals[3] = {0, 0, 0};
  if (o->call_mode == 1) {
    for (int g = 0; g < 3; ++g) {
      long long* bs = s->blocksum.as<long long>();
      scan_block_kernel<<<nb, SCAN_BLOCK, 0, st>>>(s->flag.as<int>() + (size_t)g * n, s->cidx.as<long long>() + (size_t)g * n,
                                                   bs, n);
      scan_sums_kernel<<<1, SCAN_BLOCK, 0, st>>>(bs, nb, bs + nb + 1);
      scan_add_kernel<<<nb, SCAN_BLOCK, 0, st>>>(s->cidx.as<long long>() + (size_t)g * n, bs, n);
      count_launch(3);
      CCSM_CUDA(cudaMemcpyAsync(&totals[g], bs + nb + 1, 8, cudaMemcpyDeviceToHost, st));
      CCSM_CUDA(cudaStreamSynchronize(st));
    }
  }
  CCSM_CUDA(cudaStreamSynchronize(st));
  CCSM_CUDA(cudaGetLastError());
  for (int g = 0; g < 3; ++g) s->n_high[g] = n_high[g] = totals[g];
  return CCSM_OK;
}

}  // extern "C"

static int pileup_finish_impl(ccsm_model* m, const float* const* h0s, const float* const* c0s, int32_t* cov,
                              double* cnt_mod, double* freq, uint8_t* kind) {
  if (!m || !m->pu || m->pu->n < 0) {
    set_error("ccsm_pileup_finish_host: no resident pileup (call ccsm_pileup_begin_host first)");
    return CCSM_ESTATE;
  }
  PuState* s = m->pu;
  const int64_t n = s->n;
  if (n == 0) return CCSM_OK;
  if (!cov || !cnt_mod || !freq) {
    set_error("ccsm_pileup_finish_host: null output");
    return CCSM_EINVAL;
  }
  CCSM_CUDA(cudaSetDevice(m->cfg.device));
  cudaStream_t st = nullptr;
  const int bins = s->bins, H = m->cfg.hidden, L = m->cfg.seq_len;
  const size_t state_floats = (size_t)2 * m->cfg.num_layers * H;  // per site: (2 * layers, n, hidden)
  const bool fused = aggr_fused_supported(m);
  for (int g = 0; g < 3 && s->opts.call_mode == 1; ++g) {
    const int64_t nc = s->n_high[g];
    if (nc == 0) continue;
    CCSM_TRY(s->c_pos.reserve((size_t)nc * 8));
    CCSM_TRY(s->c_site.reserve((size_t)nc * 4));
    CCSM_TRY(s->c_histo.reserve((size_t)nc * bins * 4));
    CCSM_TRY(s->c_out.reserve((size_t)nc * 4));
    pileup_histo_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(
        n, g, bins, s->pos.as<long long>(), s->ptr.as<long long>(), s->ml.as<uint8_t>(),
        s->hap.p ? s->hap.as<uint8_t>() : nullptr, s->lut.as<PileupLut>(), s->flag.as<int>(), s->cidx.as<long long>(),
        s->c_pos.as<long long>(), s->c_site.as<int>(), s->c_histo.as<float>());
    count_launch();
    const float* dh0 = nullptr;
    if (h0s[g]) {
      CCSM_TRY(s->h0.reserve(state_floats * nc * 4));
      CCSM_CUDA(cudaMemcpyAsync(s->h0.p, h0s[g], state_floats * nc * 4, cudaMemcpyHostToDevice, st));
      dh0 = s->h0.as<float>();
    }
    if (fused) {
      CCSM_TRY(aggr_fused_forward_sites(m, nc, s->c_pos.as<long long>(), s->c_histo.as<float>(),
                                        s->opts.only_close ? 1 : 0, dh0, s->c_out.as<float>(), st));
    } else {
      // other shapes / the LSTM cell: windows are materialised, then the layer-by-layer fp32 kernels run
      const float* dc0 = nullptr;
      if (m->gates == 4 && c0s && c0s[g]) {
        CCSM_TRY(s->c0.reserve(state_floats * nc * 4));
        CCSM_CUDA(cudaMemcpyAsync(s->c0.p, c0s[g], state_floats * nc * 4, cudaMemcpyHostToDevice, st));
        dc0 = s->c0.as<float>();
      }
      CCSM_TRY(s->w_off.reserve((size_t)nc * L * 4));
      CCSM_TRY(s->w_histo.reserve((size_t)nc * L * bins * 4));
      pileup_windows_kernel<<<(unsigned)((nc * L + 255) / 256), 256, 0, st>>>(
          nc, L, bins, s->opts.only_close ? 1 : 0, s->c_pos.as<long long>(), s->c_histo.as<float>(), s->w_off.as<float>(),
          s->w_histo.as<float>());
      count_launch();
      CCSM_TRY(fp32_forward_aggr(m, nc, s->w_off.as<float>(), s->w_histo.as<float>(), dh0, s->c_out.as<float>(), st, dc0));
    }
    pileup_finish_kernel<<<(unsigned)((nc + 255) / 256), 256, 0, st>>>(nc, n, g, s->c_out.as<float>(), s->c_site.as<int>(),
                                                                        s->r_cov.as<int>(), s->r_cnt.as<double>(),
                                                                        s->r_freq.as<double>());
    count_launch();
    CCSM_CUDA(cudaStreamSynchronize(st));  // c_* buffers are reused by the next group
  }
  CCSM_CUDA(cudaMemcpyAsync(cov, s->r_cov.p, (size_t)3 * n * 4, cudaMemcpyDeviceToHost, st));
  CCSM_CUDA(cudaMemcpyAsync(cnt_mod, s->r_cnt.p, (size_t)3 * n * 8, cudaMemcpyDeviceToHost, st));
  CCSM_CUDA(cudaMemcpyAsync(freq, s->r_freq.p, (size_t)3 * n * 8, cudaMemcpyDeviceToHost, st));
  if (kind) CCSM_CUDA(cudaMemcpyAsync(kind, s->r_kind.p, (size_t)3 * n, cudaMemcpyDeviceToHost, st));
  CCSM_CUDA(cudaStreamSynchronize(st));
  CCSM_CUDA(cudaGetLastError());
  return CCSM_OK;
}

extern "C" {

int ccsm_pileup_finish_host(ccsm_model* m, const float* h0_all, const float* h0_hp1, const float* h0_hp2, int32_t* cov,
                            double* cnt_mod, double* freq, uint8_t* kind) {
  const float* h0s[3] = {h0_all, h0_hp1, h0_hp2};
  return pileup_finish_impl(m, h0s, nullptr, cov, cnt_mod, freq, kind);
}

int ccsm_pileup_finish_lstm_host(ccsm_model* m, const float* const* h0, const float* const* c0, int32_t* cov,
                                 double* cnt_mod, double* freq, uint8_t* kind) {
  if (!h0 || !c0) {
    set_error("ccsm_pileup_finish_lstm_host: h0 / c0 must point at three (possibly NULL) state pointers");
    return CCSM_EINVAL;
  }
  if (m && m->gates != 4) {
    set_error("ccsm_pileup_finish_lstm_host: not an LSTM aggregate model (CCSM_AGGR_LSTM)");
    return CCSM_EINVAL;
  }
  return pileup_finish_impl(m, h0, c0, cov, cnt_mod, freq, kind);
}

}  // extern "C"
