// Fused aggregate-mode model (AggrAttRNN "attbigru": L = 11 neighbouring CpG sites, 21 inputs, H = 32, one
// bidirectional GRU layer, additive attention, fc1) -- ONE kernel per call.
//
// Replaces: reference ccsmeth/models.py:673-694 (AggrAttRNN.forward), utils/attention.py:48-70.
//
// The contractions here (K = 21 / 32 / 64, N = 96 / 32) are not tensor-core shaped, so this is an fp32 FFMA kernel
// (SURVEY.md section 8d).  Mapping: one thread per CpG site; every weight lives in shared memory in a
// [k][unit] layout, so one broadcast LDS.128 feeds four independent FFMA chains (four hidden units); the recurrent
// state of a thread stays in registers (read side) and in a per-thread shared-memory column (write side, which
// also serves the dynamic h[j] reads of the blend).  The GRU outputs a thread needs again for the attention
// (L x 64 floats) go through an L2-resident scratch slab owned by the CTA, laid out [t][unit][thread] so that all
// accesses are coalesced.  Because fc1 is linear, fc1(context) = sum_t w_t fc1(out_t): the attention pass keeps one
// scalar per step instead of the 64-float context vector.
//
// Per site: 137,632 FFMA, 924 B of windows + 256 B of h0 read, 4 B written.
#include <math.h>

#include "ccsm_internal.h"

namespace ccsm {

constexpr int AG_H = 32;           // hidden units (fixed by this kernel)
constexpr int AG_THREADS = 256;    // threads per CTA
constexpr int AG_S = 1;            // sites per thread (measured: 1 x 256 threads = 76 M sites/s, 2 x 128 = 72 M)
constexpr int AG_MAX_L = 16;

struct AggrPacked {
  // offsets into the packed float buffer
  int wx, wh, bias, wa, ua, va, fcw, fcb, total;
};

__host__ __device__ inline AggrPacked aggr_layout(int IN, int C) {
  AggrPacked p;
  int o = 0;
  p.wx = o;   o += 2 * 3 * IN * AG_H;      // [dir][gate r,z,n][k][unit]
  p.wh = o;   o += 2 * 3 * AG_H * AG_H;    // [dir][gate r,z,n][k][unit]
  p.bias = o; o += 2 * 4 * AG_H;           // [dir][b_r, b_z, b_in, b_hn][unit]
  p.wa = o;   o += 2 * AG_H * AG_H;        // [k 0..63][unit]   (query projection, transposed)
  p.ua = o;   o += 2 * AG_H * AG_H;        // [k 0..63][unit]   (key projection, transposed)
  p.va = o;   o += AG_H;
  p.fcw = o;  o += C * 2 * AG_H;           // [class][k 0..63]
  p.fcb = o;  o += 4;
  p.total = (o + 3) & ~3;
  return p;
}

__device__ __forceinline__ float sigmoid_acc(float x) { return 1.f / (1.f + expf(-x)); }

// S = sites per thread: every weight fetched from shared memory feeds S x 4 FFMA chains.  Measured on B200: the
// kernel sits at ~28 % of the FP32 pipe; a warp-wide LDS.128 occupies the shared-memory pipe for four passes even
// when every lane reads the same 16 bytes, so one weight fetch per 4 (S = 1) or 8 (S = 2, but half the warps) FFMAs
// is what binds.  The next step is a register-tiled formulation (8 sites x 2 units per thread, A and B from shared
// memory, 12 FFMA per LDS.128) -- DESIGN.md section 8.
template <int IN, int S>
__global__ void __launch_bounds__(AG_THREADS, 2)
    aggr_fused_kernel(const float* __restrict__ packed, int64_t n, int L, int C, const float* __restrict__ offsets,
                      const float* __restrict__ histos, const long long* __restrict__ site_pos, int only_close,
                      const float* __restrict__ h0, float* scratch, float* __restrict__ out) {
  extern __shared__ __align__(16) float sm[];
  const AggrPacked lay = aggr_layout(IN, C);
  constexpr int COLS = S * AG_THREADS;  // sites per CTA pass; column c = s * AG_THREADS + tid
  float* hcol = sm + lay.total;         // [unit][column]
  for (int i = threadIdx.x; i < lay.total; i += AG_THREADS) sm[i] = packed[i];
  __syncthreads();
  const int tid = threadIdx.x;
  constexpr int BINS = IN - 1;
  float* my_scratch = scratch + (size_t)blockIdx.x * L * 2 * AG_H * COLS;  // [t][unit 0..63][column]

  for (int64_t base = (int64_t)blockIdx.x * COLS; base < n; base += (int64_t)gridDim.x * COLS) {
    int64_t sidx[S];
#pragma unroll
    for (int q = 0; q < S; ++q) {
      const int64_t site = base + q * AG_THREADS + tid;
      sidx[q] = site < n ? site : n - 1;  // dead columns shadow the last site (no divergence, no stores)
    }
#pragma unroll 1
    for (int dir = 0; dir < 2; ++dir) {
      const float* WX = sm + lay.wx + dir * 3 * IN * AG_H;
      const float* WH = sm + lay.wh + dir * 3 * AG_H * AG_H;
      const float* BS = sm + lay.bias + dir * 4 * AG_H;
      float h[S][AG_H];
#pragma unroll
      for (int q = 0; q < S; ++q)
#pragma unroll
        for (int k = 0; k < AG_H; ++k) {
          h[q][k] = h0 ? h0[((size_t)dir * n + sidx[q]) * AG_H + k] : 0.f;
          hcol[k * COLS + q * AG_THREADS + tid] = h[q][k];
        }
#pragma unroll 1
      for (int step = 0; step < L; ++step) {
        const int t = dir ? L - 1 - step : step;
        float x[S][IN];
#pragma unroll
        for (int q = 0; q < S; ++q) {
          const float* hp = histos + ((size_t)sidx[q] * L + t) * BINS;
          if (site_pos) {
            // windows gathered in place from per-site rows (call_mods_freq_bam.py:272-283): neighbour j = i + t - L/2,
            // zero histogram and a position 1000 bp beyond the ends outside the region
            const int64_t j = sidx[q] + t - L / 2;
            const long long centre = site_pos[sidx[q]];
            // padded position of neighbour j (:279-280, 285-286)
            const long long pj = j < 0 ? site_pos[0] - 1000 : j >= n ? site_pos[n - 1] + 1000 : site_pos[j];
            float off;
            if (only_close) {
                // --only_close: 1 where the neighbour directly follows its predecessor as the next CpG (distance 2)
                const long long pjm = j - 1 < 0 ? site_pos[0] - 1000 : j - 1 >= n ? site_pos[n - 1] + 1000 : site_pos[j - 1];
                off = (pj - pjm == 2) ? 1.f : 0.f;
            } else {
                off = (float)llabs(pj - centre);
            }
            x[q][BINS] = off;
            if (j < 0 || j >= n) {
#pragma unroll
              for (int k = 0; k < BINS; ++k) x[q][k] = 0.f;
              continue;
            }
            hp = histos + (size_t)j * BINS;
          }
          if constexpr (BINS % 4 == 0) {
#pragma unroll
            for (int k = 0; k < BINS; k += 4) {
              const float4 v = __ldg(reinterpret_cast<const float4*>(hp + k));
              x[q][k] = v.x; x[q][k + 1] = v.y; x[q][k + 2] = v.z; x[q][k + 3] = v.w;
            }
          } else {
#pragma unroll
            for (int k = 0; k < BINS; ++k) x[q][k] = __ldg(hp + k);
          }
          if (!site_pos) x[q][BINS] = __ldg(offsets + (size_t)sidx[q] * L + t);  // cat(histos, offsets) (models.py:675-677)
        }
#pragma unroll 1
        for (int jb = 0; jb < AG_H / 4; ++jb) {
          const int j = jb * 4;
          float ar[S][4], az[S][4], ai[S][4], ah[S][4];
          {
            const float4 br = *reinterpret_cast<const float4*>(BS + 0 * AG_H + j);
            const float4 bz = *reinterpret_cast<const float4*>(BS + 1 * AG_H + j);
            const float4 bi = *reinterpret_cast<const float4*>(BS + 2 * AG_H + j);
            const float4 bh = *reinterpret_cast<const float4*>(BS + 3 * AG_H + j);
#pragma unroll
            for (int q = 0; q < S; ++q) {
              ar[q][0] = br.x; ar[q][1] = br.y; ar[q][2] = br.z; ar[q][3] = br.w;
              az[q][0] = bz.x; az[q][1] = bz.y; az[q][2] = bz.z; az[q][3] = bz.w;
              ai[q][0] = bi.x; ai[q][1] = bi.y; ai[q][2] = bi.z; ai[q][3] = bi.w;
              ah[q][0] = bh.x; ah[q][1] = bh.y; ah[q][2] = bh.z; ah[q][3] = bh.w;
            }
          }
#pragma unroll
          for (int k = 0; k < IN; ++k) {
            const float4 wr = *reinterpret_cast<const float4*>(WX + (0 * IN + k) * AG_H + j);
            const float4 wz = *reinterpret_cast<const float4*>(WX + (1 * IN + k) * AG_H + j);
            const float4 wn = *reinterpret_cast<const float4*>(WX + (2 * IN + k) * AG_H + j);
#pragma unroll
            for (int q = 0; q < S; ++q) {
              const float v = x[q][k];
              ar[q][0] = fmaf(wr.x, v, ar[q][0]); ar[q][1] = fmaf(wr.y, v, ar[q][1]);
              ar[q][2] = fmaf(wr.z, v, ar[q][2]); ar[q][3] = fmaf(wr.w, v, ar[q][3]);
              az[q][0] = fmaf(wz.x, v, az[q][0]); az[q][1] = fmaf(wz.y, v, az[q][1]);
              az[q][2] = fmaf(wz.z, v, az[q][2]); az[q][3] = fmaf(wz.w, v, az[q][3]);
              ai[q][0] = fmaf(wn.x, v, ai[q][0]); ai[q][1] = fmaf(wn.y, v, ai[q][1]);
              ai[q][2] = fmaf(wn.z, v, ai[q][2]); ai[q][3] = fmaf(wn.w, v, ai[q][3]);
            }
          }
#pragma unroll
          for (int k = 0; k < AG_H; ++k) {
            const float4 wr = *reinterpret_cast<const float4*>(WH + (0 * AG_H + k) * AG_H + j);
            const float4 wz = *reinterpret_cast<const float4*>(WH + (1 * AG_H + k) * AG_H + j);
            const float4 wn = *reinterpret_cast<const float4*>(WH + (2 * AG_H + k) * AG_H + j);
#pragma unroll
            for (int q = 0; q < S; ++q) {
              const float v = h[q][k];
              ar[q][0] = fmaf(wr.x, v, ar[q][0]); ar[q][1] = fmaf(wr.y, v, ar[q][1]);
              ar[q][2] = fmaf(wr.z, v, ar[q][2]); ar[q][3] = fmaf(wr.w, v, ar[q][3]);
              az[q][0] = fmaf(wz.x, v, az[q][0]); az[q][1] = fmaf(wz.y, v, az[q][1]);
              az[q][2] = fmaf(wz.z, v, az[q][2]); az[q][3] = fmaf(wz.w, v, az[q][3]);
              ah[q][0] = fmaf(wn.x, v, ah[q][0]); ah[q][1] = fmaf(wn.y, v, ah[q][1]);
              ah[q][2] = fmaf(wn.z, v, ah[q][2]); ah[q][3] = fmaf(wn.w, v, ah[q][3]);
            }
          }
#pragma unroll
          for (int q = 0; q < S; ++q)
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const float r = sigmoid_acc(ar[q][u]);
              const float z = sigmoid_acc(az[q][u]);
              const float nn = tanhf(fmaf(r, ah[q][u], ai[q][u]));
              const int col = q * AG_THREADS + tid;
              const float hprev = hcol[(j + u) * COLS + col];
              const float hn = fmaf(z, hprev - nn, nn);  // (1 - z) * n + z * h
              hcol[(j + u) * COLS + col] = hn;
              my_scratch[((size_t)t * 2 * AG_H + dir * AG_H + j + u) * COLS + col] = hn;
            }
        }
#pragma unroll
        for (int q = 0; q < S; ++q)
#pragma unroll
          for (int k = 0; k < AG_H; ++k) h[q][k] = hcol[k * COLS + q * AG_THREADS + tid];
      }
    }

    // ---- attention (utils/attention.py:48-70), one column at a time.  The query [h_n fwd | h_n rev] is the
    // forward output at t = L-1 and the reverse output at t = 0, both already in the scratch slab.
    const float* WA = sm + lay.wa;
    const float* UA = sm + lay.ua;
    const float* VA = sm + lay.va;
    const float* FW = sm + lay.fcw;
#pragma unroll 1
    for (int q = 0; q < S; ++q) {
      const int col = q * AG_THREADS + tid;
      float wq[AG_H];
#pragma unroll
      for (int i = 0; i < AG_H; ++i) wq[i] = 0.f;
#pragma unroll 4
      for (int k = 0; k < 2 * AG_H; ++k) {
        const int tq = k < AG_H ? L - 1 : 0;
        const float v = my_scratch[((size_t)tq * 2 * AG_H + k) * COLS + col];
#pragma unroll
        for (int i = 0; i < AG_H; i += 4) {
          const float4 w = *reinterpret_cast<const float4*>(WA + k * AG_H + i);
          wq[i] = fmaf(w.x, v, wq[i]); wq[i + 1] = fmaf(w.y, v, wq[i + 1]);
          wq[i + 2] = fmaf(w.z, v, wq[i + 2]); wq[i + 3] = fmaf(w.w, v, wq[i + 3]);
        }
      }
      float e[AG_MAX_L], g[AG_MAX_L];
#pragma unroll 1
      for (int t = 0; t < L; ++t) {
        float acc[AG_H];
#pragma unroll
        for (int i = 0; i < AG_H; ++i) acc[i] = wq[i];
        float gc = 0.f;  // fc1 . out_t (num_classes == 1)
#pragma unroll 4
        for (int k = 0; k < 2 * AG_H; ++k) {
          const float v = my_scratch[((size_t)t * 2 * AG_H + k) * COLS + col];
#pragma unroll
          for (int i = 0; i < AG_H; i += 4) {
            const float4 w = *reinterpret_cast<const float4*>(UA + k * AG_H + i);
            acc[i] = fmaf(w.x, v, acc[i]); acc[i + 1] = fmaf(w.y, v, acc[i + 1]);
            acc[i + 2] = fmaf(w.z, v, acc[i + 2]); acc[i + 3] = fmaf(w.w, v, acc[i + 3]);
          }
          gc = fmaf(FW[k], v, gc);
        }
        float et = 0.f;
#pragma unroll
        for (int i = 0; i < AG_H; ++i) et = fmaf(VA[i], tanhf(acc[i]), et);
        // static indexing keeps e / g in registers
#pragma unroll
        for (int w = 0; w < AG_MAX_L; ++w)
          if (w == t) {
            e[w] = et;
            g[w] = gc;
          }
      }
      float mx = -INFINITY;
#pragma unroll
      for (int w = 0; w < AG_MAX_L; ++w)
        if (w < L) mx = fmaxf(mx, e[w]);
      float den = 0.f, num = 0.f;
#pragma unroll
      for (int w = 0; w < AG_MAX_L; ++w)
        if (w < L) {
          const float p = expf(e[w] - mx);
          den += p;
          num = fmaf(p, g[w], num);
        }
      const int64_t site = base + col;
      if (site < n) out[site] = num / den + sm[lay.fcb];
    }
  }
}

// ---- host side ----------------------------------------------------------------------------------------------
static const HostTensor* findw(ccsm_model* m, const std::string& k) {
  auto it = m->w.find(k);
  return it == m->w.end() ? nullptr : &it->second;
}

bool aggr_fused_supported(const ccsm_model* m) {
  return m->cfg.kind == CCSM_KIND_AGGR && m->cfg.hidden == AG_H && m->cfg.num_layers == 1 && m->cfg.seq_len <= AG_MAX_L &&
         m->cfg.num_classes == 1 && (m->in_feat == 21);
}

int aggr_fused_upload(ccsm_model* m) {
  if (!aggr_fused_supported(m)) return CCSM_OK;
  const int IN = m->in_feat, C = m->cfg.num_classes, H = AG_H;
  const AggrPacked lay = aggr_layout(IN, C);
  std::vector<float> p((size_t)lay.total, 0.f);
  for (int d = 0; d < 2; ++d) {
    const std::string sfx = std::string("_l0") + (d ? "_reverse" : "");
    const HostTensor *wih = findw(m, "rnn.weight_ih" + sfx), *whh = findw(m, "rnn.weight_hh" + sfx),
                     *bih = findw(m, "rnn.bias_ih" + sfx), *bhh = findw(m, "rnn.bias_hh" + sfx);
    if (!wih || !whh || !bih || !bhh) {
      set_error("aggr_fused_upload: missing GRU tensor");
      return CCSM_EKEY;
    }
    for (int gate = 0; gate < 3; ++gate)  // PyTorch row blocks (r, z, n)
      for (int j = 0; j < H; ++j) {
        for (int k = 0; k < IN; ++k)
          p[lay.wx + ((d * 3 + gate) * IN + k) * H + j] = wih->data[(size_t)(gate * H + j) * IN + k];
        for (int k = 0; k < H; ++k)
          p[lay.wh + ((d * 3 + gate) * H + k) * H + j] = whh->data[(size_t)(gate * H + j) * H + k];
      }
    for (int j = 0; j < H; ++j) {
      p[lay.bias + (d * 4 + 0) * H + j] = bih->data[j] + bhh->data[j];
      p[lay.bias + (d * 4 + 1) * H + j] = bih->data[H + j] + bhh->data[H + j];
      p[lay.bias + (d * 4 + 2) * H + j] = bih->data[2 * H + j];
      p[lay.bias + (d * 4 + 3) * H + j] = bhh->data[2 * H + j];
    }
  }
  const HostTensor *wa = findw(m, "_att3.Wa.weight"), *ua = findw(m, "_att3.Ua.weight"), *va = findw(m, "_att3.va.weight"),
                   *fw = findw(m, "fc1.weight"), *fb = findw(m, "fc1.bias");
  if (!wa || !ua || !va || !fw || !fb) {
    set_error("aggr_fused_upload: missing attention / fc tensor");
    return CCSM_EKEY;
  }
  for (int i = 0; i < H; ++i)
    for (int k = 0; k < 2 * H; ++k) {
      p[lay.wa + k * H + i] = wa->data[(size_t)i * 2 * H + k];
      p[lay.ua + k * H + i] = ua->data[(size_t)i * 2 * H + k];
    }
  for (int i = 0; i < H; ++i) p[lay.va + i] = va->data[i];
  for (int c = 0; c < C; ++c) {
    for (int k = 0; k < 2 * H; ++k) p[lay.fcw + c * 2 * H + k] = fw->data[(size_t)c * 2 * H + k];
    p[lay.fcb + c] = fb->data[c];
  }
  CCSM_TRY(m->aggr_packed.reserve(p.size() * sizeof(float)));
  CCSM_CUDA(cudaMemcpy(m->aggr_packed.p, p.data(), p.size() * sizeof(float), cudaMemcpyHostToDevice));
  return CCSM_OK;
}

static int aggr_launch(ccsm_model* m, int64_t n, const float* offsets, const float* histos, const long long* site_pos,
                       int only_close, const float* h0, float* out, cudaStream_t st);

int aggr_fused_forward(ccsm_model* m, int64_t n, const float* offsets, const float* histos, const float* h0, float* out,
                       cudaStream_t st) {
  return aggr_launch(m, n, offsets, histos, nullptr, 0, h0, out, st);
}

// Same model over per-site rows: site_histo (n, bins) and site_pos (n) of consecutive sites; the 11-site windows and
// their |position offsets| are formed inside the kernel (84 B read per site instead of 924 B).
int aggr_fused_forward_sites(ccsm_model* m, int64_t n, const long long* site_pos, const float* site_histo, int only_close,
                             const float* h0, float* out, cudaStream_t st) {
  return aggr_launch(m, n, nullptr, site_histo, site_pos, only_close, h0, out, st);
}

static int aggr_launch(ccsm_model* m, int64_t n, const float* offsets, const float* histos, const long long* site_pos,
                       int only_close, const float* h0, float* out, cudaStream_t st) {
  const int IN = m->in_feat, C = m->cfg.num_classes;
  const AggrPacked lay = aggr_layout(IN, C);
  constexpr int COLS = AG_S * AG_THREADS;
  const size_t smem = ((size_t)lay.total + (size_t)AG_H * COLS) * sizeof(float);
  static bool attr = false;
  if (!attr) {
    CCSM_CUDA(cudaFuncSetAttribute(aggr_fused_kernel<21, AG_S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = true;
  }
  int sms = 0;
  CCSM_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, m->cfg.device));
  const int64_t tiles = (n + COLS - 1) / COLS;
  const int grid = (int)(tiles < 2LL * sms ? tiles : 2LL * sms);
  CCSM_TRY(m->aggr_scratch.reserve((size_t)grid * m->cfg.seq_len * 2 * AG_H * COLS * sizeof(float)));
  aggr_fused_kernel<21, AG_S><<<grid, AG_THREADS, smem, st>>>(m->aggr_packed.as<float>(), n, m->cfg.seq_len, C, offsets, histos,
                                                              site_pos, only_close, h0, m->aggr_scratch.as<float>(), out);
  count_launch();
  CCSM_CUDA(cudaGetLastError());
  return CCSM_OK;
}

}  // namespace ccsm
