// fp32 (FFMA) implementation of the attbigru2s / aggregate forwards for sm_100a.
//
// This is the reference-exact arithmetic mode (CCSM_PREC_FP32): every contraction is an fp32 FFMA
// GEMM, gate math uses expf/tanhf.  It exists (a) as the first parity-green CUDA path and (b) as
// the on-device cross-check for the tcgen05 path in tc_path.cu, which is the throughput product.
//
// What it computes, per chunk of `rows` = strands * sites strand-rows (row R = site * strands + strand):
//   pack_x        x0[R][t][:]   = [embed[kmer] | ipd | pw | npass ...]         reference models.py:91-123
//   per layer l:  gi            = x_l . [W_ih_fwd ; W_ih_rev]^T + b_ih          (one GEMM for all t, both dirs)
//     per step s: gh[d]         = h[d] . W_hh[d]^T + b_hh[d]                    (batched GEMM, d = fwd/rev)
//                 gate math     r,z,n,h' (PyTorch GRU cell, gate order r,z,n)   reference models.py:125-130
//   attention     e = out . Ua^T, qa = q . Wa^T, softmax_t(va . tanh(qa + e_t)) reference utils/attention.py:48-70
//   head          ctx -> fc1 -> softmax                                          reference models.py:145-150
#include <curand_kernel.h>
#include <math.h>
#include <stdio.h>

#include "ccsm_internal.h"

namespace ccsm {

// ------------------------------------------------------------------------------------------------
// C[M,N] = A[M,K] . B[N,K]^T + bias[N]     (A, B K-contiguous; K % 16 == 0; lda/ldb % 4 == 0)
// 128x128x16 tiles, 256 threads, 8x8 register micro-tiles, double-buffered shared memory.
// blockIdx.z batches independent problems through the *_bs strides (used for the two directions).
// ------------------------------------------------------------------------------------------------
constexpr int BM = 128, BN = 128, BK = 16, PADM = 4;

__global__ __launch_bounds__(256) void sgemm_nt_kernel(int M, int N, int K, const float* __restrict__ A, int lda,
                                                       long A_bs, const float* __restrict__ B, int ldb, long B_bs,
                                                       const float* __restrict__ bias, long bias_bs,
                                                       float* __restrict__ C, int ldc, long C_bs) {
  __shared__ __align__(16) float As[2][BK][BM + PADM];
  __shared__ __align__(16) float Bs[2][BK][BN + PADM];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  A += (long)blockIdx.z * A_bs;
  B += (long)blockIdx.z * B_bs;
  C += (long)blockIdx.z * C_bs;
  if (bias) bias += (long)blockIdx.z * bias_bs;

  // global->smem assignment: 512 float4 per tile, two per thread
  int lrow[2], lkq[2];
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    int idx = tid + i * 256;
    lrow[i] = idx >> 2;
    lkq[i] = idx & 3;
  }
  float4 ra[2], rb[2];
  auto gload = [&](int k0) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      int gm = m0 + lrow[i], gn = n0 + lrow[i];
      ra[i] = gm < M ? *reinterpret_cast<const float4*>(A + (long)gm * lda + k0 + lkq[i] * 4) : make_float4(0, 0, 0, 0);
      rb[i] = gn < N ? *reinterpret_cast<const float4*>(B + (long)gn * ldb + k0 + lkq[i] * 4) : make_float4(0, 0, 0, 0);
    }
  };
  auto sstore = [&](int buf) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      int k = lkq[i] * 4;
      As[buf][k + 0][lrow[i]] = ra[i].x; As[buf][k + 1][lrow[i]] = ra[i].y;
      As[buf][k + 2][lrow[i]] = ra[i].z; As[buf][k + 3][lrow[i]] = ra[i].w;
      Bs[buf][k + 0][lrow[i]] = rb[i].x; Bs[buf][k + 1][lrow[i]] = rb[i].y;
      Bs[buf][k + 2][lrow[i]] = rb[i].z; Bs[buf][k + 3][lrow[i]] = rb[i].w;
    }
  };

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  gload(0);
  sstore(0);
  __syncthreads();
  const int nk = K / BK;
  for (int kt = 0; kt < nk; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < nk) gload((kt + 1) * BK);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
      float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][64 + ty * 4]);
      float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
      float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][k][64 + tx * 4]);
      float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (kt + 1 < nk) {
      sstore(buf ^ 1);
      __syncthreads();
    }
  }

#pragma unroll
  for (int i = 0; i < 8; ++i) {
    int gm = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (gm >= M) continue;
#pragma unroll
    for (int jh = 0; jh < 2; ++jh) {
      int gn = n0 + jh * 64 + tx * 4;
      float v[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] = acc[i][jh * 4 + j] + ((bias && gn + j < N) ? bias[gn + j] : 0.f);
      float* cp = C + (long)gm * ldc + gn;
      if (gn + 3 < N && ((ldc & 3) == 0)) {
        *reinterpret_cast<float4*>(cp) = make_float4(v[0], v[1], v[2], v[3]);
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (gn + j < N) cp[j] = v[j];
      }
    }
  }
}

static int sgemm_nt(int M, int N, int K, const float* A, int lda, long A_bs, const float* B, int ldb, long B_bs,
                    const float* bias, long bias_bs, float* C, int ldc, long C_bs, int batch, cudaStream_t st) {
  if (K % BK != 0 || (lda & 3) || (ldb & 3)) {
    set_error("sgemm_nt: K=%d lda=%d ldb=%d violate alignment", K, lda, ldb);
    return CCSM_EINVAL;
  }
  dim3 grid((N + BN - 1) / BN, (M + BM - 1) / BM, batch);
  sgemm_nt_kernel<<<grid, 256, 0, st>>>(M, N, K, A, lda, A_bs, B, ldb, B_bs, bias, bias_bs, C, ldc, C_bs);
  count_launch();
  CCSM_CUDA(cudaGetLastError());
  return CCSM_OK;
}

// ------------------------------------------------------------------------------------------------
// Feature packing: two-strand embedding lookup + kinetics concat (reference models.py:91-123).
// x0[R][t][0..Kpad): [embed(kmer)(E) | ipd | pw | npass? | ipd_std, pw_std? | sn(4)? | map?] then zeros.
// ------------------------------------------------------------------------------------------------
struct StrandPtrs {
  const float *kmer, *kpass, *ipd, *ipd_sd, *pw, *pw_sd, *sns, *maps;
};

__global__ void pack_x_att2s_kernel(int64_t sites, int L, int E, int n_vocab, int flags, int Kpad, StrandPtrs s0,
                                    StrandPtrs s1, const float* __restrict__ embed, float* __restrict__ x0) {
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // over (site, strand, t)
  int64_t total = sites * 2 * L;
  if (idx >= total) return;
  int t = (int)(idx % L);
  int64_t R = idx / L;
  int strand = (int)(R & 1);
  int64_t site = R >> 1;
  const StrandPtrs& s = strand ? s1 : s0;
  int64_t o = site * L + t;
  float* x = x0 + idx * Kpad;
  int code = (int)s.kmer[o];  // float -> int truncation == tensor.int() (models.py:91)
  code = code < 0 ? 0 : (code >= n_vocab ? n_vocab - 1 : code);
  int k = 0;
  for (; k < E; ++k) x[k] = embed[code * E + k];
  x[k++] = s.ipd[o];
  x[k++] = s.pw[o];
  if (flags & CCSM_FEAT_NPASS) x[k++] = s.kpass[o];
  if (flags & CCSM_FEAT_STDS) {
    x[k++] = s.ipd_sd[o];
    x[k++] = s.pw_sd[o];
  }
  if (flags & CCSM_FEAT_SN) {
    for (int j = 0; j < 4; ++j) x[k++] = s.sns[site * 4 + j];
  }
  if (flags & CCSM_FEAT_MAP) x[k++] = s.maps[o];
  for (; k < Kpad; ++k) x[k] = 0.f;
}

// ModelAttRNN2 (models.py:319-333): x = [seq_embed[kmer] (E) | ipd_embed[int(ipd)] (8) | pw_embed[int(pw)] (8) |
// npass_embed[clamp(npass, 1, 30)] (4)?].  Indices outside a table (the reference would raise) are clamped.
__global__ void pack_x_att2s2_kernel(int64_t sites, int L, int E, int n_vocab, int flags, int Kpad, StrandPtrs s0,
                                     StrandPtrs s1, const float* __restrict__ seq_embed,
                                     const float* __restrict__ ipd_embed, const float* __restrict__ pw_embed,
                                     const float* __restrict__ npass_embed, float* __restrict__ x0) {
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // over (site, strand, t)
  if (idx >= sites * 2 * L) return;
  int t = (int)(idx % L);
  int64_t R = idx / L;
  const StrandPtrs& s = (R & 1) ? s1 : s0;
  int64_t o = (R >> 1) * L + t;
  float* x = x0 + idx * Kpad;
  auto clampi = [](int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); };
  const int code = clampi((int)s.kmer[o], 0, n_vocab - 1);
  const int ic = clampi((int)s.ipd[o], 0, 952), pc = clampi((int)s.pw[o], 0, 952);
  int k = 0;
  for (int j = 0; j < E; ++j) x[k++] = seq_embed[code * E + j];
  for (int j = 0; j < 8; ++j) x[k++] = ipd_embed[ic * 8 + j];
  for (int j = 0; j < 8; ++j) x[k++] = pw_embed[pc * 8 + j];
  if (flags & CCSM_FEAT_NPASS) {
    // torch.clamp(kpass, 1, MAX_PASSES).int(): clamp the float, then truncate
    const float kp = fminf(fmaxf(s.kpass[o], 1.f), 30.f);
    const int np = (int)kp;
    for (int j = 0; j < 4; ++j) x[k++] = npass_embed[np * 4 + j];
  }
  for (; k < Kpad; ++k) x[k] = 0.f;
}

// classifier tail of ModelAttRNN2 (models.py:275-278,378-380): hid already holds Linear(4H,4H)(ctx) + bias;
// logits = Linear(4H, classes)(relu(hid)), probs = softmax(logits).  One warp per site.
template <int MAXC>
__global__ void cls_out_kernel(int64_t sites, int D, int classes, const float* __restrict__ hid,
                               const float* __restrict__ w, const float* __restrict__ b, float* __restrict__ logits,
                               float* __restrict__ probs) {
  const int lane = threadIdx.x & 31;
  const int64_t site = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (site >= sites) return;
  float lg[MAXC];
#pragma unroll
  for (int k = 0; k < MAXC; ++k) lg[k] = 0.f;
  for (int j = lane; j < D; j += 32) {
    const float v = fmaxf(hid[site * D + j], 0.f);
#pragma unroll
    for (int k = 0; k < MAXC; ++k)
      if (k < classes) lg[k] += v * w[(int64_t)k * D + j];
  }
#pragma unroll
  for (int k = 0; k < MAXC; ++k) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) lg[k] += __shfl_xor_sync(0xffffffffu, lg[k], o);
    if (k < classes) lg[k] += b[k];
  }
  if (lane == 0) {
    float mx = -INFINITY, sum = 0.f;
    for (int k = 0; k < classes; ++k) mx = fmaxf(mx, lg[k]);
    for (int k = 0; k < classes; ++k) sum += expf(lg[k] - mx);
    for (int k = 0; k < classes; ++k) {
      if (logits) logits[site * classes + k] = lg[k];
      if (probs) probs[site * classes + k] = expf(lg[k] - mx) / sum;
    }
  }
}

// aggregate model: x = cat(histos (n,L,B), offsets (n,L,1))   (reference models.py:675-677)
__global__ void pack_x_aggr_kernel(int64_t sites, int L, int Bn, int Kpad, const float* __restrict__ offsets,
                                   const float* __restrict__ histos, float* __restrict__ x0) {
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // over (site, t)
  if (idx >= sites * L) return;
  float* x = x0 + idx * Kpad;
  const float* h = histos + idx * Bn;
  int k = 0;
  for (; k < Bn; ++k) x[k] = h[k];
  x[k++] = offsets[idx];
  for (; k < Kpad; ++k) x[k] = 0.f;
}

// h[R][d][u] = h0_strand(R)[(2*layer + d)][site][u]   (h0 index 2*layer+direction, torch nn.GRU)
__global__ void load_h0_kernel(int64_t rows, int strands, int H, int layer, int NL, int64_t n_total, int64_t site0,
                               const float* __restrict__ h0_a, const float* __restrict__ h0_b, int h0_random,
                               unsigned long long h0_seed, unsigned long long h0_offset, float* __restrict__ h) {
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // over (R, d, u)
  if (idx >= rows * 2 * H) return;
  int u = (int)(idx % H);
  int d = (int)((idx / H) & 1);
  int64_t R = idx / (2 * H);
  int strand = (int)(R % strands);
  int64_t site = site0 + R / strands;
  const float* h0 = strand ? h0_b : h0_a;
  if (!h0 && h0_random) {
    // same stream as tc_prep_kernel: output u of subsequence (row*2*layers + 2*layer+dir), see include/ccsm.h
    curandStatePhilox4_32_10_t rng;
    curand_init(h0_seed, (unsigned long long)((site * strands + strand) * 2 * NL + 2 * layer + d),
                h0_offset + (unsigned long long)(u & ~3), &rng);
    const float4 v = curand_normal4(&rng);
    h[idx] = (u & 3) == 0 ? v.x : ((u & 3) == 1 ? v.y : ((u & 3) == 2 ? v.z : v.w));
    return;
  }
  h[idx] = h0 ? h0[((int64_t)(2 * layer + d) * n_total + site) * H + u] : 0.f;
}

// One GRU time step for both directions (PyTorch cell; gi/gh already contain b_ih / b_hh):
//   r = sig(gi_r + gh_r), z = sig(gi_z + gh_z), n = tanh(gi_n + r * gh_n), h' = (1 - z) * n + z * h
__global__ void gru_step_kernel(int64_t rows, int L, int H, int step, const float* __restrict__ gi,
                                const float* __restrict__ gh, float* __restrict__ h, float* __restrict__ out) {
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // over (R, d, u)
  if (idx >= rows * 2 * H) return;
  int u = (int)(idx % H);
  int d = (int)((idx / H) & 1);
  int64_t R = idx / (2 * H);
  int t = d ? (L - 1 - step) : step;
  const float* gip = gi + ((R * L + t) * 2 + d) * 3 * (int64_t)H;
  const float* ghp = gh + (R * 2 + d) * 3 * (int64_t)H;
  float r = 1.f / (1.f + expf(-(gip[u] + ghp[u])));
  float z = 1.f / (1.f + expf(-(gip[H + u] + ghp[H + u])));
  float nn = tanhf(gip[2 * H + u] + r * ghp[2 * H + u]);
  float hp = h[idx];
  float hn = (1.f - z) * nn + z * hp;
  h[idx] = hn;
  out[(R * L + t) * 2 * (int64_t)H + d * H + u] = hn;
}

// One LSTM time step for both directions (PyTorch cell, gate row blocks i, f, g, o; gi/gh contain the biases):
//   i = sig(.), f = sig(.), g = tanh(.), o = sig(.);  c' = f * c + i * g;  h' = o * tanh(c')
// reference models.py:48-51 (nn.LSTM in ModelAttRNN(model_type="attbilstm2s")).
__global__ void lstm_step_kernel(int64_t rows, int L, int H, int step, const float* __restrict__ gi,
                                 const float* __restrict__ gh, float* __restrict__ h, float* __restrict__ c,
                                 float* __restrict__ out) {
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // over (R, d, u)
  if (idx >= rows * 2 * H) return;
  int u = (int)(idx % H);
  int d = (int)((idx / H) & 1);
  int64_t R = idx / (2 * H);
  int t = d ? (L - 1 - step) : step;
  const float* gip = gi + ((R * L + t) * 2 + d) * 4 * (int64_t)H;
  const float* ghp = gh + (R * 2 + d) * 4 * (int64_t)H;
  const float ig = 1.f / (1.f + expf(-(gip[u] + ghp[u])));
  const float fg = 1.f / (1.f + expf(-(gip[H + u] + ghp[H + u])));
  const float gg = tanhf(gip[2 * H + u] + ghp[2 * H + u]);
  const float og = 1.f / (1.f + expf(-(gip[3 * H + u] + ghp[3 * H + u])));
  const float cn = fg * c[idx] + ig * gg;
  const float hn = og * tanhf(cn);
  c[idx] = cn;
  h[idx] = hn;
  out[(R * L + t) * 2 * (int64_t)H + d * H + u] = hn;
}

// Attention reduction + head, one warp per site (both strands).
//   e_t = va . tanh(qa + E_t);  w = softmax_t(e);  ctx = sum_t w_t out_t      (attention.py:55-70)
//   logits = fc1 [ctx_strand1 | ctx_strand2] + b;  probs = softmax(logits)    (models.py:145-150)
// The aggregate model has one strand and returns the raw fc1 output (models.py:690-694).
template <int MAXC>
__global__ void att_head_kernel(int64_t sites, int strands, int L, int H, int classes, int do_softmax,
                                const float* __restrict__ E, const float* __restrict__ qa,
                                const float* __restrict__ out, const float* __restrict__ va,
                                const float* __restrict__ fc_w, const float* __restrict__ fc_b,
                                float* __restrict__ logits, float* __restrict__ probs, float* __restrict__ ctx_out) {
  // ctx_out != nullptr: only write the context vectors [ctx_strand1 | ctx_strand2] (ModelAttRNN2's classifier follows)
  const int lane = threadIdx.x & 31;
  const int64_t site = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (site >= sites) return;
  const int C2 = 2 * H;
  float lg[MAXC];
#pragma unroll
  for (int k = 0; k < MAXC; ++k) lg[k] = 0.f;
  for (int s = 0; s < strands; ++s) {
    const int64_t R = site * strands + s;
    float my_e = -INFINITY;  // lane t keeps e_t
    for (int t = 0; t < L; ++t) {
      const float* e = E + (R * L + t) * (int64_t)H;
      float p = 0.f;
      for (int j = lane; j < H; j += 32) p += va[j] * tanhf(qa[R * H + j] + e[j]);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) p += __shfl_xor_sync(0xffffffffu, p, o);
      if (lane == t) my_e = p;
    }
    float mx = my_e;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float w = lane < L ? expf(my_e - mx) : 0.f;
    float sum = w;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    w /= sum;
    for (int c = lane; c < C2; c += 32) {
      float ctx = 0.f;
      for (int t = 0; t < L; ++t) ctx += __shfl_sync(0xffffffffu, w, t) * out[(R * L + t) * (int64_t)C2 + c];
      if (ctx_out) {
        ctx_out[site * (int64_t)strands * C2 + s * C2 + c] = ctx;
        continue;
      }
#pragma unroll
      for (int k = 0; k < MAXC; ++k)
        if (k < classes) lg[k] += ctx * fc_w[(int64_t)k * strands * C2 + s * C2 + c];
    }
  }
  if (ctx_out) return;
#pragma unroll
  for (int k = 0; k < MAXC; ++k) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) lg[k] += __shfl_xor_sync(0xffffffffu, lg[k], o);
    if (k < classes) lg[k] += fc_b[k];
  }
  if (lane == 0) {
    float mx = -INFINITY, sum = 0.f;
    for (int k = 0; k < classes; ++k) mx = fmaxf(mx, lg[k]);
    for (int k = 0; k < classes; ++k) sum += expf(lg[k] - mx);
    for (int k = 0; k < classes; ++k) {
      if (logits) logits[site * classes + k] = lg[k];
      if (probs) probs[site * classes + k] = do_softmax ? expf(lg[k] - mx) / sum : lg[k];
    }
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
static const HostTensor* find(ccsm_model* m, const std::string& k) {
  auto it = m->w.find(k);
  return it == m->w.end() ? nullptr : &it->second;
}

static int upload(DevBuf& b, const std::vector<float>& v) {
  CCSM_TRY(b.reserve(v.size() * sizeof(float)));
  CCSM_CUDA(cudaMemcpy(b.p, v.data(), v.size() * sizeof(float), cudaMemcpyHostToDevice));
  return CCSM_OK;
}

int fp32_upload_weights(ccsm_model* m) {
  const int H = m->cfg.hidden, NL = m->cfg.num_layers;
  Fp32Weights& W = m->fp32;
  W.layers.resize(NL);
  static const char* sfx[2] = {"", "_reverse"};
  for (int l = 0; l < NL; ++l) {
    Fp32Layer& Lw = W.layers[l];
    Lw.K = l == 0 ? m->in_feat : 2 * H;
    Lw.Kpad = round_up(Lw.K, 16);
    const int G = m->gates;
    std::vector<float> wih((size_t)2 * G * H * Lw.Kpad, 0.f), bih((size_t)2 * G * H), whh((size_t)2 * G * H * H),
        bhh((size_t)2 * G * H);
    for (int d = 0; d < 2; ++d) {
      std::string base = "rnn.";
      const HostTensor* a = find(m, base + "weight_ih_l" + std::to_string(l) + sfx[d]);
      const HostTensor* b = find(m, base + "weight_hh_l" + std::to_string(l) + sfx[d]);
      const HostTensor* c = find(m, base + "bias_ih_l" + std::to_string(l) + sfx[d]);
      const HostTensor* e = find(m, base + "bias_hh_l" + std::to_string(l) + sfx[d]);
      if (!a || !b || !c || !e) {
        set_error("finalize: missing GRU tensors for layer %d%s", l, sfx[d]);
        return CCSM_EKEY;
      }
      for (int r = 0; r < G * H; ++r)
        for (int k = 0; k < Lw.K; ++k) wih[((size_t)d * G * H + r) * Lw.Kpad + k] = a->data[(size_t)r * Lw.K + k];
      std::copy(b->data.begin(), b->data.end(), whh.begin() + (size_t)d * G * H * H);
      std::copy(c->data.begin(), c->data.end(), bih.begin() + (size_t)d * G * H);
      std::copy(e->data.begin(), e->data.end(), bhh.begin() + (size_t)d * G * H);
    }
    CCSM_TRY(upload(Lw.w_ih, wih));
    CCSM_TRY(upload(Lw.b_ih, bih));
    CCSM_TRY(upload(Lw.w_hh, whh));
    CCSM_TRY(upload(Lw.b_hh, bhh));
  }
  CCSM_TRY(upload(W.Wa, find(m, "_att3.Wa.weight")->data));
  CCSM_TRY(upload(W.Ua, find(m, "_att3.Ua.weight")->data));
  CCSM_TRY(upload(W.va, find(m, "_att3.va.weight")->data));
  if (m->is_2s2) {
    CCSM_TRY(upload(W.embed, find(m, "seq_embed.weight")->data));
    CCSM_TRY(upload(W.ipd_embed, find(m, "ipd_embed.weight")->data));
    CCSM_TRY(upload(W.pw_embed, find(m, "pw_embed.weight")->data));
    if (m->cfg.feat_flags & CCSM_FEAT_NPASS) CCSM_TRY(upload(W.npass_embed, find(m, "npass_embed.weight")->data));
    CCSM_TRY(upload(W.cls0_w, find(m, "classifier.0.weight")->data));
    CCSM_TRY(upload(W.cls0_b, find(m, "classifier.0.bias")->data));
    CCSM_TRY(upload(W.fc_w, find(m, "classifier.3.weight")->data));
    CCSM_TRY(upload(W.fc_b, find(m, "classifier.3.bias")->data));
  } else {
    if (m->cfg.kind == CCSM_KIND_ATT2S) CCSM_TRY(upload(W.embed, find(m, "embed.weight")->data));
    CCSM_TRY(upload(W.fc_w, find(m, "fc1.weight")->data));
    CCSM_TRY(upload(W.fc_b, find(m, "fc1.bias")->data));
  }
  W.ready = true;
  return CCSM_OK;
}

static int reserve_ws(ccsm_model* m, int64_t rows) {
  Fp32Workspace& ws = m->ws32;
  if (rows <= ws.rows_cap) return CCSM_OK;
  const int64_t H = m->cfg.hidden, L = m->cfg.seq_len;
  const int64_t K0 = m->fp32.layers[0].Kpad;
  CCSM_TRY(ws.x0.reserve(rows * L * K0 * 4));
  const int64_t G = m->gates;
  CCSM_TRY(ws.gi.reserve(rows * L * 2 * G * H * 4));
  CCSM_TRY(ws.gh.reserve(rows * 2 * G * H * 4));
  CCSM_TRY(ws.h.reserve(rows * 2 * H * 4));
  if (G == 4) CCSM_TRY(ws.c.reserve(rows * 2 * H * 4));
  CCSM_TRY(ws.outA.reserve(rows * L * 2 * H * 4));
  CCSM_TRY(ws.outB.reserve(rows * L * 2 * H * 4));
  CCSM_TRY(ws.qa.reserve(rows * H * 4));
  ws.rows_cap = rows;
  return CCSM_OK;
}

static inline unsigned nblk(int64_t total, int threads) { return (unsigned)((total + threads - 1) / threads); }

// Runs layers + attention + head on x0 (already packed) for `sites` sites starting at site0.
static int run_stack(ccsm_model* m, int64_t sites, int64_t site0, int64_t n_total, const float* h0_a,
                     const float* h0_b, float* logits, float* probs, cudaStream_t st, const float* c0_a = nullptr,
                     const float* c0_b = nullptr) {
  const int H = m->cfg.hidden, L = m->cfg.seq_len, NL = m->cfg.num_layers, S = m->strands, G = m->gates;
  const int64_t rows = sites * S;
  Fp32Workspace& ws = m->ws32;
  Fp32Weights& W = m->fp32;
  const float* xin = ws.x0.as<float>();
  float* outs[2] = {ws.outA.as<float>(), ws.outB.as<float>()};
  float* out = nullptr;
  for (int l = 0; l < NL; ++l) {
    Fp32Layer& Lw = W.layers[l];
    out = outs[l & 1];
    // input projection for all time steps and both directions: gi[R][t][d][3H]
    CCSM_TRY(sgemm_nt((int)(rows * L), 2 * G * H, Lw.Kpad, xin, Lw.Kpad, 0, Lw.w_ih.as<float>(), Lw.Kpad, 0,
                      Lw.b_ih.as<float>(), 0, ws.gi.as<float>(), 2 * G * H, 0, 1, st));
    load_h0_kernel<<<nblk(rows * 2 * H, 256), 256, 0, st>>>(
        rows, S, H, l, NL, n_total, site0, h0_a, h0_b, m->h0_mode == CCSM_H0_DEVICE_RANDOM ? 1 : 0,
        (unsigned long long)m->h0_seed, (unsigned long long)(m->h0_calls * 256), ws.h.as<float>());
    count_launch();
    if (G == 4) {  // c0: explicit, zeros, or (device-random mode) a second stream with its own seed
      load_h0_kernel<<<nblk(rows * 2 * H, 256), 256, 0, st>>>(
          rows, S, H, l, NL, n_total, site0, c0_a, c0_b, m->h0_mode == CCSM_H0_DEVICE_RANDOM ? 1 : 0,
          (unsigned long long)(m->h0_seed ^ 0x9e3779b97f4a7c15ULL), (unsigned long long)(m->h0_calls * 256), ws.c.as<float>());
      count_launch();
    }
    for (int s = 0; s < L; ++s) {
      // gh[R][d][G*H] = h[R][d][:] . W_hh[d]^T + b_hh[d]
      CCSM_TRY(sgemm_nt((int)rows, G * H, H, ws.h.as<float>(), 2 * H, H, Lw.w_hh.as<float>(), H, (long)G * H * H,
                        Lw.b_hh.as<float>(), G * H, ws.gh.as<float>(), 2 * G * H, G * H, 2, st));
      if (G == 4)
        lstm_step_kernel<<<nblk(rows * 2 * H, 256), 256, 0, st>>>(rows, L, H, s, ws.gi.as<float>(), ws.gh.as<float>(),
                                                                   ws.h.as<float>(), ws.c.as<float>(), out);
      else
        gru_step_kernel<<<nblk(rows * 2 * H, 256), 256, 0, st>>>(rows, L, H, s, ws.gi.as<float>(), ws.gh.as<float>(),
                                                                  ws.h.as<float>(), out);
      count_launch();
    }
    xin = out;
  }
  CCSM_CUDA(cudaGetLastError());
  m->dbg_rnn_out = out;
  m->dbg_rnn_out_floats = rows * L * 2 * H;
  // attention: E = out . Ua^T (reuses gi), qa = q . Wa^T with q = h[R] = [h_n fwd | h_n rev] of the last layer
  float* E = ws.gi.as<float>();
  CCSM_TRY(sgemm_nt((int)(rows * L), H, 2 * H, out, 2 * H, 0, W.Ua.as<float>(), 2 * H, 0, nullptr, 0, E, H, 0, 1, st));
  CCSM_TRY(sgemm_nt((int)rows, H, 2 * H, ws.h.as<float>(), 2 * H, 0, W.Wa.as<float>(), 2 * H, 0, nullptr, 0,
                    ws.qa.as<float>(), H, 0, 1, st));
  const int warps = 4;
  if (m->cfg.num_classes > 4) {
    set_error("num_classes > 4 unsupported");
    return CCSM_EINVAL;
  }
  float* lg = logits ? logits + site0 * m->cfg.num_classes : nullptr;
  float* pr = probs ? probs + site0 * m->cfg.num_classes : nullptr;
  if (m->is_2s2) {
    // contexts -> classifier.0 (GEMM + bias) -> ReLU + classifier.3 + softmax; scratch lives in gi behind E
    const int D = S * 2 * H;
    float* ctx = E + rows * L * (int64_t)H;
    float* hid = ctx + sites * (int64_t)D;
    att_head_kernel<4><<<nblk(sites, warps), warps * 32, 0, st>>>(sites, S, L, H, 0, 0, E, ws.qa.as<float>(), out,
                                                                  W.va.as<float>(), nullptr, nullptr, nullptr, nullptr, ctx);
    count_launch();
    CCSM_TRY(sgemm_nt((int)sites, D, D, ctx, D, 0, W.cls0_w.as<float>(), D, 0, W.cls0_b.as<float>(), 0, hid, D, 0, 1, st));
    cls_out_kernel<4><<<nblk(sites, warps), warps * 32, 0, st>>>(sites, D, m->cfg.num_classes, hid, W.fc_w.as<float>(),
                                                                 W.fc_b.as<float>(), lg, pr);
    count_launch();
    CCSM_CUDA(cudaGetLastError());
    return CCSM_OK;
  }
  att_head_kernel<4><<<nblk(sites, warps), warps * 32, 0, st>>>(
      sites, S, L, H, m->cfg.num_classes, m->cfg.kind == CCSM_KIND_ATT2S ? 1 : 0, E, ws.qa.as<float>(), out,
      W.va.as<float>(), W.fc_w.as<float>(), W.fc_b.as<float>(), lg, pr, nullptr);
  count_launch();
  CCSM_CUDA(cudaGetLastError());
  return CCSM_OK;
}

static const int64_t kChunkSites = 8192;

static StrandPtrs offset_strand(const ccsm_strand* s, int64_t site0, int L) {
  StrandPtrs p;
  auto off = [&](const float* q, int64_t per) { return q ? q + site0 * per : nullptr; };
  p.kmer = off(s->kmer, L);
  p.kpass = off(s->kpass, L);
  p.ipd = off(s->ipd_means, L);
  p.ipd_sd = off(s->ipd_stds, L);
  p.pw = off(s->pw_means, L);
  p.pw_sd = off(s->pw_stds, L);
  p.sns = off(s->sns, 4);
  p.maps = off(s->maps, L);
  return p;
}

int fp32_forward_att2s(ccsm_model* m, int64_t n, const ccsm_strand* fwd, const ccsm_strand* rev, const float* h0_f,
                       const float* h0_r, float* logits, float* probs, cudaStream_t st, const float* c0_f,
                       const float* c0_r) {
  const int L = m->cfg.seq_len;
  const int64_t chunk = n < kChunkSites ? n : kChunkSites;
  CCSM_TRY(reserve_ws(m, chunk * 2));
  for (int64_t s0 = 0; s0 < n; s0 += chunk) {
    int64_t sites = (n - s0) < chunk ? (n - s0) : chunk;
    if (m->is_2s2)
      pack_x_att2s2_kernel<<<nblk(sites * 2 * L, 256), 256, 0, st>>>(
          sites, L, m->cfg.n_embed, m->cfg.n_vocab, m->cfg.feat_flags, m->fp32.layers[0].Kpad, offset_strand(fwd, s0, L),
          offset_strand(rev, s0, L), m->fp32.embed.as<float>(), m->fp32.ipd_embed.as<float>(), m->fp32.pw_embed.as<float>(),
          m->fp32.npass_embed.as<float>(), m->ws32.x0.as<float>());
    else
      pack_x_att2s_kernel<<<nblk(sites * 2 * L, 256), 256, 0, st>>>(
          sites, L, m->cfg.n_embed, m->cfg.n_vocab, m->cfg.feat_flags, m->fp32.layers[0].Kpad,
          offset_strand(fwd, s0, L), offset_strand(rev, s0, L), m->fp32.embed.as<float>(), m->ws32.x0.as<float>());
    count_launch();
    CCSM_TRY(run_stack(m, sites, s0, n, h0_f, h0_r, logits, probs, st, c0_f, c0_r));
  }
  return CCSM_OK;
}

int fp32_forward_aggr(ccsm_model* m, int64_t n, const float* offsets, const float* histos, const float* h0,
                      float* out, cudaStream_t st, const float* c0) {
  const int L = m->cfg.seq_len, Bn = m->cfg.feat_flags;
  const int64_t chunk = n < 65536 ? n : 65536;
  CCSM_TRY(reserve_ws(m, chunk));
  for (int64_t s0 = 0; s0 < n; s0 += chunk) {
    int64_t sites = (n - s0) < chunk ? (n - s0) : chunk;
    pack_x_aggr_kernel<<<nblk(sites * L, 256), 256, 0, st>>>(sites, L, Bn, m->fp32.layers[0].Kpad,
                                                             offsets + s0 * L, histos + s0 * L * Bn,
                                                             m->ws32.x0.as<float>());
    count_launch();
    CCSM_TRY(run_stack(m, sites, s0, n, h0, nullptr, nullptr, out, st, c0, nullptr));
  }
  return CCSM_OK;
}


// ================================================================================================
// ModelTransEnc ("transencoder2s", reference models.py:451-620) on fp32 FFMA kernels.  No checkpoint ships for this
// model type; it exists for `--model_type` completeness and reuses sgemm_nt for every
// linear map.  Token order: row R = site * 2 + strand, token = R * L + t.
//   embeddings (pack_x_att2s2_kernel) -> 3 x [Conv1d(k=3, p=1) as im2col + GEMM -> BatchNorm(eval) -> ReLU ->
//   MaxPool1d(k=3, s=1, p=1)] -> + pos_embed -> num_layers x [QKV GEMM -> per-(row, head) softmax attention ->
//   out-proj GEMM -> add + LayerNorm -> FFN GEMM + ReLU + GEMM -> add + LayerNorm] -> mean over t -> classifier.
// ================================================================================================
__global__ void tr_im2col_kernel(int64_t tokens, int L, int cin, int ldin, int kpad, const float* __restrict__ in,
                                 float* __restrict__ col) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // over (token, k)
  if (idx >= tokens * kpad) return;
  const int k = (int)(idx % kpad);
  const int64_t tok = idx / kpad;
  float v = 0.f;
  if (k < 3 * cin) {
    const int dk = k / cin, ci = k - dk * cin;
    const int t = (int)(tok % L) + dk - 1;
    if (t >= 0 && t < L) v = in[(tok + dk - 1) * ldin + ci];
  }
  col[idx] = v;
}

// out[R, t, c] = max over the valid neighbours t-1, t, t+1 of relu(v * scale[c] + shift[c])  (+ pos[t][c])
__global__ void tr_bn_relu_pool_kernel(int64_t tokens, int L, int C, const float* __restrict__ v,
                                       const float* __restrict__ scale, const float* __restrict__ shift,
                                       const float* __restrict__ pos, float* __restrict__ out) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // over (token, c)
  if (idx >= tokens * C) return;
  const int c = (int)(idx % C);
  const int64_t tok = idx / C;
  const int t = (int)(tok % L);
  const float sc = scale[c], sh = shift[c];
  float m = fmaxf(fmaf(v[idx], sc, sh), 0.f);
  if (t > 0) m = fmaxf(m, fmaxf(fmaf(v[idx - C], sc, sh), 0.f));
  if (t + 1 < L) m = fmaxf(m, fmaxf(fmaf(v[idx + C], sc, sh), 0.f));
  out[idx] = pos ? m + pos[t * C + c] : m;
}

// softmax(q k^T / sqrt(dh)) v for one (row, head) per warp; lane = query position (L <= 32)
__global__ void tr_attention_kernel(int64_t rows, int L, int d, int nhead, const float* __restrict__ qkv,
                                    float* __restrict__ out) {
  const int64_t w = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (w >= rows * nhead) return;
  const int head = (int)(w % nhead);
  const int64_t R = w / nhead;
  const int dh = d / nhead;
  const float* base = qkv + R * L * (int64_t)(3 * d) + head * dh;
  if (lane >= L) return;
  const float* q = base + (int64_t)lane * 3 * d;
  const float inv = rsqrtf((float)dh);
  float sc[32];
  float mx = -INFINITY;
  for (int j = 0; j < L; ++j) {
    const float* k = base + (int64_t)j * 3 * d + d;
    float a = 0.f;
    for (int c = 0; c < dh; ++c) a = fmaf(q[c], k[c], a);
    sc[j] = a * inv;
    mx = fmaxf(mx, sc[j]);
  }
  float sum = 0.f;
  for (int j = 0; j < L; ++j) {
    sc[j] = expf(sc[j] - mx);
    sum += sc[j];
  }
  float* o = out + (R * L + lane) * (int64_t)d + head * dh;
  for (int c = 0; c < dh; ++c) {
    float a = 0.f;
    for (int j = 0; j < L; ++j) a = fmaf(sc[j], base[(int64_t)j * 3 * d + 2 * d + c], a);
    o[c] = a / sum;
  }
}

// x = LayerNorm(x + y) * w + b, one warp per token (biased variance, eps 1e-5)
__global__ void tr_add_ln_kernel(int64_t tokens, int d, float* __restrict__ x, const float* __restrict__ y,
                                 const float* __restrict__ w, const float* __restrict__ b) {
  const int64_t tok = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (tok >= tokens) return;
  float* xp = x + tok * d;
  const float* yp = y + tok * d;
  float s = 0.f;
  for (int c = lane; c < d; c += 32) s += xp[c] + yp[c];
  for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mu = s / d;
  float q = 0.f;
  for (int c = lane; c < d; c += 32) {
    const float t = xp[c] + yp[c] - mu;
    q += t * t;
  }
  for (int o = 16; o; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
  const float rstd = 1.f / sqrtf(q / d + 1e-5f);
  for (int c = lane; c < d; c += 32) xp[c] = (xp[c] + yp[c] - mu) * rstd * w[c] + b[c];
}

__global__ void tr_relu_kernel(int64_t n, float* __restrict__ x) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) x[i] = fmaxf(x[i], 0.f);
}

// ctx[site][strand * d + c] = mean_t x[(site * 2 + strand) * L + t][c]
__global__ void tr_mean_pool_kernel(int64_t sites, int L, int d, const float* __restrict__ x, float* __restrict__ ctx) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // over (site, strand, c)
  if (idx >= sites * 2 * d) return;
  const int c = (int)(idx % d);
  const int64_t R = idx / d;
  float s = 0.f;
  for (int t = 0; t < L; ++t) s += x[(R * L + t) * d + c];
  ctx[idx] = s / L;
}

int trans_upload_weights(ccsm_model* m) {
  const int d = m->cfg.hidden;
  Fp32Weights& W = m->fp32;
  TrWeights& T = m->tr;
  if (m->dim_ff <= 0) {
    set_error("finalize: feed-forward weights missing");
    return CCSM_EKEY;
  }
  CCSM_TRY(upload(W.embed, find(m, "seq_embed.weight")->data));
  CCSM_TRY(upload(W.ipd_embed, find(m, "ipd_embed.weight")->data));
  CCSM_TRY(upload(W.pw_embed, find(m, "pw_embed.weight")->data));
  if (m->cfg.feat_flags & CCSM_FEAT_NPASS) CCSM_TRY(upload(W.npass_embed, find(m, "npass_embed.weight")->data));
  CCSM_TRY(upload(W.cls0_w, find(m, "classifier.0.weight")->data));
  CCSM_TRY(upload(W.cls0_b, find(m, "classifier.0.bias")->data));
  CCSM_TRY(upload(W.fc_w, find(m, "classifier.3.weight")->data));
  CCSM_TRY(upload(W.fc_b, find(m, "classifier.3.bias")->data));
  CCSM_TRY(upload(T.pos, find(m, "pos_encoder.pos_embed.weight")->data));
  static const char* convs[3] = {"trans_input.conv_embed.0", "trans_input.conv_embed.4", "trans_input.conv_embed_plus.0.conv_embed.0"};
  static const char* bns[3] = {"trans_input.conv_embed.1", "trans_input.conv_embed.5", "trans_input.conv_embed_plus.0.conv_embed.1"};
  const int cin[3] = {m->in_feat, d / 2, d}, cout[3] = {d / 2, d, d};
  for (int i = 0; i < 3; ++i) {
    TrConv& c = T.conv[i];
    c.cin = cin[i];
    c.cout = cout[i];
    c.kpad = round_up(3 * cin[i], 16);
    const HostTensor* w = find(m, std::string(convs[i]) + ".weight");
    std::vector<float> w2((size_t)cout[i] * c.kpad, 0.f), sc((size_t)cout[i]), sh((size_t)cout[i]);
    for (int co = 0; co < cout[i]; ++co)
      for (int ci = 0; ci < cin[i]; ++ci)
        for (int dk = 0; dk < 3; ++dk) w2[(size_t)co * c.kpad + dk * cin[i] + ci] = w->data[((size_t)co * cin[i] + ci) * 3 + dk];
    const HostTensor *g = find(m, std::string(bns[i]) + ".weight"), *b = find(m, std::string(bns[i]) + ".bias"),
                     *mu = find(m, std::string(bns[i]) + ".running_mean"), *var = find(m, std::string(bns[i]) + ".running_var");
    for (int co = 0; co < cout[i]; ++co) {  // eval-mode BatchNorm1d as one affine map (eps 1e-5)
      const double s = (double)g->data[co] / sqrt((double)var->data[co] + 1e-5);
      sc[co] = (float)s;
      sh[co] = (float)((double)b->data[co] - (double)mu->data[co] * s);
    }
    CCSM_TRY(upload(c.w, w2));
    CCSM_TRY(upload(c.scale, sc));
    CCSM_TRY(upload(c.shift, sh));
  }
  T.layers.resize(m->cfg.num_layers);
  for (int l = 0; l < m->cfg.num_layers; ++l) {
    const std::string pre = "transformer_encoder.layers." + std::to_string(l) + ".";
    TrLayer& Lw = T.layers[l];
    CCSM_TRY(upload(Lw.in_w, find(m, pre + "self_attn.in_proj_weight")->data));
    CCSM_TRY(upload(Lw.in_b, find(m, pre + "self_attn.in_proj_bias")->data));
    CCSM_TRY(upload(Lw.out_w, find(m, pre + "self_attn.out_proj.weight")->data));
    CCSM_TRY(upload(Lw.out_b, find(m, pre + "self_attn.out_proj.bias")->data));
    CCSM_TRY(upload(Lw.l1_w, find(m, pre + "linear1.weight")->data));
    CCSM_TRY(upload(Lw.l1_b, find(m, pre + "linear1.bias")->data));
    CCSM_TRY(upload(Lw.l2_w, find(m, pre + "linear2.weight")->data));
    CCSM_TRY(upload(Lw.l2_b, find(m, pre + "linear2.bias")->data));
    CCSM_TRY(upload(Lw.n1_w, find(m, pre + "norm1.weight")->data));
    CCSM_TRY(upload(Lw.n1_b, find(m, pre + "norm1.bias")->data));
    CCSM_TRY(upload(Lw.n2_w, find(m, pre + "norm2.weight")->data));
    CCSM_TRY(upload(Lw.n2_b, find(m, pre + "norm2.bias")->data));
  }
  return CCSM_OK;
}

void trans_release(ccsm_model* m) {
  TrWeights& T = m->tr;
  for (auto& c : T.conv) { c.w.release(); c.scale.release(); c.shift.release(); }
  T.pos.release();
  for (auto& l : T.layers)
    for (DevBuf* b : {&l.in_w, &l.in_b, &l.out_w, &l.out_b, &l.l1_w, &l.l1_b, &l.l2_w, &l.l2_b, &l.n1_w, &l.n1_b, &l.n2_w, &l.n2_b})
      b->release();
  for (DevBuf* b : {&T.x0, &T.col, &T.a, &T.b, &T.qkv, &T.ffh, &T.ctx, &T.hid}) b->release();
}

int trans_forward(ccsm_model* m, int64_t n, const ccsm_strand* fwd, const ccsm_strand* rev, float* logits, float* probs,
                  cudaStream_t st) {
  const int L = m->cfg.seq_len, d = m->cfg.hidden, ff = m->dim_ff, C = m->cfg.num_classes, nhead = m->nhead;
  Fp32Weights& W = m->fp32;
  TrWeights& T = m->tr;
  const int K0 = round_up(m->in_feat, 16);
  const int64_t chunk = n < 4096 ? n : 4096;
  const int64_t tok_cap = chunk * 2 * L;
  if (tok_cap > T.tokens_cap) {
    int colw = 0;
    for (auto& c : T.conv) colw = c.kpad > colw ? c.kpad : colw;
    CCSM_TRY(T.x0.reserve(tok_cap * K0 * 4));
    CCSM_TRY(T.col.reserve(tok_cap * colw * 4));
    CCSM_TRY(T.a.reserve(tok_cap * d * 4));
    CCSM_TRY(T.b.reserve(tok_cap * d * 4));
    CCSM_TRY(T.qkv.reserve(tok_cap * 3 * d * 4));
    CCSM_TRY(T.ffh.reserve(tok_cap * (ff > d ? ff : d) * 4));
    CCSM_TRY(T.ctx.reserve(chunk * 2 * d * 4));
    CCSM_TRY(T.hid.reserve(chunk * 2 * d * 4));
    T.tokens_cap = tok_cap;
  }
  for (int64_t s0 = 0; s0 < n; s0 += chunk) {
    const int64_t sites = (n - s0) < chunk ? (n - s0) : chunk;
    const int64_t rows = sites * 2, tokens = rows * L;
    pack_x_att2s2_kernel<<<nblk(tokens, 256), 256, 0, st>>>(
        sites, L, m->cfg.n_embed, m->cfg.n_vocab, m->cfg.feat_flags, K0, offset_strand(fwd, s0, L), offset_strand(rev, s0, L),
        W.embed.as<float>(), W.ipd_embed.as<float>(), W.pw_embed.as<float>(), W.npass_embed.as<float>(), T.x0.as<float>());
    count_launch();
    // SrcEmbed: three conv stages; activations ping-pong between a and b
    const float* in = T.x0.as<float>();
    int ldin = K0;
    float* bufs[2] = {T.a.as<float>(), T.b.as<float>()};
    float* x = nullptr;
    for (int i = 0; i < 3; ++i) {
      TrConv& c = T.conv[i];
      tr_im2col_kernel<<<nblk(tokens * c.kpad, 256), 256, 0, st>>>(tokens, L, c.cin, ldin, c.kpad, in, T.col.as<float>());
      count_launch();
      float* conv_out = T.qkv.as<float>();  // scratch: (tokens, cout)
      CCSM_TRY(sgemm_nt((int)tokens, c.cout, c.kpad, T.col.as<float>(), c.kpad, 0, c.w.as<float>(), c.kpad, 0, nullptr, 0,
                        conv_out, c.cout, 0, 1, st));
      x = bufs[i & 1];
      tr_bn_relu_pool_kernel<<<nblk(tokens * c.cout, 256), 256, 0, st>>>(tokens, L, c.cout, conv_out, c.scale.as<float>(),
                                                                          c.shift.as<float>(),
                                                                          i == 2 ? T.pos.as<float>() : nullptr, x);
      count_launch();
      in = x;
      ldin = c.cout;
    }
    float* y = x == bufs[0] ? bufs[1] : bufs[0];
    for (auto& Lw : T.layers) {
      CCSM_TRY(sgemm_nt((int)tokens, 3 * d, d, x, d, 0, Lw.in_w.as<float>(), d, 0, Lw.in_b.as<float>(), 0, T.qkv.as<float>(),
                        3 * d, 0, 1, st));
      tr_attention_kernel<<<nblk(rows * nhead, 4), 128, 0, st>>>(rows, L, d, nhead, T.qkv.as<float>(), T.ffh.as<float>());
      count_launch();
      CCSM_TRY(sgemm_nt((int)tokens, d, d, T.ffh.as<float>(), d, 0, Lw.out_w.as<float>(), d, 0, Lw.out_b.as<float>(), 0, y, d, 0,
                        1, st));
      tr_add_ln_kernel<<<nblk(tokens, 8), 256, 0, st>>>(tokens, d, x, y, Lw.n1_w.as<float>(), Lw.n1_b.as<float>());
      count_launch();
      CCSM_TRY(sgemm_nt((int)tokens, ff, d, x, d, 0, Lw.l1_w.as<float>(), d, 0, Lw.l1_b.as<float>(), 0, T.ffh.as<float>(), ff, 0,
                        1, st));
      tr_relu_kernel<<<nblk(tokens * ff, 256), 256, 0, st>>>(tokens * ff, T.ffh.as<float>());
      count_launch();
      CCSM_TRY(sgemm_nt((int)tokens, d, ff, T.ffh.as<float>(), ff, 0, Lw.l2_w.as<float>(), ff, 0, Lw.l2_b.as<float>(), 0, y, d, 0,
                        1, st));
      tr_add_ln_kernel<<<nblk(tokens, 8), 256, 0, st>>>(tokens, d, x, y, Lw.n2_w.as<float>(), Lw.n2_b.as<float>());
      count_launch();
    }
    tr_mean_pool_kernel<<<nblk(sites * 2 * d, 256), 256, 0, st>>>(sites, L, d, x, T.ctx.as<float>());
    count_launch();
    const int D = 2 * d;
    CCSM_TRY(sgemm_nt((int)sites, D, D, T.ctx.as<float>(), D, 0, W.cls0_w.as<float>(), D, 0, W.cls0_b.as<float>(), 0,
                      T.hid.as<float>(), D, 0, 1, st));
    cls_out_kernel<4><<<nblk(sites, 4), 128, 0, st>>>(sites, D, C, T.hid.as<float>(), W.fc_w.as<float>(), W.fc_b.as<float>(),
                                                      logits ? logits + s0 * C : nullptr, probs ? probs + s0 * C : nullptr);
    count_launch();
    CCSM_CUDA(cudaGetLastError());
  }
  return CCSM_OK;
}

}  // namespace ccsm
