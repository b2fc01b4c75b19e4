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
                      float* out, cudaStream_t st) {
  const int L = m->cfg.seq_len, Bn = m->cfg.feat_flags;
  const int64_t chunk = n < 65536 ? n : 65536;
  CCSM_TRY(reserve_ws(m, chunk));
  for (int64_t s0 = 0; s0 < n; s0 += chunk) {
    int64_t sites = (n - s0) < chunk ? (n - s0) : chunk;
    pack_x_aggr_kernel<<<nblk(sites * L, 256), 256, 0, st>>>(sites, L, Bn, m->fp32.layers[0].Kpad,
                                                             offsets + s0 * L, histos + s0 * L * Bn,
                                                             m->ws32.x0.as<float>());
    count_launch();
    CCSM_TRY(run_stack(m, sites, s0, n, h0, nullptr, nullptr, out, st));
  }
  return CCSM_OK;
}

}  // namespace ccsm
