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
#include <stdlib.h>

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

// Gate nonlinearities: ex2.approx-based exp and an approximate reciprocal (about 2 ulp each) -- the accurate
// expf / tanhf / IEEE division cost as many issue slots as the FFMAs of this small model.  Absolute error ~1e-7 per
// evaluation; the model output stays within 1e-6 of the CPU port (tests/test_parity_gpu.py).
__device__ __forceinline__ float sigmoid_acc(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }
__device__ __forceinline__ float tanh_acc(float x) {
  const float a = fminf(fmaxf(x, -15.f), 15.f);
  return 1.f - __fdividef(2.f, 1.f + __expf(2.f * a));
}

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
              const float nn = tanh_acc(fmaf(r, ah[q][u], ai[q][u]));
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
        for (int i = 0; i < AG_H; ++i) et = fmaf(VA[i], tanh_acc(acc[i]), et);
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

// ------------------------------------------------------------------------------------------------------------
// Register-tiled formulation (default): a CTA owns 128 sites; for each direction and step the gate pre-activations
// are one small GEMM  [128 sites x 53] . [53 x 96]  whose operands both sit in shared memory -- x_t / h_{t-1} as
// [k][site] tiles, the weights as [k][unit pair][r0 r1 z0 z1 n0 n1] -- and each thread accumulates 8 sites x 2 units
// (64 accumulators): 48 FFMA per 4 shared-memory loads instead of 4 per load in the thread-per-site kernel above.
// The attention tail runs two threads per site (16 attention units each, combined with one shuffle).
// ------------------------------------------------------------------------------------------------------------
constexpr int TL_SITES = 128;  // (the index arithmetic below uses >> 7 / & 127)
constexpr int TL_THREADS = 256;

struct TiledLayout {
  int wx, wh, bias, wa, ua, va, fcw, fcb, xs, hs, total;
};
__host__ __device__ inline TiledLayout tiled_layout(int IN) {
  TiledLayout p;
  int o = 0;
  p.wx = o;   o += 2 * IN * 16 * 6;       // [dir][k][unit pair][r0 r1 z0 z1 n0 n1]
  p.wh = o;   o += 2 * AG_H * 16 * 6;
  o = (o + 3) & ~3;
  p.bias = o; o += 2 * 4 * AG_H;          // [dir][b_r, b_z, b_in, b_hn][unit]
  p.wa = o;   o += 2 * AG_H * AG_H;       // [k 0..63][unit]
  p.ua = o;   o += 2 * AG_H * AG_H;
  p.va = o;   o += AG_H;
  p.fcw = o;  o += 2 * AG_H;
  p.fcb = o;  o += 4;
  p.xs = o;   o += 2 * IN * TL_SITES;     // x_t tiles [2][k][site] (the next step's tile is prefetched)
  o = (o + 3) & ~3;
  p.hs = o;   o += 2 * AG_H * TL_SITES;   // h tiles [2][unit][site]
  p.total = o;
  return p;
}

template <int IN>
__global__ void __launch_bounds__(TL_THREADS, 2)
    aggr_tiled_kernel(const float* __restrict__ packed, int packed_floats, int64_t n, int L, const float* __restrict__ offsets,
                      const float* __restrict__ histos, const long long* __restrict__ site_pos, int only_close,
                      const float* __restrict__ h0, float* scratch, float* __restrict__ out) {
  extern __shared__ __align__(16) float sm[];
  const TiledLayout lay = tiled_layout(IN);
  for (int i = threadIdx.x; i < packed_floats; i += TL_THREADS) sm[i] = packed[i];
  float* xs = sm + lay.xs;
  float* hs = sm + lay.hs;
  __syncthreads();
  constexpr int BINS = IN - 1;
  const int tid = threadIdx.x;
  const int sg = tid >> 4, up = tid & 15;      // site group (8 sites), unit pair
  const int s_lo = sg * 8;
  float* my_scratch = scratch + (size_t)blockIdx.x * L * 2 * AG_H * TL_SITES;  // [t][unit 0..63][site]

  for (int64_t base = (int64_t)blockIdx.x * TL_SITES; base < n; base += (int64_t)gridDim.x * TL_SITES) {
#pragma unroll 1
    for (int dir = 0; dir < 2; ++dir) {
      const float* WX = sm + lay.wx + dir * IN * 96;
      const float* WH = sm + lay.wh + dir * AG_H * 96;
      const float* BS = sm + lay.bias + dir * 4 * AG_H;
      // h0 tile -> hs[0]
      for (int idx = tid; idx < TL_SITES * AG_H; idx += TL_THREADS) {
        const int site = idx & (TL_SITES - 1), k = idx >> 7;  // consecutive lanes -> consecutive sites: no bank conflicts
        const int64_t gs = base + site < n ? base + site : n - 1;
        hs[k * TL_SITES + site] = h0 ? h0[((size_t)dir * n + gs) * AG_H + k] : 0.f;
      }
      int cur = 0;
#pragma unroll 1
      for (int step = 0; step < L; ++step) {
        // x_t tiles: 20 histogram bins + 1 offset per site (materialised windows, or gathered from per-site rows).
        // Step 0 loads its own tile; every step then fetches the NEXT step's values into registers before its
        // GEMM and parks them in the other tile afterwards, so the global-memory latency hides behind the FFMAs.
        float* xcur = xs + (step & 1) * IN * TL_SITES;
        float* xnext = xs + ((step + 1) & 1) * IN * TL_SITES;
        auto fetch = [&](int tt, float4 (&v)[3], float& off) {
#pragma unroll
          for (int i = 0; i < 3; ++i) {
            const int idx = tid + i * TL_THREADS;
            v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (idx < TL_SITES * (BINS / 4)) {
              const int q = idx >> 7, site = idx & (TL_SITES - 1);  // lanes walk the sites: conflict-free tile stores
              const int64_t gs = base + site < n ? base + site : n - 1;
              if (site_pos) {
                const int64_t j = gs + tt - L / 2;
                if (j >= 0 && j < n) v[i] = __ldg(reinterpret_cast<const float4*>(histos + (size_t)j * BINS) + q);
              } else {
                v[i] = __ldg(reinterpret_cast<const float4*>(histos + ((size_t)gs * L + tt) * BINS) + q);
              }
            }
          }
          off = 0.f;
          if (tid < TL_SITES) {
            const int64_t gs = base + tid < n ? base + tid : n - 1;
            if (site_pos) {
              const int64_t j = gs + tt - L / 2;
              const long long centre = site_pos[gs];
              const long long pj = j < 0 ? site_pos[0] - 1000 : j >= n ? site_pos[n - 1] + 1000 : site_pos[j];
              if (only_close) {
                const long long pjm = j - 1 < 0 ? site_pos[0] - 1000 : j - 1 >= n ? site_pos[n - 1] + 1000 : site_pos[j - 1];
                off = (pj - pjm == 2) ? 1.f : 0.f;
              } else {
                off = (float)llabs(pj - centre);
              }
            } else {
              off = __ldg(offsets + (size_t)gs * L + tt);
            }
          }
        };
        auto park = [&](float* xt, const float4 (&v)[3], float off) {
#pragma unroll
          for (int i = 0; i < 3; ++i) {
            const int idx = tid + i * TL_THREADS;
            if (idx < TL_SITES * (BINS / 4)) {
              const int q = idx >> 7, site = idx & (TL_SITES - 1);
              xt[(q * 4 + 0) * TL_SITES + site] = v[i].x;
              xt[(q * 4 + 1) * TL_SITES + site] = v[i].y;
              xt[(q * 4 + 2) * TL_SITES + site] = v[i].z;
              xt[(q * 4 + 3) * TL_SITES + site] = v[i].w;
            }
          }
          if (tid < TL_SITES) xt[BINS * TL_SITES + tid] = off;
        };
        const int t = dir ? L - 1 - step : step;
        float4 pv[3];
        float poff;
        if (step == 0) {
          fetch(t, pv, poff);
          park(xcur, pv, poff);
        }
        __syncthreads();
        const bool more = step + 1 < L;
        if (more) fetch(dir ? t - 1 : t + 1, pv, poff);
        const float* hc = hs + cur * AG_H * TL_SITES;
        float* hn_t = hs + (cur ^ 1) * AG_H * TL_SITES;
        float ar[8][2], az[8][2], ai[8][2], ah[8][2];
        {
          const float br0 = BS[0 * AG_H + 2 * up], br1 = BS[0 * AG_H + 2 * up + 1];
          const float bz0 = BS[1 * AG_H + 2 * up], bz1 = BS[1 * AG_H + 2 * up + 1];
          const float bi0 = BS[2 * AG_H + 2 * up], bi1 = BS[2 * AG_H + 2 * up + 1];
          const float bh0 = BS[3 * AG_H + 2 * up], bh1 = BS[3 * AG_H + 2 * up + 1];
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            ar[q][0] = br0; ar[q][1] = br1; az[q][0] = bz0; az[q][1] = bz1;
            ai[q][0] = bi0; ai[q][1] = bi1; ah[q][0] = bh0; ah[q][1] = bh1;
          }
        }
#pragma unroll 3
        for (int k = 0; k < IN; ++k) {
          const float4 a0 = *reinterpret_cast<const float4*>(xcur + k * TL_SITES + s_lo);
          const float4 a1 = *reinterpret_cast<const float4*>(xcur + k * TL_SITES + s_lo + 4);
          const float* wp = WX + (k * 16 + up) * 6;
          const float2 wr = *reinterpret_cast<const float2*>(wp), wz = *reinterpret_cast<const float2*>(wp + 2),
                       wn = *reinterpret_cast<const float2*>(wp + 4);
          const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            ar[q][0] = fmaf(a[q], wr.x, ar[q][0]); ar[q][1] = fmaf(a[q], wr.y, ar[q][1]);
            az[q][0] = fmaf(a[q], wz.x, az[q][0]); az[q][1] = fmaf(a[q], wz.y, az[q][1]);
            ai[q][0] = fmaf(a[q], wn.x, ai[q][0]); ai[q][1] = fmaf(a[q], wn.y, ai[q][1]);
          }
        }
#pragma unroll 4
        for (int k = 0; k < AG_H; ++k) {
          const float4 a0 = *reinterpret_cast<const float4*>(hc + k * TL_SITES + s_lo);
          const float4 a1 = *reinterpret_cast<const float4*>(hc + k * TL_SITES + s_lo + 4);
          const float* wp = WH + (k * 16 + up) * 6;
          const float2 wr = *reinterpret_cast<const float2*>(wp), wz = *reinterpret_cast<const float2*>(wp + 2),
                       wn = *reinterpret_cast<const float2*>(wp + 4);
          const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            ar[q][0] = fmaf(a[q], wr.x, ar[q][0]); ar[q][1] = fmaf(a[q], wr.y, ar[q][1]);
            az[q][0] = fmaf(a[q], wz.x, az[q][0]); az[q][1] = fmaf(a[q], wz.y, az[q][1]);
            ah[q][0] = fmaf(a[q], wn.x, ah[q][0]); ah[q][1] = fmaf(a[q], wn.y, ah[q][1]);
          }
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const int unit = 2 * up + u;
          const float4 p0 = *reinterpret_cast<const float4*>(hc + unit * TL_SITES + s_lo);
          const float4 p1 = *reinterpret_cast<const float4*>(hc + unit * TL_SITES + s_lo + 4);
          const float hp[8] = {p0.x, p0.y, p0.z, p0.w, p1.x, p1.y, p1.z, p1.w};
          float hv[8];
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const float r = sigmoid_acc(ar[q][u]);
            const float z = sigmoid_acc(az[q][u]);
            const float nn = tanh_acc(fmaf(r, ah[q][u], ai[q][u]));
            hv[q] = fmaf(z, hp[q] - nn, nn);  // (1 - z) * n + z * h
          }
          const float4 o0 = make_float4(hv[0], hv[1], hv[2], hv[3]), o1 = make_float4(hv[4], hv[5], hv[6], hv[7]);
          *reinterpret_cast<float4*>(hn_t + unit * TL_SITES + s_lo) = o0;
          *reinterpret_cast<float4*>(hn_t + unit * TL_SITES + s_lo + 4) = o1;
          float* sc = my_scratch + ((size_t)t * 2 * AG_H + dir * AG_H + unit) * TL_SITES + s_lo;
          *reinterpret_cast<float4*>(sc) = o0;
          *reinterpret_cast<float4*>(sc + 4) = o1;
        }
        if (more) park(xnext, pv, poff);
        __syncthreads();
        cur ^= 1;
      }
    }
    __syncthreads();  // every thread's scratch writes are visible to the block (global memory, same CTA)

    // ---- attention (utils/attention.py:48-70), register-tiled like the GRU: per step one GEMM
    // [128 sites x 64] . [64 x 32 attention units]; thread = 4 sites x 4 units; the out_t tile comes back from the
    // scratch slab into the (now idle) h tiles as [k][site].  e_t and fc1 . out_t land in small shared arrays.
    {
      float* tile = hs;                       // [64][128]
      float* e_s = xs;                        // [L][128]   (x tiles are idle now)
      float* g_s = xs + AG_MAX_L * TL_SITES;  // [L][128]
      // mapping: 32 site groups of 4 sites x 8 unit groups of 4 attention units = 16 accumulators per thread, and every
      // reduction over the attention units stays inside 8 adjacent lanes
      const int ug = tid & 7;
      const int sl = (tid >> 3) * 4;
      const float* WA = sm + lay.wa;
      const float* UA = sm + lay.ua;
      const float* VA = sm + lay.va + ug * 4;
      const float* FW = sm + lay.fcw;
      auto load_tile = [&](int t_f, int t_r) {  // rows 0..31 from step t_f (forward units), 32..63 from step t_r
        for (int idx = tid; idx < 2 * AG_H * TL_SITES / 4; idx += TL_THREADS) {
          const int k = idx / (TL_SITES / 4), c = idx - k * (TL_SITES / 4);
          const int tt = k < AG_H ? t_f : t_r;
          reinterpret_cast<float4*>(tile)[idx] =
              *reinterpret_cast<const float4*>(my_scratch + ((size_t)tt * 2 * AG_H + k) * TL_SITES + c * 4);
        }
      };
      auto gemm = [&](const float* W, float (&acc)[4][4], float (&gp)[4]) {
#pragma unroll 4
        for (int k = 0; k < 2 * AG_H; ++k) {
          const float4 a = *reinterpret_cast<const float4*>(tile + k * TL_SITES + sl);
          const float4 w = *reinterpret_cast<const float4*>(W + k * AG_H + ug * 4);
          const float av[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            acc[q][0] = fmaf(av[q], w.x, acc[q][0]); acc[q][1] = fmaf(av[q], w.y, acc[q][1]);
            acc[q][2] = fmaf(av[q], w.z, acc[q][2]); acc[q][3] = fmaf(av[q], w.w, acc[q][3]);
          }
          if ((k >> 3) == ug) {  // fc1 . out_t: each of the 8 threads of a site group takes 8 of the 64 inputs
            const float f = FW[k];
#pragma unroll
            for (int q = 0; q < 4; ++q) gp[q] = fmaf(av[q], f, gp[q]);
          }
        }
      };
      // query projection: q = [h_n fwd (step L-1) | h_n rev (step 0)]
      float wq[4][4], dummy[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int u = 0; u < 4; ++u) wq[q][u] = 0.f;
      load_tile(L - 1, 0);
      __syncthreads();
      gemm(WA, wq, dummy);
      __syncthreads();
#pragma unroll 1
      for (int t = 0; t < L; ++t) {
        load_tile(t, t);
        __syncthreads();
        float acc[4][4], gp[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int q = 0; q < 4; ++q)
#pragma unroll
          for (int u = 0; u < 4; ++u) acc[q][u] = wq[q][u];
        gemm(UA, acc, gp);
        float ep[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          ep[q] = VA[0] * tanh_acc(acc[q][0]) + VA[1] * tanh_acc(acc[q][1]) + VA[2] * tanh_acc(acc[q][2]) +
                  VA[3] * tanh_acc(acc[q][3]);
#pragma unroll
          for (int o = 1; o < 8; o <<= 1) {
            ep[q] += __shfl_xor_sync(0xffffffffu, ep[q], o);
            gp[q] += __shfl_xor_sync(0xffffffffu, gp[q], o);
          }
        }
        if (ug == 0) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            e_s[t * TL_SITES + sl + q] = ep[q];
            g_s[t * TL_SITES + sl + q] = gp[q];
          }
        }
        __syncthreads();  // tile is overwritten by the next step; e_s / g_s complete for the softmax below
      }
      if (tid < TL_SITES) {
        float mx = -INFINITY;
        for (int t = 0; t < L; ++t) mx = fmaxf(mx, e_s[t * TL_SITES + tid]);
        float den = 0.f, num = 0.f;
        for (int t = 0; t < L; ++t) {
          const float p = expf(e_s[t * TL_SITES + tid] - mx);
          den += p;
          num = fmaf(p, g_s[t * TL_SITES + tid], num);
        }
        if (base + tid < n) out[base + tid] = num / den + sm[lay.fcb];
      }
    }
    __syncthreads();  // the scratch slab and tiles are reused by the next item
  }
}

// ---- host side ----------------------------------------------------------------------------------------------
static const HostTensor* findw(ccsm_model* m, const std::string& k) {
  auto it = m->w.find(k);
  return it == m->w.end() ? nullptr : &it->second;
}

bool aggr_fused_supported(const ccsm_model* m) {
  return m->cfg.kind == CCSM_KIND_AGGR && m->gates == 3 && m->cfg.hidden == AG_H && m->cfg.num_layers == 1 && m->cfg.seq_len <= AG_MAX_L &&
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
  // the register-tiled kernel's layout: GRU weights regrouped per unit pair, the rest copied
  const TiledLayout tl = tiled_layout(IN);
  std::vector<float> t((size_t)tl.xs, 0.f);
  for (int d = 0; d < 2; ++d)
    for (int upair = 0; upair < 16; ++upair)
      for (int gate = 0; gate < 3; ++gate)
        for (int u = 0; u < 2; ++u) {
          const int j = 2 * upair + u;
          for (int k = 0; k < IN; ++k)
            t[tl.wx + ((d * IN + k) * 16 + upair) * 6 + gate * 2 + u] = p[lay.wx + ((d * 3 + gate) * IN + k) * H + j];
          for (int k = 0; k < H; ++k)
            t[tl.wh + ((d * H + k) * 16 + upair) * 6 + gate * 2 + u] = p[lay.wh + ((d * 3 + gate) * H + k) * H + j];
        }
  std::copy(p.begin() + lay.bias, p.begin() + lay.bias + 2 * 4 * H, t.begin() + tl.bias);
  std::copy(p.begin() + lay.wa, p.begin() + lay.wa + 2 * H * H, t.begin() + tl.wa);
  std::copy(p.begin() + lay.ua, p.begin() + lay.ua + 2 * H * H, t.begin() + tl.ua);
  std::copy(p.begin() + lay.va, p.begin() + lay.va + H, t.begin() + tl.va);
  std::copy(p.begin() + lay.fcw, p.begin() + lay.fcw + 2 * H, t.begin() + tl.fcw);
  t[tl.fcb] = p[lay.fcb];
  CCSM_TRY(m->aggr_packed_tiled.reserve(t.size() * sizeof(float)));
  CCSM_CUDA(cudaMemcpy(m->aggr_packed_tiled.p, t.data(), t.size() * sizeof(float), cudaMemcpyHostToDevice));
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
  int sms = 0;
  CCSM_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, m->cfg.device));
  const char* env = getenv("CCSM_AGGR_TILED");
  if (!(env && atoi(env) == 0)) {
    // register-tiled kernel (default); CCSM_AGGR_TILED=0 selects the thread-per-site kernel below
    const TiledLayout tl = tiled_layout(IN);
    const size_t smem_t = (size_t)tl.total * sizeof(float);
    static bool attr_t = false;
    if (!attr_t) {
      CCSM_CUDA(cudaFuncSetAttribute(aggr_tiled_kernel<21>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_t));
      attr_t = true;
    }
    const int64_t tiles_t = (n + TL_SITES - 1) / TL_SITES;
    const int grid_t = (int)(tiles_t < 2LL * sms ? tiles_t : 2LL * sms);
    CCSM_TRY(m->aggr_scratch.reserve((size_t)grid_t * m->cfg.seq_len * 2 * AG_H * TL_SITES * sizeof(float)));
    aggr_tiled_kernel<21><<<grid_t, TL_THREADS, smem_t, st>>>(m->aggr_packed_tiled.as<float>(), tl.xs, n, m->cfg.seq_len, offsets,
                                                              histos, site_pos, only_close, h0, m->aggr_scratch.as<float>(), out);
    count_launch();
    CCSM_CUDA(cudaGetLastError());
    return CCSM_OK;
  }
  const AggrPacked lay = aggr_layout(IN, C);
  constexpr int COLS = AG_S * AG_THREADS;
  const size_t smem = ((size_t)lay.total + (size_t)AG_H * COLS) * sizeof(float);
  static bool attr = false;
  if (!attr) {
    CCSM_CUDA(cudaFuncSetAttribute(aggr_fused_kernel<21, AG_S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = true;
  }
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
