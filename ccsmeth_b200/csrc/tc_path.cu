// Tensor-core (tcgen05 / TMEM / TMA) implementation of the attbigru2s forward for sm_100a.
//
// Replaces: reference ccsmeth/models.py:89-150 (ModelAttRNN.forward) + utils/attention.py:48-70.
//
// Data model (everything a kernel streams is a pre-tiled "image" so that one 1-D bulk copy lands an
// MMA-ready operand in shared memory; see DESIGN.md "HBM layout"):
//   row tile  = 128 strand-rows (row R = 2*site + strand) = 64 CpG sites; UMMA M = 128 (TMEM lane = row)
//   slab      = 8 consecutive K elements of all rows of a tile, rows 16 B apart: (rows x 16 B) contiguous.
//               This is the UMMA K-major SWIZZLE_NONE canonical layout with SBO = 128 B, LBO = slab bytes.
//   x0 image  [tile][t][part][2 slabs x 2048 B]                     layer-0 input, K = 11 padded to 16
//   act image [tile][t][chunk c = dir*4 + j][part][8 slabs x 2048]  layer output h_t, 64 hidden units per chunk
//   h0 image  [layer][tile][dir][kc][part][8 slabs x 2048]          initial hidden state
//   weights   [dir][j] { X: [part][KX/8 slabs x 3072 B], H: [part][32 slabs x 3072 B] }
//             192 gate rows per unit-chunk j: X part rows = (n_i, r, z), H part rows = (r, z, n_h)
//   part      = 0 (hi) or 1 (lo): P = 1 single pass; P = 2 -> x ~= hi + lo and three MMA passes
//               hi*hi + hi*lo + lo*hi (error-compensated split, fp32 accumulate in TMEM).
//   C8        = "fp16 + e4m3 corrections" (CCSM_PREC_FP16C8, P = 2): the two correction products of the split are
//               only ~2^-12 of the result, so 4 significant bits are enough for them:
//                 a . W ~= a_hi . W_hi [fp16 MMA] + e4m3(a) . e4m3(S W_lo) + e4m3(S a_lo) . e4m3(W) [e4m3 MMAs, K = 32 each]
//               with everything accumulated at scale S = 2^12 (W_hi = fp16(S W), biases x S; the gate math folds 1/S
//               into its exponent constants).  Two MMA pass-equivalents instead of three, max |dprob| ~3e-5
//               (scripts/precision_study2.py).  Part 1 of an image then holds, per 32 K elements, two 16-element
//               e4m3 slabs of the rounded value followed by two of the scaled residual (same bytes as a 16-bit lo part):
//                 activations [a8 s0 s1][alo8 s0 s1] x 2048 B, weights [Wlo8 s0 s1][W8 s0 s1] x (rows x 16 B).
//               The K = 11 input of layer 0 keeps the 3-pass fp16 split (weights x S).
//
// Kernels in this file (what runs by default: gru_variant()):
//   tc_prep_kernel        features -> x0 image, h0 -> h0 images
//   tc_gru_layer_kernel   one launch per GRU layer, persistent; warp roles = TMA producers (one bulk copy per thread and
//                         stage) | MMA issuer (one elected thread, N = 192 per instruction) | gate epilogue warps
//                         (tcgen05.ld accumulators -> sigmoid / tanh / blend -> h_t into the act image, which is also how
//                         h_t reaches the next step's A operand).  Template forms: row tiles per CTA, TMEM buffers,
//                         epilogue warps, pipelined epilogue, both directions interleaved (IL: the default for layer 0).
//   tc_gru_pair2_kernel   the same layer on CTA pairs: tcgen05.mma.cta_group::2 (M = 256), each CTA holds half of every
//                         weight slab, operands by tensor-map loads that complete on the leader's barrier (default for
//                         layers >= 1)
//   tc_gru_pair_kernel / tc_gru_duo_kernel / tc_gru_cv_kernel   earlier CTA-pair forms (relay warp) and the on-chip
//                         operand-conversion form: measured experiments, selectable with CCSM_TC_VARIANT
//   tc_att_head_kernel    attention scores, online softmax, fc1, softmax: one pass over the last layer's act image
#include <curand_kernel.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#include "ccsm_internal.h"
#include "tc_common.cuh"
#include "tmap.h"

namespace ccsm {
using namespace tc;

constexpr int TILE_ROWS = 128;
constexpr uint32_t A_SLAB = 2048;     // 128 rows x 16 B
constexpr uint32_t G_SLAB = 3072;     // 192 gate rows x 16 B
constexpr uint32_t T_SLAB = 4096;     // 256 attention rows x 16 B
constexpr uint32_t CHUNK_BYTES = 8 * A_SLAB;  // 64 K elements of a 128-row tile, one part
constexpr float C8_S = 4096.f;                // accumulator scale of the C8 mode
constexpr float C8_INV_S = 1.f / 4096.f;
// C8 part 1 of an activation chunk: byte offset (before + row * 16) of the 8 e4m3 values of 8-unit slab q (0..7):
// rounded values; the scaled residuals sit 4096 B further on.
__host__ __device__ constexpr uint32_t c8_off(int q) { return (uint32_t)(q >> 2) * 8192u + (uint32_t)((q >> 1) & 1) * 2048u + (uint32_t)(q & 1) * 8u; }

struct TcState {
  int P = 0;          // parts of the currently packed weights (0 = none)
  bool f16 = false;
  bool c8 = false;    // fp16 + e4m3 corrections (CCSM_PREC_FP16C8)
  std::vector<DevBuf> wimg;   // per layer
  std::vector<DevBuf> wpair;  // per layer, CTA-pair layout (each CTA's 96-row half contiguous)
  std::vector<size_t> kx_slabs;
  DevBuf bias;                // [layer][dir][4][H]
  DevBuf wa_img, ua_img;      // [part][64 slabs x 4096]
  DevBuf va, fc_w, fc_b, embed;
  // workspace
  int64_t tiles_cap = 0;
  int ws_P = 0;
  DevBuf x0img, h0img, act[3];
  int sm_count = 0;
  // debug bookkeeping
  int64_t last_tiles = 0;
};

// ------------------------------------------------------------------------------------------------
// math
// ------------------------------------------------------------------------------------------------
// FAST: the argument arrives pre-halved (the r/z gate rows and biases are scaled by 0.5 when the weight images
// are packed -- exact in bf16/fp16), so sigmoid(2x') = 0.5 * tanh(x') + 0.5 is one MUFU + one FFMA.
template <bool FAST>
__device__ __forceinline__ float sigmoid_(float x) {
  if constexpr (FAST) {
    float t;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(x));
    return fmaf(t, 0.5f, 0.5f);
  } else {
    return __fdividef(1.f, 1.f + __expf(-x));
  }
}
template <bool FAST>
__device__ __forceinline__ float tanh_(float x) {
  if constexpr (FAST) {
    float t;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(x));
    return t;
  } else {
    return fmaf(2.f, __fdividef(1.f, 1.f + __expf(-2.f * x)), -1.f);
  }
}

// Accurate forms on accumulators that carry the scale S (C8): sigmoid(x / S), tanh(x / S); ex2.approx + fast division.
__device__ __forceinline__ float ex2_(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float sigmoid_s(float x) { return __fdividef(1.f, 1.f + ex2_(x * (-1.4426950408889634f * C8_INV_S))); }
__device__ __forceinline__ float tanh_s(float x) {
  return fmaf(2.f, __fdividef(1.f, 1.f + ex2_(x * (-2.8853900817779268f * C8_INV_S))), -1.f);
}
// C8: both gate sigmoids of a unit with ONE reciprocal: r = (1 + eb) / ((1 + ea)(1 + eb)), z = (1 + ea) / (...): 3 MUFU
// instead of 4 (the gate epilogue of layer 0 sits on the MUFU pipe).  Pre-activations are clamped at -40 so that the
// product stays finite (sigmoid(-40) = 4e-18).
__device__ __forceinline__ void sigmoid2_s(float a, float b, float& r, float& z) {
  const float lo = -40.f * C8_S;
  const float pa = 1.f + ex2_(fmaxf(a, lo) * (-1.4426950408889634f * C8_INV_S));
  const float pb = 1.f + ex2_(fmaxf(b, lo) * (-1.4426950408889634f * C8_INV_S));
  const float inv = __fdividef(1.f, pa * pb);
  r = inv * pb;
  z = inv * pa;
}
template <bool FAST, bool C8>
__device__ __forceinline__ float sig_(float x) {
  if constexpr (C8) return sigmoid_s(x);
  else return sigmoid_<FAST>(x);
}
template <bool FAST, bool C8>
__device__ __forceinline__ float tnh_(float x) {
  if constexpr (C8) return tanh_s(x);
  else return tanh_<FAST>(x);
}

// C8: 8 fp32 -> fp16 hi (16 bytes), e4m3 of the value (8 bytes), e4m3 of S * (value - hi) (8 bytes)
__device__ __forceinline__ void split8_c8(const float (&v)[8], uint4& hi, uint2& a8, uint2& l8) {
  uint32_t h[4];
  float r[8];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    h[i] = pack2<true>(v[2 * i], v[2 * i + 1]);
    const float2 b = unpack2<true>(h[i]);
    r[2 * i] = (v[2 * i] - b.x) * C8_S;
    r[2 * i + 1] = (v[2 * i + 1] - b.y) * C8_S;
  }
  hi = make_uint4(h[0], h[1], h[2], h[3]);
  a8 = make_uint2(pack4_e4m3(v[0], v[1], v[2], v[3]), pack4_e4m3(v[4], v[5], v[6], v[7]));
  l8 = make_uint2(pack4_e4m3(r[0], r[1], r[2], r[3]), pack4_e4m3(r[4], r[5], r[6], r[7]));
}
// C8: value = hi + residual / S
__device__ __forceinline__ void join8_c8(const uint4& hi, const uint2& l8, float (&v)[8]) {
  const uint32_t h[4] = {hi.x, hi.y, hi.z, hi.w};
  float r0[4], r1[4];
  unpack4_e4m3(l8.x, r0);
  unpack4_e4m3(l8.y, r1);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 a = unpack2<true>(h[i]);
    const float ra = i < 2 ? r0[2 * i] : r1[2 * i - 4], rb = i < 2 ? r0[2 * i + 1] : r1[2 * i - 3];
    v[2 * i] = fmaf(ra, C8_INV_S, a.x);
    v[2 * i + 1] = fmaf(rb, C8_INV_S, a.y);
  }
}

// 8 fp32 -> one 16-byte vector of packed elements (hi) and optionally the residual (lo)
template <int P, bool F16>
__device__ __forceinline__ void split8(const float (&v)[8], uint4& hi, uint4& lo) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    h[i] = pack2<F16>(v[2 * i], v[2 * i + 1]);
    if constexpr (P == 2) {
      float2 b = unpack2<F16>(h[i]);
      l[i] = pack2<F16>(v[2 * i] - b.x, v[2 * i + 1] - b.y);
    } else {
      l[i] = 0;
    }
  }
  hi = make_uint4(h[0], h[1], h[2], h[3]);
  lo = make_uint4(l[0], l[1], l[2], l[3]);
}
template <int P, bool F16>
__device__ __forceinline__ void join8(const uint4& hi, const uint4& lo, float (&v)[8]) {
  const uint32_t h[4] = {hi.x, hi.y, hi.z, hi.w};
  const uint32_t l[4] = {lo.x, lo.y, lo.z, lo.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 a = unpack2<F16>(h[i]);
    if constexpr (P == 2) {
      float2 b = unpack2<F16>(l[i]);
      a.x += b.x;
      a.y += b.y;
    }
    v[2 * i] = a.x;
    v[2 * i + 1] = a.y;
  }
}

// Writes the gate biases of hidden units [u0, u0+16) into the four accumulator column groups of a buffer
// (tcgen05.st), so that every MMA of the next unit-chunk simply accumulates on top of them:
//   [0,64) n_i <- b_in, [64,128) r <- b_ir+b_hr, [128,192) z <- b_iz+b_hz, [192,256) n_h <- b_hn.
// bias_d: this CTA's direction, [4][256] floats in shared memory (gate order r, z, in, hn).
__device__ __forceinline__ void arm_bias16(uint32_t trow, int ub, const float* bias_d, int u0) {
#pragma unroll 1
  for (int g = 0; g < 4; ++g) {  // not unrolled: keeps only 16 staging registers live
    uint32_t v[16];
    const int goff = g == 0 ? 512 : (g == 1 ? 0 : (g == 2 ? 256 : 768));  // column group -> bias row: n_i, r, z, n_h
    const float4* src = reinterpret_cast<const float4*>(bias_d + goff + u0);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float4 b = src[q];
      v[4 * q + 0] = __float_as_uint(b.x);
      v[4 * q + 1] = __float_as_uint(b.y);
      v[4 * q + 2] = __float_as_uint(b.z);
      v[4 * q + 3] = __float_as_uint(b.w);
    }
    tmem_st16(trow + g * 64 + ub * 16, v);
  }
}

// Same for 8 hidden units [u0, u0+8) at accumulator columns col0 + {0, 64, 128, 192} (pipelined epilogue).
__device__ __forceinline__ void arm_bias8(uint32_t trow, int col0, const float* bias_d, int u0) {
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    uint32_t v[8];
    const int goff = g == 0 ? 512 : (g == 1 ? 0 : (g == 2 ? 256 : 768));
    const float4* src = reinterpret_cast<const float4*>(bias_d + goff + u0);
    const float4 b0 = src[0], b1 = src[1];
    v[0] = __float_as_uint(b0.x); v[1] = __float_as_uint(b0.y); v[2] = __float_as_uint(b0.z); v[3] = __float_as_uint(b0.w);
    v[4] = __float_as_uint(b1.x); v[5] = __float_as_uint(b1.y); v[6] = __float_as_uint(b1.z); v[7] = __float_as_uint(b1.w);
    tmem_st8(trow + g * 64 + col0, v);
  }
}

// ------------------------------------------------------------------------------------------------
// prep: features -> x0 image (embedding lookup + kinetics concat, reference models.py:91-106),
//       h0 (2*layers, n, H) fp32 -> h0 images.  One thread per strand-row.
// ------------------------------------------------------------------------------------------------
struct TcStrand {
  const float *kmer, *kpass, *ipd, *pw;
};

template <int P, bool F16, bool C8 = false>
__global__ void tc_prep_kernel(int64_t n_tiles, int64_t sites, int64_t site0, int64_t n_total, int L, int NL,
                               int n_vocab, int has_npass, TcStrand s0, TcStrand s1,
                               const float* __restrict__ embed, const float* __restrict__ h0_a,
                               const float* __restrict__ h0_b, int h0_random, unsigned long long h0_seed,
                               unsigned long long h0_offset, uint8_t* __restrict__ x0img,
                               uint8_t* __restrict__ h0img) {
  const int64_t gr = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // global row in chunk
  if (gr >= n_tiles * TILE_ROWS) return;
  const int64_t tile = gr / TILE_ROWS;
  const int r = (int)(gr % TILE_ROWS);
  const int strand = (int)(gr & 1);
  const int64_t site = gr >> 1;  // site within chunk
  const bool valid = site < sites;
  const TcStrand& s = strand ? s1 : s0;
  for (int t = 0; t < L; ++t) {
    float v[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = 0.f;
    if (valid) {
      const int64_t o = (site0 + site) * L + t;
      int code = (int)s.kmer[o];
      code = code < 0 ? 0 : (code >= n_vocab ? n_vocab - 1 : code);
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = embed[code * 8 + i];
      v[8] = s.ipd[o];
      v[9] = s.pw[o];
      if (has_npass) v[10] = s.kpass[o];
    }
#pragma unroll
    for (int sl = 0; sl < 2; ++sl) {
      float w[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) w[i] = v[sl * 8 + i];
      uint4 hi, lo;
      split8<P, F16>(w, hi, lo);
      uint8_t* base = x0img + ((tile * L + t) * P) * (2 * (size_t)A_SLAB) + sl * A_SLAB + r * 16;
      *reinterpret_cast<uint4*>(base) = hi;
      if constexpr (P == 2) *reinterpret_cast<uint4*>(base + 2 * A_SLAB) = lo;
    }
  }
  const float* h0 = strand ? h0_b : h0_a;
  for (int l = 0; l < NL; ++l)
    for (int d = 0; d < 2; ++d) {
      const float* src = (valid && h0) ? h0 + ((int64_t)(2 * l + d) * n_total + site0 + site) * 256 : nullptr;
      const bool rnd = valid && !h0 && h0_random;
      curandStatePhilox4_32_10_t rng;
      if (rnd)
        curand_init(h0_seed, (unsigned long long)(((site0 + site) * 2 + strand) * 2 * NL + 2 * l + d), h0_offset, &rng);
      for (int kc = 0; kc < 4; ++kc)
#pragma unroll
        for (int sl = 0; sl < 8; ++sl) {
          float w[8];
          if (rnd) {
            const float4 a = curand_normal4(&rng), b = curand_normal4(&rng);
            w[0] = a.x; w[1] = a.y; w[2] = a.z; w[3] = a.w; w[4] = b.x; w[5] = b.y; w[6] = b.z; w[7] = b.w;
          } else if (src) {
            float4 a = *reinterpret_cast<const float4*>(src + kc * 64 + sl * 8);
            float4 b = *reinterpret_cast<const float4*>(src + kc * 64 + sl * 8 + 4);
            w[0] = a.x; w[1] = a.y; w[2] = a.z; w[3] = a.w; w[4] = b.x; w[5] = b.y; w[6] = b.z; w[7] = b.w;
          } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) w[i] = 0.f;
          }
          uint8_t* cbase = h0img + ((((int64_t)l * n_tiles + tile) * 2 + d) * 4 + kc) * (size_t)(P * CHUNK_BYTES);
          uint8_t* base = cbase + sl * A_SLAB + r * 16;
          if constexpr (C8) {
            uint4 hi;
            uint2 a8, l8;
            split8_c8(w, hi, a8, l8);
            *reinterpret_cast<uint4*>(base) = hi;
            uint8_t* b8 = cbase + CHUNK_BYTES + c8_off(sl) + r * 16;
            *reinterpret_cast<uint2*>(b8) = a8;
            *reinterpret_cast<uint2*>(b8 + 4096) = l8;
          } else {
            uint4 hi, lo;
            split8<P, F16>(w, hi, lo);
            *reinterpret_cast<uint4*>(base) = hi;
            if constexpr (P == 2) *reinterpret_cast<uint4*>(base + CHUNK_BYTES) = lo;
          }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// GRU layer kernel
// ------------------------------------------------------------------------------------------------
struct GruParams {
  const uint8_t* xin;    // layer 0: x0 image; else act image of the previous layer
  const uint8_t* h0img;  // this layer's h0 images [tile][dir][kc][part][CHUNK]
  uint8_t* out;          // this layer's act image
  const uint8_t* wimg;   // this layer's weight images
  const float* bias;     // [dir][4][256]: b_r(=b_ir+b_hr), b_z, b_in, b_hn
  int n_tiles;           // even
  int L;
  int kx_slabs;          // 2 (layer 0) or 64
  int l2_hint;           // 1: bulk loads carry an L2 evict_last hint (experiment, CCSM_TC_L2HINT)
};

// NSLOT = row tiles processed together by one CTA (sharing every weight stage).
//   NSLOT = 2: one CTA per SM, 384 threads, 3 stages of 56 KB, TMEM 2 x 256 columns (one buffer per slot).
//   NSLOT = 1: two CTAs per SM, 192 threads each, 4 stages of 20 KB, TMEM 256 columns per CTA: the two
//              co-resident CTAs overlap one's gate epilogue with the other's MMAs.
//   NBUF  = accumulator buffers per slot in TMEM (2: the MMAs of unit-chunk j+1 overlap the epilogue of j).
//   KSB   = K-slabs (of 8 elements) per stage and part in the single-pass modes (x3 modes: KSB / 2, same bytes).
//           Bigger stages = more MMAs per full-barrier wait + commit of the single issuing thread (measured:
//           the issue loop, not the tensor pipe, bounds the kernel at 2 MMAs per stage).
//   (NSLOT, NBUF) = (1, 2): one CTA per SM, 192 threads, TMEM 2 x 256 columns; half the row tiles in flight of the
//           other variants, so the activation re-reads stay in L2.
//   EPIW  = epilogue warps per TMEM lane quadrant (2: two warps share a row, each takes two of the four 16-unit
//           blocks of a unit-chunk -- halves the epilogue latency when one CTA owns the SM).
//   HS    = the recurrent operand stays on chip: the gate epilogue writes h_t straight into a shared-memory buffer in
//           the UMMA A layout (two 64 KB buffers, ping-pong) and the H-part MMAs of the next step read it from
//           there, instead of h_t -> global -> fence -> TMA -> shared.  The act image is still written (it is the
//           layer's output) but is off the step-to-step critical path.  Single-pass modes only (hi+lo would need
//           2 x 128 KB); meant for layer 0, whose K_in = 16 leaves no input-projection MMAs to hide that round trip.
template <int P, int NSLOT, int NBUF, int KSB, int EPIW = 1, bool HS = false>
struct GruCfg {
  static_assert(!HS || (P == 1 && NSLOT == 1 && NBUF == 2), "HS: single pass, one row tile per CTA, TMEM double-buffered");
  static constexpr uint32_t HBUF = HS ? 4 * CHUNK_BYTES : 0;  // one h_t: 128 rows x 256 units
  static_assert(P == 1 || 8 % (KSB / P) == 0, "a stage must not straddle two 64-K chunks of a hi/lo image");
  static_assert(EPIW == 1 || NSLOT == 1, "EPIW = 2 only with one row tile per CTA");
  static constexpr int KS = KSB / P;  // slabs per stage per part
  static constexpr int THREADS = NSLOT == 2 ? 384 : 64 + 128 * EPIW;
  static constexpr int CTAS_PER_SM = (NSLOT == 1 && NBUF == 1) ? 2 : 1;
  static constexpr int EPI_WARP0 = NSLOT == 2 ? 4 : 2;
  static constexpr uint32_t TMEM_COLS = NSLOT * NBUF * 256;
  static constexpr uint32_t B_PART = KS * G_SLAB;
  static constexpr uint32_t A_PART = KS * A_SLAB;
  static constexpr uint32_t STAGE = P * (B_PART + NSLOT * A_PART);
  static constexpr uint32_t BUDGET = (CTAS_PER_SM == 2 ? 102400 : 208896) - (HS ? 2 * HBUF - 14336 : 0);  // ring bytes per CTA
  static constexpr int STAGES = (int)(BUDGET / STAGE);
  static_assert(STAGES >= 2, "ring too shallow");
  static constexpr uint32_t SMEM = STAGES * STAGE + 2 * 4 * 256 * 4 + 2 * HBUF;
  static_assert(SMEM <= 231424, "shared memory per CTA");
};

//   MC    = launched as clusters of two CTAs (same direction, neighbouring row tiles) that share every weight
//           stage: each CTA fetches HALF of the stage's weight bytes and multicasts them into both CTAs' shared
//           memory (cp.async.bulk ... .multicast::cluster), halving the L2 -> SM weight traffic (weights are
//           6/7 of what this kernel pulls through the crossbar).  A stage is refilled only when BOTH CTAs'
//           MMAs have retired it (multicast tcgen05.commit onto both empty barriers, count 2).
//   PIPE  = software-pipelined gate epilogue: the accumulators are read in 8-unit sub-blocks and the tcgen05.ld of
//           sub-block i+1 is in flight while sub-block i goes through the MUFU / pack / store work (same register
//           footprint as one 16-unit block).  Aimed at layer 0, whose epilogue (128 KB of TMEM reads per chunk at
//           64 B/cycle) is longer than its 17 MMAs.
//   C8    = fp16 main pass + two e4m3 correction MMAs per 32 K elements (see the file header); P = 2 image sizes.
//   Multi-warp producer (NSLOT == 1, no MC / HS): a stage's 2 P bulk copies are issued by 2 P threads in 2 P different warps
//   (warp 0 = weights part 0 + expect_tx; 2 P - 1 extra warps at the end of the block = weights part 1, activation parts).
//   ncu showed the single producer thread -- ~90 dependent address / uniform-datapath instructions per stage, ~660 cycles
//   -- as what every precision mode was waiting for (the MMA thread spun on the `full` barriers while the ring's `empty`
//   barriers were always already free); with one copy per thread and running pointers a stage costs each of them ~25
//   instructions.
template <int P, int NSLOT, bool MC, bool HS>
struct GruProd {
  static constexpr bool MP = NSLOT == 1 && !MC && !HS;
  static constexpr int EXTRA_WARPS = MP ? 2 * P - 1 : 0;
};
// IL: chunk order of one step of an interleaved item: F0 F1 R0 R1 F2 F3 R2 R3 (F = forward recurrence, R = reverse; digit
// = 64-unit block j).  After a recurrence's last chunk of a step (j = 3) the other one issues two chunks of MMAs before
// its next step starts -- that is the time the gate epilogue of j = 3 plus the h_t round trip (act image -> fence -> bulk
// copy) needs, which in layer 0 (K_in = 16: no input-projection MMAs to fill it) was a bubble of a third of every step.
__host__ __device__ constexpr int il_d(int jj) { return (jj >> 1) & 1; }
__host__ __device__ constexpr int il_j(int jj) { return (jj >> 2) * 2 + (jj & 1); }

template <int P, bool F16, int NSLOT, int NBUF, int KSB, int EPIW, bool MC, bool HS = false, bool PIPE = false, bool C8 = false,
          bool IL = false>
__global__ void __launch_bounds__(GruCfg<P, NSLOT, NBUF, KSB, EPIW, HS>::THREADS + 32 * GruProd<P, NSLOT, MC, HS>::EXTRA_WARPS,
                                  GruCfg<P, NSLOT, NBUF, KSB, EPIW, HS>::CTAS_PER_SM)
    tc_gru_layer_kernel(const GruParams p) {
  static_assert(!C8 || (P == 2 && F16 && PIPE && !HS && KSB == 8), "C8: fp16 images, 32 K elements per stage, pipelined epilogue");
  static_assert(NSLOT * NBUF <= 2, "TMEM holds 512 columns");
  static_assert(!MC || NSLOT == 1, "multicast variant: one row tile per CTA");
  static_assert(!(MC && HS), "HS and MC are separate experiments");
  static_assert(!IL || (NSLOT == 1 && NBUF == 2 && !MC && !HS && PIPE), "IL: one row tile per CTA, TMEM double-buffered");
  using C = GruCfg<P, NSLOT, NBUF, KSB, EPIW, HS>;
  using PR = GruProd<P, NSLOT, MC, HS>;
  constexpr int GRU_STAGES = C::STAGES;
  constexpr int GRU_THREADS = C::THREADS + 32 * PR::EXTRA_WARPS;
  constexpr int CORE_WARPS = C::THREADS / 32;  // producer role r > 0 runs in warp CORE_WARPS + r - 1
  constexpr int KS = C::KS;
  constexpr bool FAST = (P == 1);
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bars[2 * GRU_STAGES + 6];
  __shared__ uint32_t tmem_base_s;
  float* bias_s = reinterpret_cast<float*>(smem + GRU_STAGES * C::STAGE);
  // HS: two h_t buffers [kc 0..3][8 slabs x 2048 B] right after the biases (1024-byte aligned: STAGE and 8 KB are)
  const uint32_t hbuf0 = smem_u32(smem) + GRU_STAGES * C::STAGE + 2 * 4 * 256 * 4;
  uint8_t* hbuf_g = smem + GRU_STAGES * C::STAGE + 2 * 4 * 256 * 4;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t full0 = smem_u32(&bars[0]), empty0 = smem_u32(&bars[GRU_STAGES]);
  // tmem_full/empty: one barrier pair per accumulator buffer (8 bytes apart)
  const uint32_t tmem_full = smem_u32(&bars[2 * GRU_STAGES]), tmem_empty = smem_u32(&bars[2 * GRU_STAGES + 2]),
                 h_ready = smem_u32(&bars[2 * GRU_STAGES + 4]);
  if (threadIdx.x == 0) {
    for (int i = 0; i < GRU_STAGES; ++i) {
      mbar_init(full0 + 8 * i, 1);
      mbar_init(empty0 + 8 * i, MC ? 2 : 1);
    }
    for (int i = 0; i < NBUF; ++i) {
      mbar_init(tmem_full + 8 * i, 1);
      mbar_init(tmem_empty + 8 * i, NSLOT * EPIW * 128);
    }
    mbar_init(h_ready, NSLOT * EPIW * 128);
    mbar_init(h_ready + 8, NSLOT * EPIW * 128);  // IL: the reverse recurrence's steps
    fence_barrier_init();
  }
  for (int i = threadIdx.x; i < 2 * 4 * 256; i += GRU_THREADS) bias_s[i] = p.bias[i];
  if (warp == 1) {
    tmem_alloc(smem_u32(&tmem_base_s), C::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (MC) cluster_sync_all();  // the peer's barriers are initialised before anything is multicast at them
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  const uint32_t smem_base = smem_u32(smem);
  const int L = p.L;
  // work items = (group of row tiles, direction).  Plain: one item per CTA at a time, NSLOT tiles each.
  // MC: one item per CLUSTER at a time = two neighbouring row tiles, one per CTA (rank), same direction.
  const uint32_t crank = MC ? cluster_ctarank() : 0;
  // IL: one item = one row tile, BOTH directions, as two recurrences interleaved chunk by chunk (see IL_D / IL_J).
  const int n_items = IL ? p.n_tiles : MC ? (p.n_tiles / 2) * 2 : (p.n_tiles / NSLOT) * 2;
  constexpr int NCH = IL ? 8 : 4;  // unit-chunks per step of an item
  const int item0 = MC ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int item_step = MC ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  constexpr int TILES_PER_ITEM = MC ? 2 : NSLOT;
  const size_t xbytes = (size_t)P * p.kx_slabs * G_SLAB, hbytes = (size_t)P * 32 * G_SLAB;
  const size_t wj_bytes = xbytes + hbytes;

  if (PR::MP && (warp == 0 || warp >= CORE_WARPS)) {
    // ===================== TMA producers: one bulk copy per stage and thread =====================
    if (elect_one()) {
      const int role = warp == 0 ? 0 : warp - CORE_WARPS + 1;  // [0, P): weight part; [P, 2 P): activation part
      const bool is_b = role < P;
      const int pp = is_b ? role : role - P;
      const bool x_short = p.kx_slabs == 2;  // layer 0: one K = 16 stage, x0 image
      const uint64_t pol_last = make_policy_evict_last(), pol_first = make_policy_evict_first();
      uint32_t stage = 0, use = 0, gstep = 0;
      for (int item = item0; item < n_items; item += item_step) {
        const int64_t tile = IL ? item : item >> 1;
        for (int s = 0; s < L; ++s, ++gstep) {
          for (int jj = 0; jj < NCH; ++jj) {
            const int d = IL ? il_d(jj) : (item & 1);
            const int j = IL ? il_j(jj) : jj;
            const int t = d ? (L - 1 - s) : s;
            const int tprev = d ? t + 1 : t - 1;
            // this thread's part of the step's activation operands
            const uint8_t* xa = x_short ? p.xin + ((tile * L + t) * P + pp) * (2 * (size_t)A_SLAB)
                                        : p.xin + (((tile * L + t) * 8) * P + pp) * (size_t)CHUNK_BYTES;
            const uint8_t* ha = (s == 0 ? p.h0img + (((tile * 2 + d) * 4) * P + pp) * (size_t)CHUNK_BYTES
                                        : p.out + (((tile * L + tprev) * 8 + d * 4) * P + pp) * (size_t)CHUNK_BYTES);
            const uint8_t* wj = p.wimg + (size_t)(d * 4 + j) * wj_bytes;
            for (int part = 0; part < 2; ++part) {
              const int total = part == 0 ? p.kx_slabs : 32;
              // running source pointer of this thread's copy, and its bump per stage
              const uint8_t* src;
              size_t bump;
              if (is_b) {
                src = wj + (part ? xbytes : 0) + (size_t)pp * total * G_SLAB;
                bump = (size_t)KS * G_SLAB;
              } else {
                src = part ? ha : xa;
                bump = 0;  // activations: (so >> 3) chunks of P * CHUNK_BYTES, then slabs inside the chunk
              }
              for (int so = 0; so < total; so += KS) {
                const int ns = (total - so) < KS ? (total - so) : KS;
                if (!is_b && part == 1 && so == 0 && j == 0 && gstep > 0) {
                  mbar_wait(h_ready + (IL ? 8 * d : 0), (gstep - 1) & 1);  // h_{t_prev} is in the act image
                  fence_proxy_async_all();
                }
                mbar_wait(empty0 + 8 * stage, (use & 1) ^ 1);
                const uint32_t fb = full0 + 8 * stage;
                const uint32_t sb = smem_base + stage * C::STAGE;
                if (role == 0) mbar_expect_tx(fb, (uint32_t)(P * ns) * (G_SLAB + A_SLAB));
                if (is_b) {
                  // L2 policy experiments (CCSM_TC_L2HINT): 1, 2 = weights evict_last
                  if (p.l2_hint == 1 || p.l2_hint == 2) bulk_g2s_hint(sb + pp * C::B_PART, src, ns * G_SLAB, fb, pol_last);
                  else bulk_g2s(sb + pp * C::B_PART, src, ns * G_SLAB, fb);
                  src += bump;
                } else {
                  const uint8_t* a = (x_short && part == 0) ? src
                                                            : src + (size_t)(so >> 3) * (P * CHUNK_BYTES) + (so & 7) * A_SLAB;
                  // 1 = activations evict_last too; 3 = the last (fourth) read of x_t and every read of h_{t-1} evict_first;
                  // 4 = the last read of x_t and of h_{t-1} evict_first
                  if (p.l2_hint == 1) bulk_g2s_hint(sb + P * C::B_PART + pp * C::A_PART, a, ns * A_SLAB, fb, pol_last);
                  else if ((p.l2_hint == 3 && (j == 3 || part == 1)) || (p.l2_hint == 4 && j == 3))
                    bulk_g2s_hint(sb + P * C::B_PART + pp * C::A_PART, a, ns * A_SLAB, fb, pol_first);
                  else bulk_g2s(sb + P * C::B_PART + pp * C::A_PART, a, ns * A_SLAB, fb);
                }
                if (++stage == GRU_STAGES) {
                  stage = 0;
                  ++use;
                }
              }
            }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      uint32_t stage = 0, use = 0;  // use = how many times the ring wrapped
      uint32_t gstep = 0;
      const uint64_t pol = make_policy_evict_last();
      const bool hint = p.l2_hint != 0;
      for (int item = item0; item < n_items; item += item_step) {
        const int pair = item >> 1, d = item & 1;
        const int64_t tile0 = TILES_PER_ITEM * (int64_t)pair + crank;
        for (int s = 0; s < L; ++s, ++gstep) {
          const int t = d ? (L - 1 - s) : s;
          const int tprev = d ? t + 1 : t - 1;
          for (int j = 0; j < 4; ++j) {
            const uint8_t* wj = p.wimg + (size_t)(d * 4 + j) * wj_bytes;
            for (int part = 0; part < 2; ++part) {
              const int total = part == 0 ? p.kx_slabs : 32;
              for (int so = 0; so < total; so += KS) {
                const int ns = (total - so) < KS ? (total - so) : KS;
                if (!HS && part == 1 && so == 0 && j == 0 && gstep > 0) {
                  mbar_wait(h_ready, (gstep - 1) & 1);  // h_{t_prev} of both slots is in the act image
                  fence_proxy_async_all();
                }
                mbar_wait(empty0 + 8 * stage, (use & 1) ^ 1);
                const uint32_t fb = full0 + 8 * stage;
                const uint32_t sb = smem_base + stage * C::STAGE;
                const bool load_a = !(HS && part == 1);  // HS: the recurrent operand is already in shared memory
                mbar_expect_tx(fb, (uint32_t)(P * ns) * (G_SLAB + (load_a ? NSLOT * A_SLAB : 0)));
                // weights
                const uint8_t* wsrc = wj + (part ? xbytes : 0);
#pragma unroll
                for (int pp = 0; pp < P; ++pp) {
                  if constexpr (MC) {
                    // this CTA's half of the stage's weight slabs, delivered to both CTAs of the cluster
                    const uint32_t half = (uint32_t)(ns / 2) * G_SLAB;
                    bulk_g2s_mc(sb + pp * C::B_PART + crank * half,
                                wsrc + ((size_t)pp * total + so) * G_SLAB + crank * half, half, fb, (uint16_t)3);
                  } else if (hint) {
                    bulk_g2s_hint(sb + pp * C::B_PART, wsrc + ((size_t)pp * total + so) * G_SLAB, ns * G_SLAB, fb, pol);
                  } else {
                    bulk_g2s(sb + pp * C::B_PART, wsrc + ((size_t)pp * total + so) * G_SLAB, ns * G_SLAB, fb);
                  }
                }
                // activations of every slot
#pragma unroll
                for (int sl = 0; sl < NSLOT; ++sl) {
                  if (!load_a) break;
                  const int64_t tile = tile0 + sl;
                  const uint8_t* asrc;
                  size_t part_stride;
                  if (part == 0) {
                    if (p.kx_slabs == 2) {
                      asrc = p.xin + ((tile * L + t) * P) * (2 * (size_t)A_SLAB);
                      part_stride = 2 * A_SLAB;
                    } else {
                      asrc = p.xin + (((tile * L + t) * 8 + (so >> 3)) * P) * (size_t)CHUNK_BYTES + (so & 7) * A_SLAB;
                      part_stride = CHUNK_BYTES;
                    }
                  } else {
                    if (s == 0)
                      asrc = p.h0img + (((tile * 2 + d) * 4 + (so >> 3)) * P) * (size_t)CHUNK_BYTES + (so & 7) * A_SLAB;
                    else
                      asrc = p.out + (((tile * L + tprev) * 8 + d * 4 + (so >> 3)) * P) * (size_t)CHUNK_BYTES +
                             (so & 7) * A_SLAB;
                    part_stride = CHUNK_BYTES;
                  }
#pragma unroll
                  for (int pp = 0; pp < P; ++pp) {
                    const uint32_t dsta = sb + P * C::B_PART + (sl * P + pp) * C::A_PART;
                    if (hint) bulk_g2s_hint(dsta, asrc + pp * part_stride, ns * A_SLAB, fb, pol);
                    else bulk_g2s(dsta, asrc + pp * part_stride, ns * A_SLAB, fb);
                  }
                }
                if (++stage == GRU_STAGES) {
                  stage = 0;
                  ++use;
                }
              }
            }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (elect_one()) {
      constexpr uint32_t idesc192 = make_idesc(128, 192, F16);
      constexpr uint32_t idesc192_8 = make_idesc_e4m3(128, 192);
      uint32_t stage = 0, use = 0, chunk = 0;
      uint32_t hcount = 0;  // HS: completions of h_ready consumed so far (one per step: item start, then every step)
      for (int item = item0; item < n_items; item += item_step) {
        for (int s = 0; s < L; ++s, ++hcount) {
          for (int j = 0; j < NCH; ++j, ++chunk) {
            const uint32_t buf = chunk % NBUF, bphase = (chunk / NBUF) & 1;
            // completion #u of tmem_empty[buf]: #0 = initial bias arming, #k = drain + re-arm after use k-1
            mbar_wait(tmem_empty + 8 * buf, bphase);
            tc_fence_after();
            for (int part = 0; part < 2; ++part) {
              const int total = part == 0 ? p.kx_slabs : 32;
              for (int so = 0; so < total; so += KS) {
                const int ns = (total - so) < KS ? (total - so) : KS;
                if (HS && part == 1 && so == 0 && j == 0) mbar_wait(h_ready, hcount & 1);  // h_{t_prev} is in hbuf
                mbar_wait(full0 + 8 * stage, use & 1);
                tc_fence_after();
                const uint32_t sb = smem_base + stage * C::STAGE;
                if (C8 && ns == 4) {
                  // 32 K elements: two fp16 MMAs (K = 16) on the hi parts, then a8 . Wlo8 and alo8 . W8 (K = 32 each)
#pragma unroll
                  for (int sl = 0; sl < NSLOT; ++sl) {
                    const uint32_t dcol = tmem + (NBUF == 2 ? buf : sl) * 256 + (part == 0 ? 0 : 64);
                    const uint32_t a0 = sb + P * C::B_PART + (sl * P) * C::A_PART, b0 = sb;
#pragma unroll
                    for (int ks = 0; ks < 2; ++ks)
                      umma_f16(dcol, make_smem_desc(a0 + ks * 2 * A_SLAB, A_SLAB, 128),
                               make_smem_desc(b0 + ks * 2 * G_SLAB, G_SLAB, 128), idesc192, 1u);
#pragma unroll
                    for (int c = 0; c < 2; ++c)
                      umma_f8(dcol, make_smem_desc(a0 + C::A_PART + c * 2 * A_SLAB, A_SLAB, 128),
                              make_smem_desc(b0 + C::B_PART + c * 2 * G_SLAB, G_SLAB, 128), idesc192_8, 1u);
                  }
                } else
                for (int ks = 0; ks < ns / 2; ++ks) {
#pragma unroll
                  for (int sl = 0; sl < NSLOT; ++sl) {
                    const uint32_t dcol = tmem + (NBUF == 2 ? buf : sl) * 256;
#pragma unroll
                    for (int pass = 0; pass < (P == 2 ? 3 : 1); ++pass) {
                      const int pa = pass == 2 ? 1 : 0, pb = pass == 1 ? 1 : 0;
                      const uint32_t a_addr = (HS && part == 1)
                                                  ? hbuf0 + (hcount & 1) * C::HBUF + (so + ks * 2) * A_SLAB
                                                  : sb + P * C::B_PART + (sl * P + pa) * C::A_PART + ks * 2 * A_SLAB;
                      const uint32_t b_addr = sb + pb * C::B_PART + ks * 2 * G_SLAB;
                      const uint64_t ad = make_smem_desc(a_addr, A_SLAB, 128);
                      // X part -> columns [0,192) = (n_i, r, z); H part -> columns [64,256) = (r, z, n_h).
                      // The epilogue pre-loaded every column with its gate bias, so all MMAs accumulate.
                      umma_f16(dcol + (part == 0 ? 0 : 64), ad, make_smem_desc(b_addr, G_SLAB, 128), idesc192, 1u);
                    }
                  }
                }
                // frees the smem stage when these MMAs retire (MC: in both CTAs -- the peer refills half of it)
                if constexpr (MC) umma_commit_mc(empty0 + 8 * stage, (uint16_t)3);
                else umma_commit(empty0 + 8 * stage);
                if (++stage == GRU_STAGES) {
                  stage = 0;
                  ++use;
                }
              }
            }
            umma_commit(tmem_full + 8 * buf);  // accumulators of this unit-chunk complete
          }
        }
      }
    }
    __syncwarp();
  } else if (warp >= C::EPI_WARP0) {
    // ===================== gate epilogue =====================
    // tcgen05.ld lane rule: a warp may only touch TMEM lanes [32 * (warp % 4), +32)
    constexpr int NUB = 4 / EPIW;                                       // 16-unit blocks per warp
    const int slot = EPIW >= 2 ? 0 : (warp - C::EPI_WARP0) >> 2, quad = warp & 3;
    const int ub0 = EPIW >= 2 ? NUB * ((warp - C::EPI_WARP0) >> 2) : 0;  // first 16-unit block of this warp
    const int row = quad * 32 + lane;
    const uint32_t trow0 = tmem + ((uint32_t)(quad * 32) << 16) + slot * 256;
    uint32_t chunk = 0;
    // gridDim.x is even (host), so every item of this CTA has the same direction: the biases to arm are fixed
    // (IL: the direction changes from chunk to chunk; the first two chunks are the forward recurrence's.)
    const float* bz0 = bias_s + (IL ? 0 : (item0 & 1)) * 4 * 256;
    for (int b = 0; b < NBUF; ++b) {  // arm the first NBUF unit-chunks (j = b)
#pragma unroll
      for (int k = 0; k < NUB; ++k)
        arm_bias16(trow0 + (NBUF == 2 ? b * 256 : 0), ub0 + k, bz0, b * 64 + (ub0 + k) * 16);
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(tmem_empty + 8 * b);
    }
    uint32_t hcnt = 0;  // HS: steps started so far (selects the h buffer: read (hcnt & 1), write the other one)
    for (int item = item0; item < n_items; item += item_step) {
      const int pair = item >> 1, d_item = item & 1;
      const int64_t tile = IL ? (int64_t)item : TILES_PER_ITEM * (int64_t)pair + crank + slot;
      if constexpr (HS) {
        const int d = d_item;
        // h0 of this row -> the buffer step 0 reads.  Safe to overwrite: every MMA that read this buffer (step L-2 of
        // the previous item) retired before the tmem_full of the previous item's last step, which this thread passed.
        const uint8_t* src = p.h0img + ((tile * 2 + d) * 4) * (size_t)CHUNK_BYTES + row * 16;
        uint8_t* dst = hbuf_g + (hcnt & 1) * C::HBUF + row * 16;
        constexpr int NSL = 32 / EPIW;  // K-slabs this warp copies
        const int sl0 = EPIW == 2 ? ((warp - C::EPI_WARP0) >> 2) * NSL : 0;
#pragma unroll 8
        for (int sl = sl0; sl < sl0 + NSL; ++sl)
          *reinterpret_cast<uint4*>(dst + sl * A_SLAB) = __ldcg(reinterpret_cast<const uint4*>(src + sl * A_SLAB));
        fence_proxy_async_smem();
        mbar_arrive(h_ready);
      }
      for (int s = 0; s < L; ++s, ++hcnt) {
        for (int jj = 0; jj < NCH; ++jj, ++chunk) {
          const int d = IL ? il_d(jj) : d_item;
          const int j = IL ? il_j(jj) : jj;
          const int t = d ? (L - 1 - s) : s;
          const int tprev = d ? t + 1 : t - 1;
          // biases of the unit-chunk that uses this accumulator buffer next (IL: two chunks further on in the interleaved order)
          const float* bz = IL ? bias_s + il_d((jj + 2) & 7) * 4 * 256 : bz0;
          const int jnext = IL ? il_j((jj + 2) & 7) : ((j + NBUF) & 3);
          const uint8_t* hp_base =
              HS ? hbuf_g + (hcnt & 1) * C::HBUF + j * (size_t)CHUNK_BYTES
                 : ((s == 0) ? p.h0img + (((tile * 2 + d) * 4 + j) * P) * (size_t)CHUNK_BYTES
                             : p.out + (((tile * L + tprev) * 8 + d * 4 + j) * P) * (size_t)CHUNK_BYTES);
          uint8_t* out_base = p.out + (((tile * L + t) * 8 + d * 4 + j) * P) * (size_t)CHUNK_BYTES;
          uint8_t* hnext = hbuf_g + ((hcnt + 1) & 1) * C::HBUF + j * (size_t)CHUNK_BYTES;
          // prefetch h_{t_prev} for this row's 64 units (L2 latency overlaps the MMAs of this chunk)
          uint4 hph[2 * NUB], hpl[2 * NUB];
#pragma unroll
          for (int q = 0; q < 2 * NUB; ++q) {
            if constexpr (HS)
              hph[q] = *reinterpret_cast<const uint4*>(hp_base + (2 * ub0 + q) * A_SLAB + row * 16);
            else
              hph[q] = __ldcg(reinterpret_cast<const uint4*>(hp_base + (2 * ub0 + q) * A_SLAB + row * 16));
            if constexpr (C8) {
              const uint2 t8 = __ldcg(reinterpret_cast<const uint2*>(hp_base + CHUNK_BYTES + 4096 + c8_off(2 * ub0 + q) + row * 16));
              hpl[q] = make_uint4(t8.x, t8.y, 0, 0);
            } else if constexpr (P == 2)
              hpl[q] = __ldcg(reinterpret_cast<const uint4*>(hp_base + CHUNK_BYTES + (2 * ub0 + q) * A_SLAB + row * 16));
            else
              hpl[q] = make_uint4(0, 0, 0, 0);
          }
          const uint32_t buf = chunk % NBUF, bphase = (chunk / NBUF) & 1;
          const uint32_t trow = trow0 + (NBUF == 2 ? buf * 256 : 0);
          mbar_wait(tmem_full + 8 * buf, bphase);
          tc_fence_after();
          if constexpr (PIPE) {
            constexpr int NSB = 2 * NUB;  // 8-unit sub-blocks of this warp
            uint32_t acc[2][4][8];        // [ping-pong][n_i, r, z, n_h][8 units]
            uint2 a8_even = make_uint2(0, 0), l8_even = make_uint2(0, 0);  // C8: first half of a 16-element e4m3 row
            const int c00 = ub0 * 16;
#pragma unroll
            for (int g = 0; g < 4; ++g) tmem_ld8(trow + g * 64 + c00, acc[0][g]);
#pragma unroll
            for (int sb = 0; sb < NSB; ++sb) {
              const int col = c00 + sb * 8;
              tmem_ld_wait();  // sub-block sb has landed (issued one iteration ago)
              if (sb + 1 < NSB) {
#pragma unroll
                for (int g = 0; g < 4; ++g) tmem_ld8(trow + g * 64 + col + 8, acc[(sb + 1) & 1][g]);
              }
              // re-arm these columns with the biases of the unit-chunk that uses this buffer next
              arm_bias8(trow, col, bz, jnext * 64 + col);
              float hp[8], hn[8];
              if constexpr (C8) join8_c8(hph[sb], make_uint2(hpl[sb].x, hpl[sb].y), hp);
              else join8<P, F16>(hph[sb], hpl[sb], hp);
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                float r, z;
                if constexpr (C8) {
                  sigmoid2_s(__uint_as_float(acc[sb & 1][1][i]), __uint_as_float(acc[sb & 1][2][i]), r, z);
                } else {
                  r = sig_<FAST, C8>(__uint_as_float(acc[sb & 1][1][i]));
                  z = sig_<FAST, C8>(__uint_as_float(acc[sb & 1][2][i]));
                }
                const float n = tnh_<FAST, C8>(fmaf(r, __uint_as_float(acc[sb & 1][3][i]), __uint_as_float(acc[sb & 1][0][i])));
                hn[i] = fmaf(z, hp[i] - n, n);  // (1 - z) * n + z * h
              }
              const int slab = (col >> 3);  // K-slab of these 8 units inside the 64-unit chunk
              if constexpr (C8) {
                uint4 hi;
                uint2 a8, l8;
                split8_c8(hn, hi, a8, l8);
                *reinterpret_cast<uint4*>(out_base + slab * A_SLAB + row * 16) = hi;
                if ((sb & 1) == 0) {
                  a8_even = a8;
                  l8_even = l8;
                } else {  // both halves of a 16-element e4m3 row: one 16-byte store each
                  uint8_t* b8 = out_base + CHUNK_BYTES + c8_off(slab - 1) + row * 16;
                  *reinterpret_cast<uint4*>(b8) = make_uint4(a8_even.x, a8_even.y, a8.x, a8.y);
                  *reinterpret_cast<uint4*>(b8 + 4096) = make_uint4(l8_even.x, l8_even.y, l8.x, l8.y);
                }
              } else {
              uint4 hi, lo;
              split8<P, F16>(hn, hi, lo);
              if constexpr (HS) {
                if (s + 1 < L) *reinterpret_cast<uint4*>(hnext + slab * A_SLAB + row * 16) = hi;
              }
              *reinterpret_cast<uint4*>(out_base + slab * A_SLAB + row * 16) = hi;
              if constexpr (P == 2) *reinterpret_cast<uint4*>(out_base + CHUNK_BYTES + slab * A_SLAB + row * 16) = lo;
              }
            }
          } else {
#pragma unroll
          for (int k = 0; k < NUB; ++k) {
            const int ub = ub0 + k;
            uint32_t ani[16], ar[16], az[16], anh[16];
            tmem_ld16(trow + 0 + ub * 16, ani);
            tmem_ld16(trow + 64 + ub * 16, ar);
            tmem_ld16(trow + 128 + ub * 16, az);
            tmem_ld16(trow + 192 + ub * 16, anh);
            tmem_ld_wait();
            // re-arm these columns with the biases of the unit-chunk that uses this buffer next
            arm_bias16(trow, ub, bz, jnext * 64 + ub * 16);
#pragma unroll
            for (int q = 0; q < 2; ++q) {
              float hp[8], hn[8];
              join8<P, F16>(hph[k * 2 + q], hpl[k * 2 + q], hp);
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const int c = q * 8 + i;
                const float r = sigmoid_<FAST>(__uint_as_float(ar[c]));
                const float z = sigmoid_<FAST>(__uint_as_float(az[c]));
                const float n = tanh_<FAST>(fmaf(r, __uint_as_float(anh[c]), __uint_as_float(ani[c])));
                hn[i] = fmaf(z, hp[i] - n, n);  // (1 - z) * n + z * h
              }
              uint4 hi, lo;
              split8<P, F16>(hn, hi, lo);
              if constexpr (HS) {
                // next step's A operand, written in place in the UMMA layout (the last step's output feeds nobody)
                if (s + 1 < L) *reinterpret_cast<uint4*>(hnext + (ub * 2 + q) * A_SLAB + row * 16) = hi;
              }
              *reinterpret_cast<uint4*>(out_base + (ub * 2 + q) * A_SLAB + row * 16) = hi;
              if constexpr (P == 2)
                *reinterpret_cast<uint4*>(out_base + CHUNK_BYTES + (ub * 2 + q) * A_SLAB + row * 16) = lo;
            }
          }
          }
          tmem_st_wait();
          tc_fence_before();
          mbar_arrive(tmem_empty + 8 * buf);
          if (j == 3) {
            if constexpr (HS) {
              if (s + 1 < L) {
                fence_proxy_async_smem();  // generic-proxy shared-memory writes -> visible to the MMA's operand reads
                mbar_arrive(h_ready);
              }
            } else {
              fence_proxy_async_all();  // generic-proxy global writes -> visible to the producer's bulk copies
              mbar_arrive(h_ready + (IL ? 8 * d : 0));
            }
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (MC) cluster_sync_all();  // the peer may still be arriving on this CTA's empty barriers
  if (warp == 1) tmem_dealloc(tmem, C::TMEM_COLS);
}

// ------------------------------------------------------------------------------------------------
// fp16c8 GRU layer kernel with ON-CHIP operand conversion ("CV").
//
// The layer kernels are bound by the bytes each SM pulls in through its L2 port (~64 B/clk/SM: profiles/r02_bound.md), and
// half of the fp16c8 correction operands are redundant on the wire: e4m3(a) and e4m3(W) are just roundings of the fp16
// hi parts that are loaded anyway.  This kernel loads, per 32 K elements, only hi (fp16) and the scaled residuals
// (alo8, Wlo8) -- 30 KB per stage instead of 40 KB -- and four converter warps produce a8 = e4m3(hi_a) and
// W8 = e4m3(hi_W / S) in shared memory, right where the e4m3 MMAs expect them.  The fp16 MMAs of a stage issue as soon as
// it lands; its two e4m3 MMAs issue one stage later, when the conversion is done, so the converters are off the
// critical path.  Work item, TMEM plan, bias arming and the pipelined epilogue are those of tc_gru_layer_kernel with
// (NSLOT 1, NBUF 2): one row tile per CTA, one direction per CTA, two accumulator buffers.
//   warp 0 producer | warp 1 MMA issuer | 4 * EPIW epilogue warps | 4 converter warps
// ------------------------------------------------------------------------------------------------
template <int EPIW>
struct CvCfg {
  static constexpr int KS = 4;                          // 8-element K-slabs per stage = 32 K elements
  static constexpr uint32_t B_PART = KS * G_SLAB;       // 12288: fp16 hi | [Wlo8 6144][W8 6144]
  static constexpr uint32_t A_PART = KS * A_SLAB;       // 8192:  fp16 hi | [a8 4096][alo8 4096]
  static constexpr uint32_t STAGE = 2 * (B_PART + A_PART);  // 40960
  static constexpr int STAGES = 5;
  static constexpr uint32_t SMEM = STAGES * STAGE + 2 * 4 * 256 * 4;
  static constexpr int EPI_WARP0 = 2;
  static constexpr int CV_WARP0 = 2 + 4 * EPIW;
  static constexpr int THREADS = 32 * (CV_WARP0 + 4);
};

// 16 fp16 (two 16-byte slab rows) -> 16 e4m3 (one 16-byte row), optionally scaled by 2^-12 first (exact in fp16)
template <bool SCALE>
__device__ __forceinline__ uint4 cvt16_e4m3(const uint4& a, const uint4& b) {
  const uint32_t in[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
  uint32_t out[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    uint32_t r[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      __half2 v = *reinterpret_cast<const __half2*>(&in[2 * i + h]);
      if constexpr (SCALE) v = __hmul2(v, __float2half2_rn(C8_INV_S));
      r[h] = __nv_cvt_halfraw2_to_fp8x2(*reinterpret_cast<__half2_raw*>(&v), __NV_SATFINITE, __NV_E4M3);
    }
    out[i] = r[0] | (r[1] << 16);
  }
  return make_uint4(out[0], out[1], out[2], out[3]);
}

template <int EPIW>
__global__ void __launch_bounds__(CvCfg<EPIW>::THREADS, 1) tc_gru_cv_kernel(const GruParams p) {
  using C = CvCfg<EPIW>;
  constexpr int S = C::STAGES, KS = C::KS, P = 2;
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bars[3 * S + 5];
  __shared__ uint32_t tmem_base_s;
  float* bias_s = reinterpret_cast<float*>(smem + S * C::STAGE);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t full0 = smem_u32(&bars[0]), empty0 = smem_u32(&bars[S]), conv0 = smem_u32(&bars[2 * S]);
  const uint32_t tmem_full = smem_u32(&bars[3 * S]), tmem_empty = smem_u32(&bars[3 * S + 2]),
                 h_ready = smem_u32(&bars[3 * S + 4]);
  if (threadIdx.x == 0) {
    for (int i = 0; i < S; ++i) {
      mbar_init(full0 + 8 * i, 1);
      mbar_init(empty0 + 8 * i, 1);
      mbar_init(conv0 + 8 * i, 4);  // one arrival per converter warp
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(tmem_full + 8 * i, 1);
      mbar_init(tmem_empty + 8 * i, EPIW * 128);
    }
    mbar_init(h_ready, EPIW * 128);
    fence_barrier_init();
  }
  for (int i = threadIdx.x; i < 2 * 4 * 256; i += C::THREADS) bias_s[i] = p.bias[i];
  if (warp == 1) {
    tmem_alloc(smem_u32(&tmem_base_s), 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  const uint32_t smem_base = smem_u32(smem);
  const int L = p.L;
  const int n_items = p.n_tiles * 2;   // (row tile, direction); the grid is even, so a CTA keeps one direction
  const int item0 = (int)blockIdx.x, item_step = (int)gridDim.x;
  const size_t xbytes = (size_t)P * p.kx_slabs * G_SLAB, hbytes = (size_t)P * 32 * G_SLAB;
  const size_t wj_bytes = xbytes + hbytes;
  const bool x_short = p.kx_slabs < KS;  // layer 0: one 16-K stage in the 3-pass fp16 layout, nothing to convert

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      uint32_t stage = 0, use = 0, gstep = 0;
      for (int item = item0; item < n_items; item += item_step) {
        const int64_t tile = item >> 1;
        const int d = item & 1;
        for (int s = 0; s < L; ++s, ++gstep) {
          const int t = d ? (L - 1 - s) : s;
          const int tprev = d ? t + 1 : t - 1;
          for (int j = 0; j < 4; ++j) {
            const uint8_t* wj = p.wimg + (size_t)(d * 4 + j) * wj_bytes;
            for (int part = 0; part < 2; ++part) {
              const int total = part == 0 ? p.kx_slabs : 32;
              const uint8_t* wsrc = wj + (part ? xbytes : 0);
              for (int so = 0; so < total; so += KS) {
                const int ns = (total - so) < KS ? (total - so) : KS;
                if (part == 1 && so == 0 && j == 0 && gstep > 0) {
                  mbar_wait(h_ready, (gstep - 1) & 1);  // h_{t_prev} is in the act image
                  fence_proxy_async_all();
                }
                mbar_wait(empty0 + 8 * stage, (use & 1) ^ 1);
                const uint32_t fb = full0 + 8 * stage;
                const uint32_t sb = smem_base + stage * C::STAGE;
                const uint8_t* asrc;
                size_t part_stride;
                if (part == 0) {
                  if (x_short) {
                    asrc = p.xin + ((tile * L + t) * P) * (2 * (size_t)A_SLAB);
                    part_stride = 2 * A_SLAB;
                  } else {
                    asrc = p.xin + (((tile * L + t) * 8 + (so >> 3)) * P) * (size_t)CHUNK_BYTES + (so & 7) * A_SLAB;
                    part_stride = CHUNK_BYTES;
                  }
                } else {
                  if (s == 0)
                    asrc = p.h0img + (((tile * 2 + d) * 4 + (so >> 3)) * P) * (size_t)CHUNK_BYTES + (so & 7) * A_SLAB;
                  else
                    asrc = p.out + (((tile * L + tprev) * 8 + d * 4 + (so >> 3)) * P) * (size_t)CHUNK_BYTES + (so & 7) * A_SLAB;
                  part_stride = CHUNK_BYTES;
                }
                if (ns == KS) {
                  // hi parts whole; of the e4m3 parts only the scaled residuals: Wlo8 = first half of the weight group,
                  // alo8 = second half of the activation group.  e4m3(W) / e4m3(a) are produced on chip.
                  mbar_expect_tx(fb, C::B_PART + C::B_PART / 2 + C::A_PART + C::A_PART / 2);
                  bulk_g2s(sb, wsrc + (size_t)so * G_SLAB, C::B_PART, fb);
                  bulk_g2s(sb + C::B_PART, wsrc + ((size_t)total + so) * G_SLAB, C::B_PART / 2, fb);
                  bulk_g2s(sb + 2 * C::B_PART, asrc, C::A_PART, fb);
                  bulk_g2s(sb + 2 * C::B_PART + C::A_PART + C::A_PART / 2, asrc + part_stride + C::A_PART / 2, C::A_PART / 2, fb);
                } else {
                  mbar_expect_tx(fb, (uint32_t)(P * ns) * (G_SLAB + A_SLAB));
#pragma unroll
                  for (int pp = 0; pp < P; ++pp) {
                    bulk_g2s(sb + pp * C::B_PART, wsrc + ((size_t)pp * total + so) * G_SLAB, ns * G_SLAB, fb);
                    bulk_g2s(sb + 2 * C::B_PART + pp * C::A_PART, asrc + pp * part_stride, ns * A_SLAB, fb);
                  }
                }
                if (++stage == S) {
                  stage = 0;
                  ++use;
                }
              }
            }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (elect_one()) {
      constexpr uint32_t idesc = make_idesc(128, 192, true);
      constexpr uint32_t idesc8 = make_idesc_e4m3(128, 192);
      uint32_t stage = 0, use = 0, chunk = 0;
      // the stage whose e4m3 MMAs are still owed (issued once its conversion is done, one stage later)
      bool owe = false;
      uint32_t owe_stage = 0, owe_use = 0, owe_dcol = 0;
      auto settle = [&]() {
        if (!owe) return;
        mbar_wait(conv0 + 8 * owe_stage, owe_use & 1);
        tc_fence_after();
        const uint32_t sb = smem_base + owe_stage * C::STAGE;
        const uint32_t a1 = sb + 2 * C::B_PART + C::A_PART, b1 = sb + C::B_PART;
        // a8 . Wlo8, then alo8 . W8 (K = 32 each)
        umma_f8(owe_dcol, make_smem_desc(a1, A_SLAB, 128), make_smem_desc(b1, G_SLAB, 128), idesc8, 1u);
        umma_f8(owe_dcol, make_smem_desc(a1 + 2 * A_SLAB, A_SLAB, 128), make_smem_desc(b1 + 2 * G_SLAB, G_SLAB, 128), idesc8, 1u);
        umma_commit(empty0 + 8 * owe_stage);
        owe = false;
      };
      for (int item = item0; item < n_items; item += item_step) {
        for (int s = 0; s < L; ++s) {
          for (int j = 0; j < 4; ++j, ++chunk) {
            const uint32_t buf = chunk & 1, bphase = (chunk >> 1) & 1;
            mbar_wait(tmem_empty + 8 * buf, bphase);
            tc_fence_after();
            for (int part = 0; part < 2; ++part) {
              const int total = part == 0 ? p.kx_slabs : 32;
              const uint32_t dcol = tmem + buf * 256 + (part == 0 ? 0 : 64);  // X -> (n_i, r, z); H -> (r, z, n_h)
              for (int so = 0; so < total; so += KS) {
                const int ns = (total - so) < KS ? (total - so) : KS;
                mbar_wait(full0 + 8 * stage, use & 1);
                tc_fence_after();
                const uint32_t sb = smem_base + stage * C::STAGE;
                const uint32_t a0 = sb + 2 * C::B_PART, b0 = sb;
                if (ns == KS) {
#pragma unroll
                  for (int q = 0; q < 2; ++q)
                    umma_f16(dcol, make_smem_desc(a0 + q * 2 * A_SLAB, A_SLAB, 128),
                             make_smem_desc(b0 + q * 2 * G_SLAB, G_SLAB, 128), idesc, 1u);
                  settle();
                  owe = true;
                  owe_stage = stage;
                  owe_use = use;
                  owe_dcol = dcol;
                } else {
                  settle();
                  for (int ks = 0; ks < ns / 2; ++ks) {
#pragma unroll
                    for (int pass = 0; pass < 3; ++pass) {
                      const int pa = pass == 2 ? 1 : 0, pb = pass == 1 ? 1 : 0;
                      umma_f16(dcol, make_smem_desc(a0 + pa * C::A_PART + ks * 2 * A_SLAB, A_SLAB, 128),
                               make_smem_desc(b0 + pb * C::B_PART + ks * 2 * G_SLAB, G_SLAB, 128), idesc, 1u);
                    }
                  }
                  umma_commit(empty0 + 8 * stage);
                }
                if (++stage == S) {
                  stage = 0;
                  ++use;
                }
              }
            }
            settle();
            umma_commit(tmem_full + 8 * buf);  // accumulators of this unit-chunk complete
          }
        }
      }
    }
    __syncwarp();
  } else if (warp >= C::CV_WARP0) {
    // ===================== converters: a8 = e4m3(a_hi), W8 = e4m3(W_hi / S) of every full stage =====================
    const int ct = (warp - C::CV_WARP0) * 32 + lane;  // 0..127
    uint32_t stage = 0, use = 0;
    const int per_chunk_x = (p.kx_slabs + KS - 1) / KS;
    for (int item = item0; item < n_items; item += item_step)
      for (int c = 0; c < L * 4; ++c)
        for (int st = 0; st < per_chunk_x + 8; ++st) {
          const bool full_stage = !(x_short && st < per_chunk_x);
          if (full_stage) {
            mbar_wait(full0 + 8 * stage, use & 1);
            uint8_t* sb = smem + stage * C::STAGE;
            // activations: 128 rows x 2 groups of 16 K elements
#pragma unroll
            for (int q = 0; q < 2; ++q) {
              const int task = ct + q * 128, row = task & 127, k16 = task >> 7;
              const uint8_t* src = sb + 2 * C::B_PART + (2 * k16) * A_SLAB + row * 16;
              *reinterpret_cast<uint4*>(sb + 2 * C::B_PART + C::A_PART + k16 * A_SLAB + row * 16) =
                  cvt16_e4m3<false>(*reinterpret_cast<const uint4*>(src), *reinterpret_cast<const uint4*>(src + A_SLAB));
            }
            // weights: 192 rows x 2 groups
#pragma unroll
            for (int q = 0; q < 3; ++q) {
              const int task = ct + q * 128, row = task % 192, k16 = task / 192;
              const uint8_t* src = sb + (2 * k16) * G_SLAB + row * 16;
              *reinterpret_cast<uint4*>(sb + C::B_PART + (2 + k16) * G_SLAB + row * 16) =
                  cvt16_e4m3<true>(*reinterpret_cast<const uint4*>(src), *reinterpret_cast<const uint4*>(src + G_SLAB));
            }
            fence_proxy_async_smem();  // generic-proxy shared-memory writes -> visible to the MMA's operand reads
          }
          // every use of a stage completes one phase of its conv barrier (short stages too: the parity follows `use`)
          __syncwarp();
          if (lane == 0) mbar_arrive(conv0 + 8 * stage);
          if (++stage == S) {
            stage = 0;
            ++use;
          }
        }
  } else {
    // ===================== gate epilogue (pipelined, one thread per row and half-chunk) =====================
    const int quad = warp & 3;  // tcgen05.ld lane rule: a warp touches TMEM lanes [32 * (warp % 4), +32)
    const int ub0 = EPIW == 2 ? 2 * ((warp - C::EPI_WARP0) >> 2) : 0;  // first 16-unit block of this warp
    constexpr int NUB = 4 / EPIW, NSB = 2 * NUB;
    const int row = quad * 32 + lane;
    const uint32_t trow0 = tmem + ((uint32_t)(quad * 32) << 16);
    uint32_t chunk = 0;
    const float* bz = bias_s + (item0 & 1) * 4 * 256;
    for (int b = 0; b < 2; ++b) {  // arm the first two unit-chunks (j = b)
#pragma unroll
      for (int k = 0; k < NUB; ++k) arm_bias16(trow0 + b * 256, ub0 + k, bz, b * 64 + (ub0 + k) * 16);
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(tmem_empty + 8 * b);
    }
    for (int item = item0; item < n_items; item += item_step) {
      const int64_t tile = item >> 1;
      const int d = item & 1;
      for (int s = 0; s < L; ++s) {
        const int t = d ? (L - 1 - s) : s;
        const int tprev = d ? t + 1 : t - 1;
        for (int j = 0; j < 4; ++j, ++chunk) {
          const uint8_t* hp_base = (s == 0) ? p.h0img + (((tile * 2 + d) * 4 + j) * P) * (size_t)CHUNK_BYTES
                                            : p.out + (((tile * L + tprev) * 8 + d * 4 + j) * P) * (size_t)CHUNK_BYTES;
          uint8_t* out_base = p.out + (((tile * L + t) * 8 + d * 4 + j) * P) * (size_t)CHUNK_BYTES;
          uint4 hph[NSB];
          uint2 hpl[NSB];
#pragma unroll
          for (int q = 0; q < NSB; ++q) {
            hph[q] = __ldcg(reinterpret_cast<const uint4*>(hp_base + (2 * ub0 + q) * A_SLAB + row * 16));
            hpl[q] = __ldcg(reinterpret_cast<const uint2*>(hp_base + CHUNK_BYTES + 4096 + c8_off(2 * ub0 + q) + row * 16));
          }
          const uint32_t buf = chunk & 1, bphase = (chunk >> 1) & 1;
          const uint32_t trow = trow0 + buf * 256;
          mbar_wait(tmem_full + 8 * buf, bphase);
          tc_fence_after();
          uint32_t acc[2][4][8];
          uint2 a8_even = make_uint2(0, 0), l8_even = make_uint2(0, 0);
          const int c00 = ub0 * 16;
#pragma unroll
          for (int g = 0; g < 4; ++g) tmem_ld8(trow + g * 64 + c00, acc[0][g]);
#pragma unroll
          for (int sb = 0; sb < NSB; ++sb) {
            const int col = c00 + sb * 8;
            tmem_ld_wait();
            if (sb + 1 < NSB) {
#pragma unroll
              for (int g = 0; g < 4; ++g) tmem_ld8(trow + g * 64 + col + 8, acc[(sb + 1) & 1][g]);
            }
            arm_bias8(trow, col, bz, ((j + 2) & 3) * 64 + col);
            float hp[8], hn[8];
            join8_c8(hph[sb], hpl[sb], hp);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float r = sigmoid_s(__uint_as_float(acc[sb & 1][1][i]));
              const float z = sigmoid_s(__uint_as_float(acc[sb & 1][2][i]));
              const float n = tanh_s(fmaf(r, __uint_as_float(acc[sb & 1][3][i]), __uint_as_float(acc[sb & 1][0][i])));
              hn[i] = fmaf(z, hp[i] - n, n);
            }
            uint4 hi;
            uint2 a8, l8;
            split8_c8(hn, hi, a8, l8);
            const int slab = col >> 3;
            *reinterpret_cast<uint4*>(out_base + slab * A_SLAB + row * 16) = hi;
            if ((sb & 1) == 0) {
              a8_even = a8;
              l8_even = l8;
            } else {
              // e4m3(h) is still stored for the kernels that load it (attention, the non-converting variants)
              uint8_t* b8 = out_base + CHUNK_BYTES + c8_off(slab - 1) + row * 16;
              *reinterpret_cast<uint4*>(b8) = make_uint4(a8_even.x, a8_even.y, a8.x, a8.y);
              *reinterpret_cast<uint4*>(b8 + 4096) = make_uint4(l8_even.x, l8_even.y, l8.x, l8.y);
            }
          }
          tmem_st_wait();
          tc_fence_before();
          mbar_arrive(tmem_empty + 8 * buf);
          if (j == 3) {
            fence_proxy_async_all();
            mbar_arrive(h_ready);
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 512);
}

// ------------------------------------------------------------------------------------------------
// CTA-pair GRU layer kernel (tcgen05.mma.cta_group::2): a cluster of two CTAs (one TPC) runs M = 256:
// CTA c owns row tile 2*pair + c (its 128 rows of A in its own shared memory, its 128 TMEM lanes) and
// HALF of every weight tile (96 of the 192 gate rows), so each weight byte is fetched from L2 and read
// from shared memory once per 256 rows.  TMEM holds two accumulator buffers per CTA: the MMAs of
// unit-chunk j+1 overlap the gate epilogue of j.
//   warp 0      producer (both CTAs): own A K-slabs + own half of the weights -> 7-stage ring
//   warp 1      rank 0: MMA issuer for the pair; rank 1: relay (local full barrier -> leader's peer_full)
//   warps 2-5   gate epilogue of the CTA's own row tile; zero the n_h columns with tcgen05.st after
//               draining a buffer so that every H-part MMA can accumulate
// Weight image: [dir][j]{X: [half][part][K_in/8 slabs x 1536 B], H: [half][part][32 slabs x 1536 B]}.
// ------------------------------------------------------------------------------------------------
constexpr uint32_t GH_SLAB = 1536;  // 96 gate rows x 16 B
constexpr int PAIR_THREADS = 192;

// NBUF = 2: one cluster per TPC, TMEM double-buffered, 7 stages.
// NBUF = 1: two clusters per TPC (two CTAs per SM, 256 TMEM columns each), 3 stages each: one cluster's gate
//           epilogue and step-boundary latency overlap the other cluster's MMAs.
template <int P, int NBUF>
struct PairCfg {
  static constexpr int KS = 8 / P;
  static constexpr int STAGES = NBUF == 2 ? 7 : 3;
  static constexpr int CTAS_PER_SM = NBUF == 2 ? 1 : 2;
  static constexpr uint32_t TMEM_COLS = NBUF * 256;
  static constexpr uint32_t B_PART = KS * GH_SLAB;
  static constexpr uint32_t A_PART = KS * A_SLAB;
  static constexpr uint32_t STAGE = P * (B_PART + A_PART);  // 28672
  static constexpr uint32_t SMEM = STAGES * STAGE + 2 * 4 * 256 * 4;
};

template <int P, bool F16, int NBUF>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(PAIR_THREADS, PairCfg<P, NBUF>::CTAS_PER_SM)
    tc_gru_pair_kernel(const GruParams p) {
  using C = PairCfg<P, NBUF>;
  constexpr int KS = C::KS;
  constexpr bool FAST = (P == 1);
  constexpr int S = C::STAGES;
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bars[3 * S + 5];
  __shared__ uint32_t tmem_base_s;
  float* bias_s = reinterpret_cast<float*>(smem + S * C::STAGE);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const uint32_t full0 = smem_u32(&bars[0]), empty0 = smem_u32(&bars[S]), peer0 = smem_u32(&bars[2 * S]);
  const uint32_t tmem_full = smem_u32(&bars[3 * S]), tmem_empty = smem_u32(&bars[3 * S + 2]),
                 h_ready = smem_u32(&bars[3 * S + 4]);
  if (threadIdx.x == 0) {
    for (int i = 0; i < S; ++i) {
      mbar_init(full0 + 8 * i, 1);
      mbar_init(empty0 + 8 * i, 1);
      mbar_init(peer0 + 8 * i, 1);
    }
    for (int i = 0; i < NBUF; ++i) {
      mbar_init(tmem_full + 8 * i, 1);
      mbar_init(tmem_empty + 8 * i, 8);  // one arrival per epilogue warp of both CTAs
    }
    mbar_init(h_ready, 4);
    fence_barrier_init();
  }
  for (int i = threadIdx.x; i < 2 * 4 * 256; i += PAIR_THREADS) bias_s[i] = p.bias[i];
  if (warp == 1) {
    tmem_alloc2(smem_u32(&tmem_base_s), C::TMEM_COLS);
    tmem_relinquish2();
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  const uint32_t smem_base = smem_u32(smem);
  const int L = p.L;
  const int n_items = (p.n_tiles / 2) * 2;  // pairs of row tiles x 2 directions
  const int cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;
  const size_t xbytes = (size_t)2 * P * p.kx_slabs * GH_SLAB, hbytes = (size_t)2 * P * 32 * GH_SLAB;
  const size_t wj_bytes = xbytes + hbytes;

  if (warp == 0) {
    // ===================== TMA producer (each CTA: own A tile, own half of B) =====================
    if (elect_one()) {
      uint32_t stage = 0, use = 0, gstep = 0;
      for (int item = cluster_id; item < n_items; item += n_clusters) {
        const int pair = item >> 1, d = item & 1;
        const int64_t tile = 2 * (int64_t)pair + rank;
        for (int s = 0; s < L; ++s, ++gstep) {
          const int t = d ? (L - 1 - s) : s;
          const int tprev = d ? t + 1 : t - 1;
          for (int j = 0; j < 4; ++j) {
            const uint8_t* wj = p.wimg + (size_t)(d * 4 + j) * wj_bytes;
            for (int part = 0; part < 2; ++part) {
              const int total = part == 0 ? p.kx_slabs : 32;
              const uint8_t* wsrc = wj + (part ? xbytes : 0) + (size_t)rank * P * total * GH_SLAB;
              for (int so = 0; so < total; so += KS) {
                const int ns = (total - so) < KS ? (total - so) : KS;
                if (part == 1 && so == 0 && j == 0 && gstep > 0) {
                  mbar_wait(h_ready, (gstep - 1) & 1);
                  fence_proxy_async_all();
                }
                mbar_wait(empty0 + 8 * stage, (use & 1) ^ 1);
                const uint32_t fb = full0 + 8 * stage;
                const uint32_t sb = smem_base + stage * C::STAGE;
                mbar_expect_tx(fb, (uint32_t)(P * ns) * (GH_SLAB + A_SLAB));
                const uint8_t* asrc;
                size_t part_stride;
                if (part == 0) {
                  if (p.kx_slabs == 2) {
                    asrc = p.xin + ((tile * L + t) * P) * (2 * (size_t)A_SLAB);
                    part_stride = 2 * A_SLAB;
                  } else {
                    asrc = p.xin + (((tile * L + t) * 8 + (so >> 3)) * P) * (size_t)CHUNK_BYTES + (so & 7) * A_SLAB;
                    part_stride = CHUNK_BYTES;
                  }
                } else {
                  if (s == 0)
                    asrc = p.h0img + (((tile * 2 + d) * 4 + (so >> 3)) * P) * (size_t)CHUNK_BYTES + (so & 7) * A_SLAB;
                  else
                    asrc = p.out + (((tile * L + tprev) * 8 + d * 4 + (so >> 3)) * P) * (size_t)CHUNK_BYTES +
                           (so & 7) * A_SLAB;
                  part_stride = CHUNK_BYTES;
                }
#pragma unroll
                for (int pp = 0; pp < P; ++pp) {
                  bulk_g2s(sb + pp * C::B_PART, wsrc + ((size_t)pp * total + so) * GH_SLAB, ns * GH_SLAB, fb);
                  bulk_g2s(sb + P * C::B_PART + pp * C::A_PART, asrc + pp * part_stride, ns * A_SLAB, fb);
                }
                if (++stage == S) {
                  stage = 0;
                  ++use;
                }
              }
            }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (elect_one()) {
      if (rank == 1) {
        // ===================== relay: my operands landed -> tell the leader =====================
        uint32_t stage = 0, use = 0;
        for (int item = cluster_id; item < n_items; item += n_clusters)
          for (int s = 0; s < L; ++s)
            for (int j = 0; j < 4; ++j)
              for (int part = 0; part < 2; ++part) {
                const int total = part == 0 ? p.kx_slabs : 32;
                for (int so = 0; so < total; so += KS) {
                  mbar_wait(full0 + 8 * stage, use & 1);
                  mbar_arrive_remote(mapa_u32(peer0 + 8 * stage, 0));
                  if (++stage == S) {
                    stage = 0;
                    ++use;
                  }
                }
              }
      } else {
        // ===================== MMA issuer for the pair =====================
        constexpr uint32_t idesc = make_idesc(256, 192, F16);
        uint32_t stage = 0, use = 0, chunk = 0;
        for (int item = cluster_id; item < n_items; item += n_clusters) {
          for (int s = 0; s < L; ++s) {
            for (int j = 0; j < 4; ++j, ++chunk) {
              const uint32_t buf = chunk % NBUF, u = chunk / NBUF;
              mbar_wait(tmem_empty + 8 * buf, u & 1);  // completion #u: #0 = initial arming, #k = drain of use k-1
              tc_fence_after();
              const uint32_t dcol = tmem + buf * 256;
              for (int part = 0; part < 2; ++part) {
                const int total = part == 0 ? p.kx_slabs : 32;
                for (int so = 0; so < total; so += KS) {
                  const int ns = (total - so) < KS ? (total - so) : KS;
                  mbar_wait(full0 + 8 * stage, use & 1);
                  mbar_wait(peer0 + 8 * stage, use & 1);
                  tc_fence_after();
                  const uint32_t sb = smem_base + stage * C::STAGE;
                  for (int ks = 0; ks < ns / 2; ++ks) {
#pragma unroll
                    for (int pass = 0; pass < (P == 2 ? 3 : 1); ++pass) {
                      const int pa = pass == 2 ? 1 : 0, pb = pass == 1 ? 1 : 0;
                      const uint64_t ad = make_smem_desc(sb + P * C::B_PART + pa * C::A_PART + ks * 2 * A_SLAB, A_SLAB, 128);
                      const uint64_t bd = make_smem_desc(sb + pb * C::B_PART + ks * 2 * GH_SLAB, GH_SLAB, 128);
                      umma_f16_pair(dcol + (part == 0 ? 0 : 64), ad, bd, idesc, 1u);  // columns pre-loaded with biases
                    }
                  }
                  umma_commit_pair(empty0 + 8 * stage, 0x3);
                  if (++stage == S) {
                    stage = 0;
                    ++use;
                  }
                }
              }
              umma_commit_pair(tmem_full + 8 * buf, 0x3);
            }
          }
        }
      }
    }
    __syncwarp();
  } else {
    // ===================== gate epilogue (warps 2-5) of this CTA's row tile =====================
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    const uint32_t trow0 = tmem + ((uint32_t)(quad * 32) << 16);
    const uint32_t remote_empty = mapa_u32(tmem_empty, 0);
    // every item of this cluster has the same direction (the cluster count is even)
    const float* bz = bias_s + (cluster_id & 1) * 4 * 256;
    // initial state: every buffer armed with the biases of its first unit-chunk
#pragma unroll
    for (int b = 0; b < NBUF; ++b) {
#pragma unroll
      for (int ub = 0; ub < 4; ++ub) arm_bias16(trow0 + b * 256, ub, bz, b * 64 + ub * 16);
    }
    tmem_st_wait();
    tc_fence_before();
    __syncwarp();
    if (lane == 0) {
#pragma unroll
      for (int b = 0; b < NBUF; ++b) mbar_arrive_remote(remote_empty + 8 * b);
    }
    uint32_t chunk = 0;
    for (int item = cluster_id; item < n_items; item += n_clusters) {
      const int pair = item >> 1, d = item & 1;
      const int64_t tile = 2 * (int64_t)pair + rank;
      for (int s = 0; s < L; ++s) {
        const int t = d ? (L - 1 - s) : s;
        const int tprev = d ? t + 1 : t - 1;
        for (int j = 0; j < 4; ++j, ++chunk) {
          const uint8_t* hp_base =
              (s == 0) ? p.h0img + (((tile * 2 + d) * 4 + j) * P) * (size_t)CHUNK_BYTES
                       : p.out + (((tile * L + tprev) * 8 + d * 4 + j) * P) * (size_t)CHUNK_BYTES;
          uint8_t* out_base = p.out + (((tile * L + t) * 8 + d * 4 + j) * P) * (size_t)CHUNK_BYTES;
          uint4 hph[8], hpl[8];
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            hph[q] = __ldcg(reinterpret_cast<const uint4*>(hp_base + q * A_SLAB + row * 16));
            if constexpr (P == 2)
              hpl[q] = __ldcg(reinterpret_cast<const uint4*>(hp_base + CHUNK_BYTES + q * A_SLAB + row * 16));
            else
              hpl[q] = make_uint4(0, 0, 0, 0);
          }
          const uint32_t buf = chunk % NBUF, u = chunk / NBUF;
          const uint32_t trow = trow0 + buf * 256;
          mbar_wait(tmem_full + 8 * buf, u & 1);
          tc_fence_after();
#pragma unroll
          for (int ub = 0; ub < 4; ++ub) {
            uint32_t ani[16], ar[16], az[16], anh[16];
            tmem_ld16(trow + 0 + ub * 16, ani);
            tmem_ld16(trow + 64 + ub * 16, ar);
            tmem_ld16(trow + 128 + ub * 16, az);
            tmem_ld16(trow + 192 + ub * 16, anh);
            tmem_ld_wait();
            arm_bias16(trow, ub, bz, ((j + NBUF) & 3) * 64 + ub * 16);  // biases of the next user of this buffer
#pragma unroll
            for (int q = 0; q < 2; ++q) {
              float hp[8], hn[8];
              join8<P, F16>(hph[ub * 2 + q], hpl[ub * 2 + q], hp);
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const int c = q * 8 + i;
                const float r = sigmoid_<FAST>(__uint_as_float(ar[c]));
                const float z = sigmoid_<FAST>(__uint_as_float(az[c]));
                const float n = tanh_<FAST>(fmaf(r, __uint_as_float(anh[c]), __uint_as_float(ani[c])));
                hn[i] = fmaf(z, hp[i] - n, n);
              }
              uint4 hi, lo;
              split8<P, F16>(hn, hi, lo);
              *reinterpret_cast<uint4*>(out_base + (ub * 2 + q) * A_SLAB + row * 16) = hi;
              if constexpr (P == 2)
                *reinterpret_cast<uint4*>(out_base + CHUNK_BYTES + (ub * 2 + q) * A_SLAB + row * 16) = lo;
            }
          }
          tmem_st_wait();
          tc_fence_before();
          if (j == 3) fence_proxy_async_all();
          __syncwarp();
          if (lane == 0) {
            mbar_arrive_remote(remote_empty + 8 * buf);
            if (j == 3) mbar_arrive(h_ready);
          }
        }
      }
    }
  }
  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) tmem_dealloc2(tmem, C::TMEM_COLS);
}

// ------------------------------------------------------------------------------------------------
// "Duo" GRU layer kernel: CTA pair (tcgen05.mma.cta_group::2, M = 256) running BOTH directions of its two row tiles
// as two interleaved recurrences.
//
// Why: the layer kernels above are bound by the bytes each SM pulls in through its L2 port (~64 B/clk/SM: bf16,
// fp16x3 and fp16c8 all land on ~16 cycles per KB of bulk-copy traffic, profiles/r02_*), not by the tensor pipe.
// In a CTA pair each SM fetches only HALF of every weight slab (96 of the 192 gate rows; the pair's MMA reads the
// other half from the peer's shared memory), which removes 30 % of the bytes per row tile.  And with the forward
// and the reverse scan of the same tiles alternating chunk by chunk -- recurrence 0 = forward on TMEM buffer 0,
// recurrence 1 = reverse on buffer 1 -- one recurrence's MMAs run while the other's gate epilogue drains and while
// its h_t makes the round trip through L2 at a step boundary (what left layer 0 latency-bound).
//
//   warp 0      producer (both CTAs): own A K-slabs + own half of the weight slabs -> 7-stage ring (28 KB stages)
//   warp 1      rank 0: MMA issuer for the pair; rank 1: relay (local full barrier -> leader's peer_full)
//   warps 2-5   gate epilogue of recurrence 0 (forward), warps 6-9 of recurrence 1 (reverse); thread = row
// Work item = one pair of row tiles (CTA c owns tile 2 * pair + c); cluster i takes items i, i + n_clusters, ...
// Weight image: the CTA-pair layout [dir][j]{X: [half][part][slabs x 1536 B], H: [half][part][32 slabs x 1536 B]}.
// ------------------------------------------------------------------------------------------------
constexpr int DUO_THREADS = 320;
constexpr int DUO_STAGES = 7;

template <int P>
struct DuoCfg {
  static constexpr int KS = 8 / P;  // K-slabs (8 elements) per stage and part
  static constexpr uint32_t B_PART = KS * GH_SLAB;
  static constexpr uint32_t A_PART = KS * A_SLAB;
  static constexpr uint32_t STAGE = P * (B_PART + A_PART);  // 28672
  static constexpr uint32_t SMEM = DUO_STAGES * STAGE + 2 * 4 * 256 * 4;
};

template <int P, bool F16, bool C8>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(DUO_THREADS, 1) tc_gru_duo_kernel(const GruParams p) {
  static_assert(!C8 || (P == 2 && F16), "C8: fp16 images");
  using C = DuoCfg<P>;
  constexpr int KS = C::KS;
  constexpr bool FAST = (P == 1);
  constexpr int S = DUO_STAGES;
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bars[3 * S + 6];
  __shared__ uint32_t tmem_base_s;
  float* bias_s = reinterpret_cast<float*>(smem + S * C::STAGE);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const uint32_t full0 = smem_u32(&bars[0]), empty0 = smem_u32(&bars[S]), peer0 = smem_u32(&bars[2 * S]);
  // per recurrence r: tmem_full + 8r, tmem_empty + 8r, h_ready + 8r
  const uint32_t tmem_full = smem_u32(&bars[3 * S]), tmem_empty = smem_u32(&bars[3 * S + 2]),
                 h_ready = smem_u32(&bars[3 * S + 4]);
  if (threadIdx.x == 0) {
    for (int i = 0; i < S; ++i) {
      mbar_init(full0 + 8 * i, 1);
      mbar_init(empty0 + 8 * i, 1);
      mbar_init(peer0 + 8 * i, 1);
    }
    for (int r = 0; r < 2; ++r) {
      mbar_init(tmem_full + 8 * r, 1);
      mbar_init(tmem_empty + 8 * r, 8);  // one arrival per epilogue warp of the recurrence, both CTAs
      mbar_init(h_ready + 8 * r, 4);     // the recurrence's four epilogue warps of this CTA
    }
    fence_barrier_init();
  }
  for (int i = threadIdx.x; i < 2 * 4 * 256; i += DUO_THREADS) bias_s[i] = p.bias[i];
  if (warp == 1) {
    tmem_alloc2(smem_u32(&tmem_base_s), 512);
    tmem_relinquish2();
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  const uint32_t smem_base = smem_u32(smem);
  const int L = p.L;
  const int n_items = p.n_tiles / 2;  // pairs of row tiles
  const int cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;
  const size_t xbytes = (size_t)2 * P * p.kx_slabs * GH_SLAB, hbytes = (size_t)2 * P * 32 * GH_SLAB;
  const size_t wj_bytes = xbytes + hbytes;
  // the K = 16 input of layer 0 is a single short stage in the 3-pass layout (also in the C8 mode)
  const bool x_short = p.kx_slabs < KS;

  if (warp == 0) {
    // ===================== TMA producer (each CTA: own A tile, own half of B) =====================
    if (elect_one()) {
      uint32_t stage = 0, use = 0, gstep = 0;
      for (int item = cluster_id; item < n_items; item += n_clusters) {
        const int64_t tile = 2 * (int64_t)item + rank;
        for (int s = 0; s < L; ++s, ++gstep) {
          for (int j = 0; j < 4; ++j) {
            for (int d = 0; d < 2; ++d) {  // recurrence d: forward / reverse scan
              const int t = d ? (L - 1 - s) : s;
              const int tprev = d ? t + 1 : t - 1;
              const uint8_t* wj = p.wimg + (size_t)(d * 4 + j) * wj_bytes;
              for (int part = 0; part < 2; ++part) {
                const int total = part == 0 ? p.kx_slabs : 32;
                const uint8_t* wsrc = wj + (part ? xbytes : 0) + (size_t)rank * P * total * GH_SLAB;
                for (int so = 0; so < total; so += KS) {
                  const int ns = (total - so) < KS ? (total - so) : KS;
                  if (part == 1 && so == 0 && j == 0 && gstep > 0) {
                    mbar_wait(h_ready + 8 * d, (gstep - 1) & 1);  // h_{t_prev} of this recurrence is in the act image
                    fence_proxy_async_all();
                  }
                  mbar_wait(empty0 + 8 * stage, (use & 1) ^ 1);
                  const uint32_t fb = full0 + 8 * stage;
                  const uint32_t sb = smem_base + stage * C::STAGE;
                  mbar_expect_tx(fb, (uint32_t)(P * ns) * (GH_SLAB + A_SLAB));
                  const uint8_t* asrc;
                  size_t part_stride;
                  if (part == 0) {
                    if (x_short) {
                      asrc = p.xin + ((tile * L + t) * P) * (2 * (size_t)A_SLAB);
                      part_stride = 2 * A_SLAB;
                    } else {
                      asrc = p.xin + (((tile * L + t) * 8 + (so >> 3)) * P) * (size_t)CHUNK_BYTES + (so & 7) * A_SLAB;
                      part_stride = CHUNK_BYTES;
                    }
                  } else {
                    if (s == 0)
                      asrc = p.h0img + (((tile * 2 + d) * 4 + (so >> 3)) * P) * (size_t)CHUNK_BYTES + (so & 7) * A_SLAB;
                    else
                      asrc = p.out + (((tile * L + tprev) * 8 + d * 4 + (so >> 3)) * P) * (size_t)CHUNK_BYTES +
                             (so & 7) * A_SLAB;
                    part_stride = CHUNK_BYTES;
                  }
#pragma unroll
                  for (int pp = 0; pp < P; ++pp) {
                    bulk_g2s(sb + pp * C::B_PART, wsrc + ((size_t)pp * total + so) * GH_SLAB, ns * GH_SLAB, fb);
                    bulk_g2s(sb + P * C::B_PART + pp * C::A_PART, asrc + pp * part_stride, ns * A_SLAB, fb);
                  }
                  if (++stage == S) {
                    stage = 0;
                    ++use;
                  }
                }
              }
            }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (elect_one()) {
      if (rank == 1) {
        // ===================== relay: my operands landed -> tell the leader =====================
        uint32_t stage = 0, use = 0;
        const int per_chunk = (p.kx_slabs + KS - 1) / KS + 32 / KS;
        for (int item = cluster_id; item < n_items; item += n_clusters)
          for (int c = 0; c < L * 8 * per_chunk; ++c) {
            mbar_wait(full0 + 8 * stage, use & 1);
            mbar_arrive_remote(mapa_u32(peer0 + 8 * stage, 0));
            if (++stage == S) {
              stage = 0;
              ++use;
            }
          }
      } else {
        // ===================== MMA issuer for the pair =====================
        constexpr uint32_t idesc = make_idesc(256, 192, F16);
        constexpr uint32_t idesc8 = make_idesc_e4m3(256, 192);
        uint32_t stage = 0, use = 0, chunk = 0;
        for (int item = cluster_id; item < n_items; item += n_clusters) {
          for (int s = 0; s < L; ++s) {
            for (int j = 0; j < 4; ++j, ++chunk) {
              for (int d = 0; d < 2; ++d) {
                // completion #chunk of tmem_empty[d]: #0 = initial arming, #k = drain + re-arm after chunk k-1
                mbar_wait(tmem_empty + 8 * d, chunk & 1);
                tc_fence_after();
                const uint32_t dcol = tmem + d * 256;
                for (int part = 0; part < 2; ++part) {
                  const int total = part == 0 ? p.kx_slabs : 32;
                  const uint32_t dpart = dcol + (part == 0 ? 0 : 64);  // X -> (n_i, r, z); H -> (r, z, n_h)
                  for (int so = 0; so < total; so += KS) {
                    const int ns = (total - so) < KS ? (total - so) : KS;
                    mbar_wait(full0 + 8 * stage, use & 1);
                    mbar_wait(peer0 + 8 * stage, use & 1);
                    tc_fence_after();
                    const uint32_t sb = smem_base + stage * C::STAGE;
                    const uint32_t a0 = sb + P * C::B_PART, b0 = sb;
                    if (C8 && ns == KS) {
                      // 32 K elements: two fp16 MMAs on the hi parts, then a8 . Wlo8 and alo8 . W8 (K = 32 each)
#pragma unroll
                      for (int q = 0; q < 2; ++q) {
                        umma_f16_pair(dpart, make_smem_desc(a0 + q * 2 * A_SLAB, A_SLAB, 128),
                                      make_smem_desc(b0 + q * 2 * GH_SLAB, GH_SLAB, 128), idesc, 1u);
                        umma_f8_pair(dpart, make_smem_desc(a0 + C::A_PART + q * 2 * A_SLAB, A_SLAB, 128),
                                     make_smem_desc(b0 + C::B_PART + q * 2 * GH_SLAB, GH_SLAB, 128), idesc8, 1u);
                      }
                    } else {
                      for (int ks = 0; ks < ns / 2; ++ks) {
#pragma unroll
                        for (int pass = 0; pass < (P == 2 ? 3 : 1); ++pass) {
                          const int pa = pass == 2 ? 1 : 0, pb = pass == 1 ? 1 : 0;
                          umma_f16_pair(dpart, make_smem_desc(a0 + pa * C::A_PART + ks * 2 * A_SLAB, A_SLAB, 128),
                                        make_smem_desc(b0 + pb * C::B_PART + ks * 2 * GH_SLAB, GH_SLAB, 128), idesc, 1u);
                        }
                      }
                    }
                    umma_commit_pair(empty0 + 8 * stage, 0x3);
                    if (++stage == S) {
                      stage = 0;
                      ++use;
                    }
                  }
                }
                umma_commit_pair(tmem_full + 8 * d, 0x3);
              }
            }
          }
        }
      }
    }
    __syncwarp();
  } else {
    // ===================== gate epilogue: warps 2-5 recurrence 0 (forward), warps 6-9 recurrence 1 (reverse) =====
    const int d = (warp - 2) >> 2;
    const int quad = warp & 3;  // tcgen05.ld lane rule: a warp touches TMEM lanes [32 * (warp % 4), +32)
    const int row = quad * 32 + lane;
    const uint32_t trow = tmem + ((uint32_t)(quad * 32) << 16) + d * 256;
    const uint32_t remote_empty = mapa_u32(tmem_empty + 8 * d, 0);
    const float* bz = bias_s + d * 4 * 256;
#pragma unroll
    for (int ub = 0; ub < 4; ++ub) arm_bias16(trow, ub, bz, ub * 16);  // unit-chunk 0
    tmem_st_wait();
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive_remote(remote_empty);
    uint32_t chunk = 0;
    for (int item = cluster_id; item < n_items; item += n_clusters) {
      const int64_t tile = 2 * (int64_t)item + rank;
      for (int s = 0; s < L; ++s) {
        const int t = d ? (L - 1 - s) : s;
        const int tprev = d ? t + 1 : t - 1;
        for (int j = 0; j < 4; ++j, ++chunk) {
          const uint8_t* hp_base =
              (s == 0) ? p.h0img + (((tile * 2 + d) * 4 + j) * P) * (size_t)CHUNK_BYTES
                       : p.out + (((tile * L + tprev) * 8 + d * 4 + j) * P) * (size_t)CHUNK_BYTES;
          uint8_t* out_base = p.out + (((tile * L + t) * 8 + d * 4 + j) * P) * (size_t)CHUNK_BYTES;
          // prefetch h_{t_prev} of this row's 64 units (the L2 latency overlaps the chunk's MMAs)
          uint4 hph[8], hpl[8];
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            hph[q] = __ldcg(reinterpret_cast<const uint4*>(hp_base + q * A_SLAB + row * 16));
            if constexpr (C8) {
              const uint2 t8 = __ldcg(reinterpret_cast<const uint2*>(hp_base + CHUNK_BYTES + 4096 + c8_off(q) + row * 16));
              hpl[q] = make_uint4(t8.x, t8.y, 0, 0);
            } else if constexpr (P == 2)
              hpl[q] = __ldcg(reinterpret_cast<const uint4*>(hp_base + CHUNK_BYTES + q * A_SLAB + row * 16));
            else
              hpl[q] = make_uint4(0, 0, 0, 0);
          }
          mbar_wait(tmem_full + 8 * d, chunk & 1);
          tc_fence_after();
          uint32_t acc[2][4][8];  // [ping-pong][n_i, r, z, n_h][8 units]
          uint2 a8_even = make_uint2(0, 0), l8_even = make_uint2(0, 0);
#pragma unroll
          for (int g = 0; g < 4; ++g) tmem_ld8(trow + g * 64, acc[0][g]);
#pragma unroll
          for (int sb = 0; sb < 8; ++sb) {
            const int col = sb * 8;
            tmem_ld_wait();  // sub-block sb has landed (issued one iteration ago)
            if (sb + 1 < 8) {
#pragma unroll
              for (int g = 0; g < 4; ++g) tmem_ld8(trow + g * 64 + col + 8, acc[(sb + 1) & 1][g]);
            }
            arm_bias8(trow, col, bz, ((j + 1) & 3) * 64 + col);  // biases of this recurrence's next unit-chunk
            float hp[8], hn[8];
            if constexpr (C8) join8_c8(hph[sb], make_uint2(hpl[sb].x, hpl[sb].y), hp);
            else join8<P, F16>(hph[sb], hpl[sb], hp);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float r = sig_<FAST, C8>(__uint_as_float(acc[sb & 1][1][i]));
              const float z = sig_<FAST, C8>(__uint_as_float(acc[sb & 1][2][i]));
              const float n = tnh_<FAST, C8>(fmaf(r, __uint_as_float(acc[sb & 1][3][i]), __uint_as_float(acc[sb & 1][0][i])));
              hn[i] = fmaf(z, hp[i] - n, n);  // (1 - z) * n + z * h
            }
            if constexpr (C8) {
              uint4 hi;
              uint2 a8, l8;
              split8_c8(hn, hi, a8, l8);
              *reinterpret_cast<uint4*>(out_base + sb * A_SLAB + row * 16) = hi;
              if ((sb & 1) == 0) {
                a8_even = a8;
                l8_even = l8;
              } else {
                uint8_t* b8 = out_base + CHUNK_BYTES + c8_off(sb - 1) + row * 16;
                *reinterpret_cast<uint4*>(b8) = make_uint4(a8_even.x, a8_even.y, a8.x, a8.y);
                *reinterpret_cast<uint4*>(b8 + 4096) = make_uint4(l8_even.x, l8_even.y, l8.x, l8.y);
              }
            } else {
              uint4 hi, lo;
              split8<P, F16>(hn, hi, lo);
              *reinterpret_cast<uint4*>(out_base + sb * A_SLAB + row * 16) = hi;
              if constexpr (P == 2) *reinterpret_cast<uint4*>(out_base + CHUNK_BYTES + sb * A_SLAB + row * 16) = lo;
            }
          }
          tmem_st_wait();
          tc_fence_before();
          if (j == 3) fence_proxy_async_all();  // generic-proxy global writes -> visible to the producer's bulk copies
          __syncwarp();
          if (lane == 0) {
            mbar_arrive_remote(remote_empty);
            if (j == 3) mbar_arrive(h_ready + 8 * d);
          }
        }
      }
    }
  }
  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) tmem_dealloc2(tmem, 512);
}

// ------------------------------------------------------------------------------------------------
// CTA-pair GRU layer kernel, second form ("pair2", variant n): tcgen05.mma.cta_group::2 (M = 256: CTA c owns row tile
// 2 * pair + c, its 128 rows of A in its own shared memory, and HALF of every weight slab -- 96 of the 192 gate rows),
// so every weight byte crosses L2 -> SM once per 256 rows: 688 KB instead of 983 KB per row tile and unit-chunk in fp16c8.
// Against round 1's pair kernel: operands arrive by 2-D tensor-map loads (cp.async.bulk.tensor ... .cta_group::2) that
// complete on the LEADER CTA's `full` barrier from both CTAs -- no relay warp, no relay hop on the critical path (plain
// bulk copies can only signal a barrier of the destination CTA: tc_selftest.cu, flag 2 vs flag 4) --, one load per
// producer thread in 2 P warps per CTA, the software-pipelined gate epilogue, and the fp16c8 operand forms.
//   warp 0 + 2 P - 1 warps at the end   producers (weights part pp / activations part pp of this CTA)
//   warp 1                              rank 0: MMA issuer for the pair (rank 1: idle)
//   warps 2-5                           gate epilogue of this CTA's row tile
// Work item = (pair of row tiles, direction); the cluster count is even, so a cluster keeps one direction.
// Layers >= 1 only (K_in = 512: every stage is full).  Weight image: the CTA-pair layout of tc_gru_pair_kernel.
// ------------------------------------------------------------------------------------------------
template <int P, int MODE>
struct Pair2Cfg {
  static constexpr int KS = 8 / P;
  static constexpr int STAGES = 7;
  static constexpr uint32_t B_PART = KS * GH_SLAB;
  static constexpr uint32_t A_PART = KS * A_SLAB;
  static constexpr uint32_t STAGE = P * (B_PART + A_PART);  // 28672
  static constexpr uint32_t SMEM = STAGES * STAGE + 2 * 4 * 256 * 4;
  static constexpr int EPI_WARPS = MODE ? 8 : 4;
  static constexpr int CORE_WARPS = 2 + EPI_WARPS;
  static constexpr int THREADS = 32 * (CORE_WARPS + 2 * P - 1);
};

// MODE 0: work item = (pair of row tiles, direction), the two TMEM buffers alternate chunk by chunk (layers >= 1).
// MODE 1: work item = pair of row tiles, BOTH directions as two recurrences that alternate chunk by chunk (F0 R0 F1 R1
//         ...), recurrence d on TMEM buffer d with its own four epilogue warps.
// MODE 2: both directions in the order of tc_gru_layer_kernel's IL form (F0 F1 R0 R1 F2 F3 R2 R3), the two TMEM buffers
//         alternating chunk by chunk, eight epilogue warps that each take half of a chunk's units.
// Modes 1 / 2 are meant for layer 0, where nothing else hides the h_t round trip at a step boundary; its K = 16 input part is
// one short stage (two slabs, 3-pass form).
template <int P, bool F16, bool C8, int MODE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(Pair2Cfg<P, MODE>::THREADS, 1)
    tc_gru_pair2_kernel(const GruParams p, const __grid_constant__ CUtensorMap tm_w, const __grid_constant__ CUtensorMap tm_x,
                        const __grid_constant__ CUtensorMap tm_o, const __grid_constant__ CUtensorMap tm_h0,
                        const __grid_constant__ CUtensorMap tm_wx, const __grid_constant__ CUtensorMap tm_xs) {
  static_assert(!C8 || (P == 2 && F16), "C8: fp16 images");
  using C = Pair2Cfg<P, MODE>;
  constexpr bool DUO = MODE != 0;  // both directions in one item
  constexpr int NCH = DUO ? 8 : 4;  // unit-chunks per step of an item
  constexpr int KS = C::KS;
  constexpr bool FAST = (P == 1);
  constexpr int S = C::STAGES;
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bars[2 * S + 6];
  __shared__ uint32_t tmem_base_s;
  float* bias_s = reinterpret_cast<float*>(smem + S * C::STAGE);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const uint32_t full0 = smem_u32(&bars[0]), empty0 = smem_u32(&bars[S]);
  const uint32_t tmem_full = smem_u32(&bars[2 * S]), tmem_empty = smem_u32(&bars[2 * S + 2]),
                 h_ready = smem_u32(&bars[2 * S + 4]);
  if (threadIdx.x == 0) {
    for (int i = 0; i < S; ++i) {
      mbar_init(full0 + 8 * i, 1);   // used in the leader: its expect_tx arrival + the bytes of both CTAs
      mbar_init(empty0 + 8 * i, 1);  // multicast commit of the stage's MMAs
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(tmem_full + 8 * i, 1);
      mbar_init(tmem_empty + 8 * i, MODE == 2 ? 16 : 8);  // one arrival per epilogue warp working on the buffer, both CTAs
      mbar_init(h_ready + 8 * i, MODE == 2 ? 8 : 4);      // DUO: one per recurrence
    }
    fence_barrier_init();
  }
  for (int i = threadIdx.x; i < 2 * 4 * 256; i += C::THREADS) bias_s[i] = p.bias[i];
  if (warp == 1) {
    tmem_alloc2(smem_u32(&tmem_base_s), 512);
    tmem_relinquish2();
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  const uint32_t smem_base = smem_u32(smem);
  const int L = p.L;
  const int n_items = DUO ? p.n_tiles / 2 : (p.n_tiles / 2) * 2;  // pairs of row tiles (x 2 directions)
  const int cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;
  // weight image in 1536-byte slabs: [dir][j]{X: [half][part][kx], H: [half][part][32]}
  const int kx = p.kx_slabs;
  const bool x_short = kx < KS;  // layer 0: K = 16, one stage of two slabs
  const int XS = 2 * P * kx, WJ = XS + 2 * P * 32;

  if (warp == 0 || warp >= C::CORE_WARPS) {
    // ===================== producers: one tensor-map load per stage and thread =====================
    if (elect_one()) {
      const int role = warp == 0 ? 0 : warp - C::CORE_WARPS + 1;  // [0, P): weight part; [P, 2 P): activation part
      const bool is_b = role < P;
      const int pp = is_b ? role : role - P;
      const uint32_t leader_full0 = mapa_u32(full0, 0);
      const uint64_t pol_first = make_policy_evict_first();
      tma_prefetch_desc(is_b ? (const void*)&tm_w : (const void*)&tm_x);
      tma_prefetch_desc(is_b ? (const void*)&tm_wx : (const void*)&tm_xs);
      if (!is_b) {
        tma_prefetch_desc(&tm_o);
        tma_prefetch_desc(&tm_h0);
      }
      uint32_t stage = 0, use = 0, gstep = 0;
      for (int item = cluster_id; item < n_items; item += n_clusters) {
        const int64_t tile = 2 * (int64_t)(DUO ? item : item >> 1) + rank;
        for (int s = 0; s < L; ++s, ++gstep) {
          for (int c = 0; c < NCH; ++c) {
            {
              const int dd = MODE == 0 ? 0 : MODE == 1 ? (c & 1) : il_d(c);
              const int j = MODE == 0 ? c : MODE == 1 ? (c >> 1) : il_j(c);
              const int d = DUO ? dd : (item & 1);
              const int t = d ? (L - 1 - s) : s;
              const int tprev = d ? t + 1 : t - 1;
              // first slab (2048-byte units) of this thread's part of x_t / h_{t-1}
              const int xs0 = x_short ? (int)(((tile * L + t) * P + pp) * 2) : (int)((((tile * L + t) * 8) * P + pp) * 8);
              const int hs0 = s == 0 ? (int)((((tile * 2 + d) * 4) * P + pp) * 8)
                                     : (int)((((tile * L + tprev) * 8 + d * 4) * P + pp) * 8);
              const int wj = (d * 4 + j) * WJ;
              for (int part = 0; part < 2; ++part) {
                const int total = part == 0 ? kx : 32;
                const int ws0 = wj + (part ? XS : 0) + (int)rank * P * total + pp * total;
                for (int so = 0; so < total; so += KS) {
                  const int ns = (total - so) < KS ? (total - so) : KS;
                  if (!is_b && part == 1 && so == 0 && j == 0 && gstep > 0) {
                    mbar_wait(h_ready + 8 * dd, (gstep - 1) & 1);  // h_{t_prev} of this CTA's rows is in the act image
                    fence_proxy_async_all();
                  }
                  mbar_wait(empty0 + 8 * stage, (use & 1) ^ 1);
                  const uint32_t sb = smem_base + stage * C::STAGE;
                  if (role == 0 && rank == 0) mbar_expect_tx(full0 + 8 * stage, 2u * (uint32_t)(P * ns) * (GH_SLAB + A_SLAB));
                  const uint32_t fb = leader_full0 + 8 * stage;
                  if (is_b) {
                    tma2d_pair(sb + pp * C::B_PART, (part == 0 && x_short) ? &tm_wx : &tm_w, 0, ws0 + so, fb);
                  } else {
                    const uint32_t dst = sb + P * C::B_PART + pp * C::A_PART;
                    if (part == 0 && x_short) {
                      tma2d_pair(dst, &tm_xs, 0, xs0, fb);
                    } else {
                      const int sl = (so >> 3) * (P * 8) + (so & 7);
                      const void* tm = part == 0 ? &tm_x : (s == 0 ? &tm_h0 : &tm_o);
                      const int c1 = (part == 0 ? xs0 : hs0) + sl;
                      // the step's last read of x_t / h_{t-1} leaves L2 first (CCSM_TC_L2HINT 4, see tc_gru_layer_kernel)
                      if (p.l2_hint == 4 && j == 3) tma2d_pair_hint(dst, tm, 0, c1, fb, pol_first);
                      else tma2d_pair(dst, tm, 0, c1, fb);
                    }
                  }
                  if (++stage == S) {
                    stage = 0;
                    ++use;
                  }
                }
              }
            }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (rank == 0 && elect_one()) {
      // ===================== MMA issuer for the pair =====================
      constexpr uint32_t idesc = make_idesc(256, 192, F16);
      constexpr uint32_t idesc8 = make_idesc_e4m3(256, 192);
      uint32_t stage = 0, use = 0, chunk = 0;  // chunk: per TMEM buffer pair (non-DUO: every chunk; DUO: per recurrence)
      for (int item = cluster_id; item < n_items; item += n_clusters) {
        for (int s = 0; s < L; ++s) {
          for (int c = 0; c < NCH; ++c) {
            {
              // MODE 1: buffer = recurrence, chunk counts (s, j) pairs; otherwise the buffers alternate chunk by chunk
              const uint32_t buf = MODE == 1 ? (uint32_t)(c & 1) : (chunk & 1), u = MODE == 1 ? chunk : (chunk >> 1);
              mbar_wait(tmem_empty + 8 * buf, u & 1);  // completion #u: #0 = initial arming, #k = drain of use k-1
              tc_fence_after();
              const uint32_t dcol = tmem + buf * 256;
              for (int part = 0; part < 2; ++part) {
                const int total = part == 0 ? kx : 32;
                const uint32_t dpart = dcol + (part == 0 ? 0 : 64);  // X -> (n_i, r, z); H -> (r, z, n_h)
                for (int so = 0; so < total; so += KS) {
                  const int ns = (total - so) < KS ? (total - so) : KS;
                  mbar_wait(full0 + 8 * stage, use & 1);
                  tc_fence_after();
                  const uint32_t sb = smem_base + stage * C::STAGE;
                  const uint32_t a0 = sb + P * C::B_PART, b0 = sb;
                  if (C8 && ns == KS) {
#pragma unroll
                    for (int q = 0; q < 2; ++q)
                      umma_f16_pair(dpart, make_smem_desc(a0 + q * 2 * A_SLAB, A_SLAB, 128),
                                    make_smem_desc(b0 + q * 2 * GH_SLAB, GH_SLAB, 128), idesc, 1u);
#pragma unroll
                    for (int q = 0; q < 2; ++q)
                      umma_f8_pair(dpart, make_smem_desc(a0 + C::A_PART + q * 2 * A_SLAB, A_SLAB, 128),
                                   make_smem_desc(b0 + C::B_PART + q * 2 * GH_SLAB, GH_SLAB, 128), idesc8, 1u);
                  } else {
                    for (int ks = 0; ks < ns / 2; ++ks) {
#pragma unroll
                      for (int pass = 0; pass < (P == 2 ? 3 : 1); ++pass) {
                        const int pa = pass == 2 ? 1 : 0, pb = pass == 1 ? 1 : 0;
                        umma_f16_pair(dpart, make_smem_desc(a0 + pa * C::A_PART + ks * 2 * A_SLAB, A_SLAB, 128),
                                      make_smem_desc(b0 + pb * C::B_PART + ks * 2 * GH_SLAB, GH_SLAB, 128), idesc, 1u);
                      }
                    }
                  }
                  umma_commit_pair(empty0 + 8 * stage, 0x3);
                  if (++stage == S) {
                    stage = 0;
                    ++use;
                  }
                }
              }
              umma_commit_pair(tmem_full + 8 * buf, 0x3);
              if (MODE != 1 || (c & 1)) ++chunk;
            }
          }
        }
      }
    }
    __syncwarp();
  } else {
    // ===================== gate epilogue of this CTA's row tile =====================
    // MODE 0: warps 2-5, every chunk.  MODE 1: warps 2-5 the forward, 6-9 the reverse recurrence.  MODE 2: warps 2-9, every
    // chunk, warps 2-5 the first 32 units of a chunk and 6-9 the last 32.
    const int grp = (warp - 2) >> 2;
    const int rec = MODE == 1 ? grp : 0;
    constexpr int NSB = MODE == 2 ? 4 : 8;           // 8-unit sub-blocks per thread and chunk
    const int sb0 = MODE == 2 ? grp * 4 : 0;
    const int quad = warp & 3;  // tcgen05.ld lane rule: a warp touches TMEM lanes [32 * (warp % 4), +32)
    const int row = quad * 32 + lane;
    const uint32_t trow0 = tmem + ((uint32_t)(quad * 32) << 16);
    const uint32_t remote_empty = mapa_u32(tmem_empty, 0);
    // MODE 0: every item of this cluster has the same direction (even cluster count)
    const int d_fixed = MODE == 1 ? rec : (MODE == 0 ? (cluster_id & 1) : 0);
    if constexpr (MODE == 1) {
#pragma unroll
      for (int ub = 0; ub < 4; ++ub) arm_bias16(trow0 + rec * 256, ub, bias_s + d_fixed * 1024, ub * 16);  // unit-chunk 0
    } else {
      // buffers 0 / 1 <- unit-chunks 0 / 1 of the (first) direction
#pragma unroll
      for (int b = 0; b < 2; ++b) {
#pragma unroll
        for (int q = 0; q < NSB; ++q) arm_bias8(trow0 + b * 256, (sb0 + q) * 8, bias_s + d_fixed * 1024, b * 64 + (sb0 + q) * 8);
      }
    }
    tmem_st_wait();
    tc_fence_before();
    __syncwarp();
    if (lane == 0) {
      if (MODE == 1) {
        mbar_arrive_remote(remote_empty + 8 * rec);
      } else {
        mbar_arrive_remote(remote_empty);
        mbar_arrive_remote(remote_empty + 8);
      }
    }
    uint32_t chunk = 0;
    for (int item = cluster_id; item < n_items; item += n_clusters) {
      const int64_t tile = 2 * (int64_t)(DUO ? item : item >> 1) + rank;
      for (int s = 0; s < L; ++s) {
        for (int c = 0; c < (MODE == 2 ? 8 : 4); ++c, ++chunk) {
          const int d = MODE == 2 ? il_d(c) : d_fixed;
          const int j = MODE == 2 ? il_j(c) : c;
          const int t = d ? (L - 1 - s) : s;
          const int tprev = d ? t + 1 : t - 1;
          // the unit-chunk that uses this accumulator buffer next, and its direction's biases
          const int cn = (c + 2) & 7;
          const int jnext = MODE == 0 ? ((j + 2) & 3) : MODE == 1 ? ((j + 1) & 3) : il_j(cn);
          const float* bz = bias_s + (MODE == 2 ? il_d(cn) : d_fixed) * 1024;
          const uint8_t* hp_base =
              (s == 0) ? p.h0img + (((tile * 2 + d) * 4 + j) * P) * (size_t)CHUNK_BYTES
                       : p.out + (((tile * L + tprev) * 8 + d * 4 + j) * P) * (size_t)CHUNK_BYTES;
          uint8_t* out_base = p.out + (((tile * L + t) * 8 + d * 4 + j) * P) * (size_t)CHUNK_BYTES;
          // prefetch h_{t_prev} of this row's units (the L2 latency overlaps the chunk's MMAs)
          uint4 hph[NSB], hpl[NSB];
#pragma unroll
          for (int q = 0; q < NSB; ++q) {
            hph[q] = __ldcg(reinterpret_cast<const uint4*>(hp_base + (sb0 + q) * A_SLAB + row * 16));
            if constexpr (C8) {
              const uint2 t8 = __ldcg(reinterpret_cast<const uint2*>(hp_base + CHUNK_BYTES + 4096 + c8_off(sb0 + q) + row * 16));
              hpl[q] = make_uint4(t8.x, t8.y, 0, 0);
            } else if constexpr (P == 2)
              hpl[q] = __ldcg(reinterpret_cast<const uint4*>(hp_base + CHUNK_BYTES + (sb0 + q) * A_SLAB + row * 16));
            else
              hpl[q] = make_uint4(0, 0, 0, 0);
          }
          const uint32_t buf = MODE == 1 ? (uint32_t)rec : (chunk & 1), u = MODE == 1 ? chunk : (chunk >> 1);
          const uint32_t trow = trow0 + buf * 256;
          mbar_wait(tmem_full + 8 * buf, u & 1);
          tc_fence_after();
          uint32_t acc[2][4][8];  // [ping-pong][n_i, r, z, n_h][8 units]
          uint2 a8_even = make_uint2(0, 0), l8_even = make_uint2(0, 0);
#pragma unroll
          for (int g = 0; g < 4; ++g) tmem_ld8(trow + g * 64 + sb0 * 8, acc[0][g]);
#pragma unroll
          for (int q = 0; q < NSB; ++q) {
            const int sb = sb0 + q;
            const int col = sb * 8;
            tmem_ld_wait();  // sub-block q has landed (issued one iteration ago)
            if (q + 1 < NSB) {
#pragma unroll
              for (int g = 0; g < 4; ++g) tmem_ld8(trow + g * 64 + col + 8, acc[(q + 1) & 1][g]);
            }
            arm_bias8(trow, col, bz, jnext * 64 + col);
            float hp[8], hn[8];
            if constexpr (C8) join8_c8(hph[q], make_uint2(hpl[q].x, hpl[q].y), hp);
            else join8<P, F16>(hph[q], hpl[q], hp);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              float r, z;
              if constexpr (C8) {
                sigmoid2_s(__uint_as_float(acc[q & 1][1][i]), __uint_as_float(acc[q & 1][2][i]), r, z);
              } else {
                r = sig_<FAST, C8>(__uint_as_float(acc[q & 1][1][i]));
                z = sig_<FAST, C8>(__uint_as_float(acc[q & 1][2][i]));
              }
              const float n = tnh_<FAST, C8>(fmaf(r, __uint_as_float(acc[q & 1][3][i]), __uint_as_float(acc[q & 1][0][i])));
              hn[i] = fmaf(z, hp[i] - n, n);  // (1 - z) * n + z * h
            }
            if constexpr (C8) {
              uint4 hi;
              uint2 a8, l8;
              split8_c8(hn, hi, a8, l8);
              *reinterpret_cast<uint4*>(out_base + sb * A_SLAB + row * 16) = hi;
              if ((q & 1) == 0) {
                a8_even = a8;
                l8_even = l8;
              } else {
                uint8_t* b8 = out_base + CHUNK_BYTES + c8_off(sb - 1) + row * 16;
                *reinterpret_cast<uint4*>(b8) = make_uint4(a8_even.x, a8_even.y, a8.x, a8.y);
                *reinterpret_cast<uint4*>(b8 + 4096) = make_uint4(l8_even.x, l8_even.y, l8.x, l8.y);
              }
            } else {
              uint4 hi, lo;
              split8<P, F16>(hn, hi, lo);
              *reinterpret_cast<uint4*>(out_base + sb * A_SLAB + row * 16) = hi;
              if constexpr (P == 2) *reinterpret_cast<uint4*>(out_base + CHUNK_BYTES + sb * A_SLAB + row * 16) = lo;
            }
          }
          tmem_st_wait();
          tc_fence_before();
          if (j == 3) fence_proxy_async_all();  // generic-proxy global writes -> visible to the producers' tensor-map loads
          __syncwarp();
          if (lane == 0) {
            mbar_arrive_remote(remote_empty + 8 * buf);
            if (j == 3) mbar_arrive(h_ready + 8 * (MODE == 0 ? 0 : d));
          }
        }
      }
    }
  }
  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) tmem_dealloc2(tmem, 512);
}

// ------------------------------------------------------------------------------------------------
// attention + head kernel (reference utils/attention.py:48-70, models.py:135-150)
//   Qa = q . Wa^T -> TMEM cols [0,256);  per t: D_t = out_t . Ua^T -> TMEM cols [256,512)
//   e_t = va . tanh(Qa + D_t) (thread-local: TMEM lane = row), softmax over t, ctx = sum_t w_t out_t,
//   partial logits = fc1[:, strand*512 : +512] . ctx, two strands combined with a warp shuffle.
//   ONE pass over the layer's act image: fc1 . ctx = sum_t w_t (fc1 . out_t), and fc1 . out_t is taken by the head warps
//   from the A-operand stages of D_t while they sit in shared memory (each stage is released by the MMA commit AND the
//   four head warps); the softmax over t runs online (running max, rescaled sums), one step behind the scores.
// ------------------------------------------------------------------------------------------------
struct AttParams {
  const uint8_t* act;   // last layer's act image
  const uint8_t* wa_img;
  const uint8_t* ua_img;
  const float* va;      // [256]
  const float* fc_w;    // [2][1024]
  const float* fc_b;    // [2]
  float* logits;        // (n, 2) or null, already offset to the chunk's first site
  float* probs;
  int n_tiles;
  int L;
  int64_t sites;        // valid sites in this chunk
};

constexpr int ATT_STAGES = 4;
constexpr int ATT_THREADS = 384;  // warps 0, 2, 3 producers, 1 MMA, 4-7 score warps, 8-11 head warps (fc1 . out_t, online softmax)

template <int P>
struct AttCfg {
  static constexpr int KS = 8 / P;
  static constexpr uint32_t B_PART = KS * T_SLAB;
  static constexpr uint32_t A_PART = KS * A_SLAB;
  static constexpr uint32_t STAGE = P * (B_PART + A_PART);  // 49152
  static constexpr uint32_t SMEM = ATT_STAGES * STAGE + (256 + 2048 + 2 * 128) * 4;
};

template <int P, bool F16, bool C8 = false>
__global__ void __launch_bounds__(ATT_THREADS, 1) tc_att_head_kernel(const AttParams p) {
  static_assert(!C8 || (P == 2 && F16), "C8: fp16 images");
  using C = AttCfg<P>;
  constexpr int KS = C::KS;
  constexpr bool FAST = (P == 1);
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bars[2 * ATT_STAGES + 4];
  __shared__ uint32_t tmem_base_s;
  float* va_s = reinterpret_cast<float*>(smem + ATT_STAGES * C::STAGE);
  float* fc_s = va_s + 256;
  float* e_s = fc_s + 2048;  // [2][128]: score e_t of each row, slot = (running step count) & 1

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t full0 = smem_u32(&bars[0]), empty0 = smem_u32(&bars[ATT_STAGES]);
  const uint32_t d_full = smem_u32(&bars[2 * ATT_STAGES]), d_empty = smem_u32(&bars[2 * ATT_STAGES + 1]);
  const uint32_t e_full = smem_u32(&bars[2 * ATT_STAGES + 2]);  // [2]: score warps -> head warps, alternating per step
  if (threadIdx.x == 0) {
    for (int i = 0; i < ATT_STAGES; ++i) {
      mbar_init(full0 + 8 * i, 1);
      mbar_init(empty0 + 8 * i, 1 + 4);  // MMA commit + one arrival per head warp
    }
    mbar_init(d_full, 1);
    mbar_init(d_empty, 128);
    for (int i = 0; i < 2; ++i) mbar_init(e_full + 8 * i, 128);
    fence_barrier_init();
  }
  for (int i = threadIdx.x; i < 256; i += ATT_THREADS) va_s[i] = p.va[i];
  for (int i = threadIdx.x; i < 2048; i += ATT_THREADS) fc_s[i] = p.fc_w[i];
  if (warp == 1) {
    tmem_alloc(smem_u32(&tmem_base_s), 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  const uint32_t smem_base = smem_u32(smem);
  const int L = p.L;

  if (warp == 0 || warp == 2 || warp == 3) {
    // ===================== TMA producers: warp 0 = weights part 0 (+ expect_tx), warp 2 = weights part 1, warp 3 = activations
    // (one thread each, running source pointers: a single producer thread needed longer per stage than the stage's MMAs)
    const int role = warp == 0 ? 0 : warp - 1;
    if ((role != 1 || P == 2) && elect_one()) {
      uint32_t stage = 0, use = 0;
      for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
        for (int g = 0; g <= L; ++g) {  // g == 0: Qa; g >= 1: D_{t = g-1}
          const uint8_t* wsrc = (g == 0 ? p.wa_img : p.ua_img) + (role == 1 ? (size_t)64 * T_SLAB : 0);
          for (int so = 0; so < 64; so += KS) {
            mbar_wait(empty0 + 8 * stage, (use & 1) ^ 1);
            const uint32_t fb = full0 + 8 * stage;
            const uint32_t sb = smem_base + stage * C::STAGE;
            if (role == 0) mbar_expect_tx(fb, (uint32_t)(P * KS) * (T_SLAB + A_SLAB));
            if (role < 2) {
              bulk_g2s(sb + role * C::B_PART, wsrc, KS * T_SLAB, fb);
              wsrc += (size_t)KS * T_SLAB;
            } else {
              const int c = so >> 3;
              const int t = g == 0 ? (c < 4 ? L - 1 : 0) : g - 1;  // q = [h_n fwd (t = L-1) | h_n rev (t = 0)]
              const uint8_t* asrc = p.act + ((((int64_t)tile * L + t) * 8 + c) * P) * (size_t)CHUNK_BYTES + (so & 7) * A_SLAB;
#pragma unroll
              for (int pp = 0; pp < P; ++pp)
                bulk_g2s(sb + P * C::B_PART + pp * C::A_PART, asrc + (size_t)pp * CHUNK_BYTES, KS * A_SLAB, fb);
            }
            if (++stage == ATT_STAGES) {
              stage = 0;
              ++use;
            }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (elect_one()) {
      constexpr uint32_t idesc256 = make_idesc(128, 256, F16);
      constexpr uint32_t idesc256_8 = make_idesc_e4m3(128, 256);
      uint32_t stage = 0, use = 0, dcount = 0;
      for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
        for (int g = 0; g <= L; ++g) {
          if (g != 1) {
            // g == 0 overwrites Qa (previous tile fully drained), g >= 2 overwrites D: wait for the epilogue
            mbar_wait(d_empty, (dcount & 1) ^ 1);
            tc_fence_after();
          }
          const uint32_t dcol = tmem + (g == 0 ? 0 : 256);
          for (int so = 0; so < 64; so += KS) {
            mbar_wait(full0 + 8 * stage, use & 1);
            tc_fence_after();
            const uint32_t sb = smem_base + stage * C::STAGE;
            if constexpr (C8) {
              // 32 K elements per stage: two fp16 MMAs on the hi parts, then a8 . Wlo8 and alo8 . W8
              const uint32_t a0 = sb + P * C::B_PART, b0 = sb;
#pragma unroll
              for (int ks = 0; ks < 2; ++ks)
                umma_f16(dcol, make_smem_desc(a0 + ks * 2 * A_SLAB, A_SLAB, 128),
                         make_smem_desc(b0 + ks * 2 * T_SLAB, T_SLAB, 128), idesc256, (so == 0 && ks == 0) ? 0u : 1u);
#pragma unroll
              for (int c = 0; c < 2; ++c)
                umma_f8(dcol, make_smem_desc(a0 + C::A_PART + c * 2 * A_SLAB, A_SLAB, 128),
                        make_smem_desc(b0 + C::B_PART + c * 2 * T_SLAB, T_SLAB, 128), idesc256_8, 1u);
            } else
            for (int ks = 0; ks < KS / 2; ++ks) {
#pragma unroll
              for (int pass = 0; pass < (P == 2 ? 3 : 1); ++pass) {
                const int pa = pass == 2 ? 1 : 0, pb = pass == 1 ? 1 : 0;
                const uint64_t ad = make_smem_desc(sb + P * C::B_PART + pa * C::A_PART + ks * 2 * A_SLAB, A_SLAB, 128);
                const uint64_t bd = make_smem_desc(sb + pb * C::B_PART + ks * 2 * T_SLAB, T_SLAB, 128);
                umma_f16(dcol, ad, bd, idesc256, (so == 0 && ks == 0 && pass == 0) ? 0u : 1u);
              }
            }
            umma_commit(empty0 + 8 * stage);
            if (++stage == ATT_STAGES) {
              stage = 0;
              ++use;
            }
          }
          if (g >= 1) {
            umma_commit(d_full);
            ++dcount;
          }
        }
      }
    }
    __syncwarp();
  } else if (warp >= 4 && warp < 8) {
    // ===================== score warps: e_t = va . tanh(Qa + D_t) =====================
    const int quad = warp - 4;
    const int row = quad * 32 + lane;
    const uint32_t trow = tmem + ((uint32_t)(quad * 32) << 16);
    uint32_t dcount = 0;
    for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
      for (int t = 0; t < L; ++t, ++dcount) {
        mbar_wait(d_full, dcount & 1);
        tc_fence_after();
        // 16-column blocks, the tcgen05.ld of block cb + 1 in flight while block cb goes through the MUFU pipe
        float e0 = 0.f, e1 = 0.f;
        uint32_t q[2][16], dd[2][16];
        tmem_ld16(trow, q[0]);
        tmem_ld16(trow + 256, dd[0]);
#pragma unroll
        for (int cb = 0; cb < 16; ++cb) {
          tmem_ld_wait();
          if (cb + 1 < 16) {
            tmem_ld16(trow + (cb + 1) * 16, q[(cb + 1) & 1]);
            tmem_ld16(trow + 256 + (cb + 1) * 16, dd[(cb + 1) & 1]);
          }
#pragma unroll
          for (int i = 0; i < 16; i += 2) {
            e0 = fmaf(va_s[cb * 16 + i], tnh_<FAST, C8>(__uint_as_float(q[cb & 1][i]) + __uint_as_float(dd[cb & 1][i])), e0);
            e1 = fmaf(va_s[cb * 16 + i + 1],
                      tnh_<FAST, C8>(__uint_as_float(q[cb & 1][i + 1]) + __uint_as_float(dd[cb & 1][i + 1])), e1);
          }
        }
        const float e = e0 + e1;
        tc_fence_before();
        mbar_arrive(d_empty);
        // Slot dcount & 1 is free: the head warps read e of step dcount - 2 before they release the first stage of step
        // dcount, and D of step dcount needs all of its 16 stages.
        e_s[(dcount & 1) * 128 + row] = e;
        mbar_arrive(e_full + 8 * (dcount & 1));
      }
    }
  } else if (warp >= 8) {
    // ===================== head warps: g_t = fc1[:, strand half] . out_t from the operand stages, online softmax =====================
    const int row = (warp - 8) * 32 + lane;
    const int strand = row & 1;
    const float* f0 = fc_s + strand * 512;
    const uint32_t arow = (uint32_t)row * 16;
    uint32_t stage = 0, use = 0, ecount = 0;
    float mx = -INFINITY, sum = 0.f, a0 = 0.f, a1 = 0.f;  // running max, sum of weights, weighted g
    float g0 = 0.f, g1 = 0.f;                              // g of the step whose score is still on its way
    bool pending = false, pend_last = false;
    int pend_tile = 0;
    auto settle = [&]() {
      // the score of the pending step: it was finished while this warp went through the 16 stages that followed
      mbar_wait(e_full + 8 * (ecount & 1), (ecount >> 1) & 1);
      const float e = e_s[(ecount & 1) * 128 + row];
      ++ecount;
      const float m2 = fmaxf(mx, e);
      const float c = __expf(mx - m2), w = __expf(e - m2);
      sum = fmaf(sum, c, w);
      a0 = fmaf(a0, c, w * g0);
      a1 = fmaf(a1, c, w * g1);
      mx = m2;
      if (pend_last) {
        const float inv = 1.f / sum;
        float lg0 = a0 * inv, lg1 = a1 * inv;
        lg0 += __shfl_xor_sync(0xffffffffu, lg0, 1);  // strand 1 + strand 2 of the same site (adjacent rows)
        lg1 += __shfl_xor_sync(0xffffffffu, lg1, 1);
        const int64_t site = ((int64_t)pend_tile * TILE_ROWS + row) >> 1;
        if (strand == 0 && site < p.sites) {
          lg0 += p.fc_b[0];
          lg1 += p.fc_b[1];
          const float m3 = fmaxf(lg0, lg1);
          const float e0 = __expf(lg0 - m3), e1 = __expf(lg1 - m3);
          const float is = 1.f / (e0 + e1);
          if (p.logits) {
            p.logits[site * 2] = lg0;
            p.logits[site * 2 + 1] = lg1;
          }
          if (p.probs) {
            p.probs[site * 2] = e0 * is;
            p.probs[site * 2 + 1] = e1 * is;
          }
        }
        mx = -INFINITY;
        sum = a0 = a1 = 0.f;
      }
      pending = false;
    };
    for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
      for (int g = 0; g <= L; ++g) {
        float n0 = 0.f, n1 = 0.f;
        for (int so = 0; so < 64; so += KS) {
          mbar_wait(full0 + 8 * stage, use & 1);
          if (g >= 1) {
            const uint8_t* ap = smem + stage * C::STAGE + P * C::B_PART + arow;
#pragma unroll
            for (int q = 0; q < KS; ++q) {
              const uint4 hi = *reinterpret_cast<const uint4*>(ap + q * A_SLAB);
              float v[8];
              if constexpr (C8) {
                const uint2 l8 = *reinterpret_cast<const uint2*>(ap + C::A_PART + 4096 + (q >> 1) * 2048 + (q & 1) * 8);
                join8_c8(hi, l8, v);
              } else if constexpr (P == 2) {
                join8<P, F16>(hi, *reinterpret_cast<const uint4*>(ap + C::A_PART + q * A_SLAB), v);
              } else {
                join8<P, F16>(hi, make_uint4(0, 0, 0, 0), v);
              }
              const float* f = f0 + (so + q) * 8;
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                n0 = fmaf(v[i], f[i], n0);
                n1 = fmaf(v[i], f[1024 + i], n1);
              }
            }
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(empty0 + 8 * stage);
          if (++stage == ATT_STAGES) {
            stage = 0;
            ++use;
          }
        }
        if (pending) settle();
        if (g >= 1) {
          g0 = n0;
          g1 = n1;
          pending = true;
          pend_last = g == L;
          pend_tile = tile;
        }
      }
    }
    if (pending) settle();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 512);
}

// act image (one layer) -> (rows, L, 512) fp32, for tests
template <int P, bool F16, bool C8 = false>
__global__ void tc_unpack_act_kernel(const uint8_t* __restrict__ act, float* __restrict__ out, int64_t rows, int L) {
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // over (row, t, slab 0..63)
  if (idx >= rows * L * 64) return;
  const int sl = (int)(idx % 64);
  const int t = (int)((idx / 64) % L);
  const int64_t R = idx / (64 * L);
  const int64_t tile = R / TILE_ROWS;
  const int r = (int)(R % TILE_ROWS);
  const uint8_t* src = act + (((tile * L + t) * 8 + (sl >> 3)) * P) * (size_t)CHUNK_BYTES + (sl & 7) * A_SLAB + r * 16;
  uint4 hi = *reinterpret_cast<const uint4*>(src);
  uint4 lo = make_uint4(0, 0, 0, 0);
  float v[8];
  if constexpr (C8) {
    const uint8_t* cb = act + (((tile * L + t) * 8 + (sl >> 3)) * P) * (size_t)CHUNK_BYTES;
    join8_c8(hi, *reinterpret_cast<const uint2*>(cb + CHUNK_BYTES + 4096 + c8_off(sl & 7) + r * 16), v);
  } else {
    if constexpr (P == 2) lo = *reinterpret_cast<const uint4*>(src + CHUNK_BYTES);
    join8<P, F16>(hi, lo, v);
  }
  float* o = out + (R * L + t) * 512 + sl * 8;
#pragma unroll
  for (int i = 0; i < 8; ++i) o[i] = v[i];
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
static uint16_t to_elem(float v, bool f16) {
  if (f16) {
    __half h = __float2half_rn(v);
    return *reinterpret_cast<uint16_t*>(&h);
  }
  __nv_bfloat16 h = __float2bfloat16_rn(v);
  return *reinterpret_cast<uint16_t*>(&h);
}
static float from_elem(uint16_t b, bool f16) {
  if (f16) return __half2float(*reinterpret_cast<__half*>(&b));
  return __bfloat162float(*reinterpret_cast<__nv_bfloat16*>(&b));
}

// Packs `rows` weight rows (given by row_of(n) -> pointer to K floats, K_src valid entries) into
// [part][kslabs][rows x 8 elems] at dst (uint16 elements).
template <class RowFn>
static void pack_image(uint16_t* dst, int rows, int kslabs, int K_src, int P, bool f16, RowFn row_of) {
  const size_t part_elems = (size_t)kslabs * rows * 8;
  for (int n = 0; n < rows; ++n) {
    const float* w = row_of(n);
    for (int k = 0; k < kslabs * 8; ++k) {
      const float v = k < K_src ? w[k] : 0.f;
      const size_t off = (size_t)(k / 8) * rows * 8 + (size_t)n * 8 + (k % 8);
      const uint16_t hi = to_elem(v, f16);
      dst[off] = hi;
      if (P == 2) dst[part_elems + off] = to_elem(v - from_elem(hi, f16), f16);
    }
  }
}

// C8 weight image: part 0 = fp16(S w) in 8-element K-slabs; part 1 = per 32 K elements two 16-element slabs of
// e4m3(S w - part 0) followed by two of e4m3(w); a slab = rows x 16 B.  Returns false if S w leaves the fp16 range.
template <class RowFn>
static bool pack_image_c8(uint16_t* dst, int rows, int kslabs, int K_src, RowFn row_of) {
  const size_t part_elems = (size_t)kslabs * rows * 8;
  uint8_t* p1 = reinterpret_cast<uint8_t*>(dst + part_elems);
  for (int n = 0; n < rows; ++n) {
    const float* w = row_of(n);
    for (int k = 0; k < kslabs * 8; ++k) {
      const float v = k < K_src ? w[k] : 0.f;
      const float sv = v * C8_S;
      if (!(fabsf(sv) < 60000.f)) return false;
      const uint16_t hi = to_elem(sv, true);
      dst[(size_t)(k / 8) * rows * 8 + (size_t)n * 8 + (k % 8)] = hi;
      const size_t off = (size_t)(k / 32) * 4 * rows * 16 + (size_t)((k % 32) / 16) * rows * 16 + (size_t)n * 16 + (k % 16);
      p1[off] = (uint8_t)__nv_cvt_float_to_fp8(sv - from_elem(hi, true), __NV_SATFINITE, __NV_E4M3);
      p1[off + (size_t)2 * rows * 16] = (uint8_t)__nv_cvt_float_to_fp8(v, __NV_SATFINITE, __NV_E4M3);
    }
  }
  return true;
}

static const HostTensor* findw(ccsm_model* m, const std::string& k) {
  auto it = m->w.find(k);
  return it == m->w.end() ? nullptr : &it->second;
}

int tc_upload_weights(ccsm_model* m) {
  const int H = m->cfg.hidden, NL = m->cfg.num_layers;
  if (m->cfg.kind != CCSM_KIND_ATT2S || H != 256 || m->cfg.num_classes != 2 || m->cfg.n_embed != 8 ||
      m->in_feat > 16 || (m->cfg.feat_flags & ~CCSM_FEAT_NPASS) != 0) {
    set_error("tensor-core path supports the shipped attbigru2s configuration (hidden 256, 2 classes, "
              "embed 8, kmer+ipd+pw[+npass]); use precision fp32 for other configurations");
    return CCSM_EUNSUPPORTED;
  }
  if (!m->tc) m->tc = new TcState();
  TcState& T = *m->tc;
  const int prec = m->cfg.precision;
  const bool c8 = prec == CCSM_PREC_FP16C8;
  const int P = (prec == CCSM_PREC_BF16X3 || prec == CCSM_PREC_FP16X3 || c8) ? 2 : 1;
  const bool f16 = (prec == CCSM_PREC_FP16X3 || prec == CCSM_PREC_FP16 || c8);
  bool c8_ok = true;
  cudaDeviceProp prop;
  CCSM_CUDA(cudaGetDeviceProperties(&prop, m->cfg.device));
  T.sm_count = prop.multiProcessorCount;
  static const char* sfx[2] = {"", "_reverse"};
  T.wimg.resize(NL);
  T.wpair.resize(NL);
  T.kx_slabs.resize(NL);
  std::vector<float> bias((size_t)NL * 2 * 4 * H);
  for (int l = 0; l < NL; ++l) {
    const int K = l == 0 ? m->in_feat : 2 * H;
    const int kxs = l == 0 ? 2 : (2 * H) / 8;
    T.kx_slabs[l] = kxs;
    const size_t x_elems = (size_t)P * kxs * 192 * 8, h_elems = (size_t)P * 32 * 192 * 8;
    std::vector<uint16_t> img((x_elems + h_elems) * 8);
    for (int d = 0; d < 2; ++d) {
      const HostTensor* wih = findw(m, "rnn.weight_ih_l" + std::to_string(l) + sfx[d]);
      const HostTensor* whh = findw(m, "rnn.weight_hh_l" + std::to_string(l) + sfx[d]);
      const HostTensor* bih = findw(m, "rnn.bias_ih_l" + std::to_string(l) + sfx[d]);
      const HostTensor* bhh = findw(m, "rnn.bias_hh_l" + std::to_string(l) + sfx[d]);
      if (!wih || !whh || !bih || !bhh) {
        set_error("tc finalize: missing GRU tensors for layer %d", l);
        return CCSM_EKEY;
      }
      for (int j = 0; j < 4; ++j) {
        uint16_t* base = img.data() + (size_t)(d * 4 + j) * (x_elems + h_elems);
        // X part rows: n_i, r, z   (PyTorch gate row order in weight_ih: r [0,H), z [H,2H), n [2H,3H))
        // single-pass modes: r/z rows pre-scaled by 0.5 (exact) for the one-MUFU sigmoid, see sigmoid_<FAST>
        std::vector<float> tmp((size_t)(K > H ? K : H));
        auto xrow = [&](int n) {
          const int g = n / 64, u = j * 64 + n % 64;
          const int row = (g == 0 ? 2 * H : (g == 1 ? 0 : H)) + u;
          return wih->data.data() + (size_t)row * K;
        };
        auto hrow = [&](int n) {
          const int g = n / 64, u = j * 64 + n % 64;
          const int row = (g == 0 ? 0 : (g == 1 ? H : 2 * H)) + u;
          return whh->data.data() + (size_t)row * H;
        };
        if (c8) {
          // layer 0's K = 11 input keeps the 3-pass fp16 split, at the accumulator scale S
          if (kxs % 4 != 0)
            pack_image(base, 192, kxs, K, P, f16, [&](int n) {
              const float* w = xrow(n);
              for (int k = 0; k < K; ++k) {
                tmp[k] = C8_S * w[k];
                if (!(fabsf(tmp[k]) < 60000.f)) c8_ok = false;
              }
              return (const float*)tmp.data();
            });
          else
            c8_ok = pack_image_c8(base, 192, kxs, K, xrow) && c8_ok;
          c8_ok = pack_image_c8(base + x_elems, 192, 32, H, hrow) && c8_ok;
        } else {
        pack_image(base, 192, kxs, K, P, f16, [&](int n) {
          const float* w = xrow(n);
          if (P == 2 || n / 64 == 0) return w;
          for (int k = 0; k < K; ++k) tmp[k] = 0.5f * w[k];
          return (const float*)tmp.data();
        });
        // H part rows: r, z, n_h
        pack_image(base + x_elems, 192, 32, H, P, f16, [&](int n) {
          const float* w = hrow(n);
          if (P == 2 || n / 64 == 2) return w;
          for (int k = 0; k < H; ++k) tmp[k] = 0.5f * w[k];
          return (const float*)tmp.data();
        });
        }
      }
      float* b = bias.data() + ((size_t)l * 2 + d) * 4 * H;
      for (int u = 0; u < H; ++u) {
        const float gs = P == 1 ? 0.5f : 1.f;                // matches the r/z row scaling above
        const float as = c8 ? C8_S : 1.f;                    // C8: accumulators carry the scale S
        b[u] = as * gs * (bih->data[u] + bhh->data[u]);             // b_r
        b[H + u] = as * gs * (bih->data[H + u] + bhh->data[H + u]); // b_z
        b[2 * H + u] = as * bih->data[2 * H + u];            // b_in
        b[3 * H + u] = as * bhh->data[2 * H + u];            // b_hn (inside r * (.))
      }
    }
    CCSM_TRY(T.wimg[l].reserve(img.size() * 2));
    CCSM_CUDA(cudaMemcpy(T.wimg[l].p, img.data(), img.size() * 2, cudaMemcpyHostToDevice));
    {
      // CTA-pair layout: same rows, split into halves of 96: [half][part][slab][96 rows x 8]
      std::vector<uint16_t> pimg(img.size());
      for (int dj = 0; dj < 8; ++dj)
        for (int part = 0; part < 2; ++part) {
          const int ks = part == 0 ? kxs : 32;
          const uint16_t* src = img.data() + (size_t)dj * (x_elems + h_elems) + (part ? x_elems : 0);
          uint16_t* dst = pimg.data() + (size_t)dj * (x_elems + h_elems) + (part ? x_elems : 0);
          for (int pp = 0; pp < P; ++pp)
            for (int sl = 0; sl < ks; ++sl)
              for (int n = 0; n < 192; ++n) {
                const int c = n / 96, nn = n % 96;
                const uint16_t* a = src + ((size_t)pp * ks + sl) * 192 * 8 + (size_t)n * 8;
                uint16_t* b = dst + (((size_t)c * P + pp) * ks + sl) * 96 * 8 + (size_t)nn * 8;
                for (int e = 0; e < 8; ++e) b[e] = a[e];
              }
        }
      CCSM_TRY(T.wpair[l].reserve(pimg.size() * 2));
      CCSM_CUDA(cudaMemcpy(T.wpair[l].p, pimg.data(), pimg.size() * 2, cudaMemcpyHostToDevice));
    }
  }
  CCSM_TRY(T.bias.reserve(bias.size() * 4));
  CCSM_CUDA(cudaMemcpy(T.bias.p, bias.data(), bias.size() * 4, cudaMemcpyHostToDevice));
  for (int which = 0; which < 2; ++which) {
    const HostTensor* w = findw(m, which == 0 ? "_att3.Wa.weight" : "_att3.Ua.weight");
    std::vector<uint16_t> img((size_t)P * 64 * 256 * 8);
    if (c8) c8_ok = pack_image_c8(img.data(), 256, 64, 2 * H, [&](int n) { return w->data.data() + (size_t)n * 2 * H; }) && c8_ok;
    else pack_image(img.data(), 256, 64, 2 * H, P, f16, [&](int n) { return w->data.data() + (size_t)n * 2 * H; });
    DevBuf& dst = which == 0 ? T.wa_img : T.ua_img;
    CCSM_TRY(dst.reserve(img.size() * 2));
    CCSM_CUDA(cudaMemcpy(dst.p, img.data(), img.size() * 2, cudaMemcpyHostToDevice));
  }
  auto up = [&](DevBuf& b, const char* key) -> int {
    const HostTensor* w = findw(m, key);
    CCSM_TRY(b.reserve(w->data.size() * 4));
    CCSM_CUDA(cudaMemcpy(b.p, w->data.data(), w->data.size() * 4, cudaMemcpyHostToDevice));
    return CCSM_OK;
  };
  CCSM_TRY(up(T.va, "_att3.va.weight"));
  CCSM_TRY(up(T.fc_w, "fc1.weight"));
  CCSM_TRY(up(T.fc_b, "fc1.bias"));
  CCSM_TRY(up(T.embed, "embed.weight"));
  if (!c8_ok) {
    T.P = 0;
    set_error("precision fp16c8: a weight times 2^12 leaves the fp16 range; use fp16x3 for this checkpoint");
    return CCSM_EUNSUPPORTED;
  }
  T.P = P;
  T.f16 = f16;
  T.c8 = c8;
  return CCSM_OK;
}

void tc_release(ccsm_model* m) {
  if (!m->tc) return;
  TcState& T = *m->tc;
  for (auto& b : T.wimg) b.release();
  for (auto& b : T.wpair) b.release();
  T.bias.release(); T.wa_img.release(); T.ua_img.release(); T.va.release(); T.fc_w.release(); T.fc_b.release();
  T.embed.release(); T.x0img.release(); T.h0img.release();
  for (auto& b : T.act) b.release();
  delete m->tc;
  m->tc = nullptr;
}

static int tc_reserve(ccsm_model* m, int64_t tiles) {
  TcState& T = *m->tc;
  if (tiles <= T.tiles_cap && T.ws_P == T.P) return CCSM_OK;
  const int L = m->cfg.seq_len, NL = m->cfg.num_layers, P = T.P;
  CCSM_TRY(T.x0img.reserve((size_t)tiles * L * P * 2 * A_SLAB));
  CCSM_TRY(T.h0img.reserve((size_t)NL * tiles * 2 * 4 * P * CHUNK_BYTES));
  for (int i = 0; i < 3; ++i) CCSM_TRY(T.act[i].reserve((size_t)tiles * L * 8 * P * CHUNK_BYTES));
  T.tiles_cap = tiles;
  T.ws_P = P;
  return CCSM_OK;
}

// GRU kernel variant: 0 = (NSLOT 1, NBUF 1, two CTAs per SM), 1 = (NSLOT 2, NBUF 1), 2 = (NSLOT 1, NBUF 2),
// 5 = variant 0 with 40 KB stages, 6 = variant 2 with two epilogue warps per quadrant,
// 7 / 8 / 9 = variants 2 / 0 / 5 as clusters of two CTAs sharing every weight stage by TMA multicast,
// a / b (10 / 11) = h_t kept in shared memory between steps (HS), 8 / 4 epilogue warps,
// c / d / e (12 / 13 / 14) = variants 0 / 2 / 6 with the software-pipelined gate epilogue (PIPE),
// 3 = CTA pair (cta_group::2, M = 256, TMEM double-buffered), 4 = CTA pair, two clusters per TPC (NBUF 1).
// Selectable per layer class for experiments: CCSM_TC_VARIANT="<layer0><layers>=1>", e.g. "02".
// Measured defaults (profiles/r01_variants.md): layer 0 (K_in = 16, latency-bound) -> 0 (two CTAs per SM);
// layers >= 1 -> 2 (one CTA per SM, TMEM double-buffered, 40 KB stages = 4-6 MMAs per barrier round trip of the
// issuing thread, half the row tiles in flight so the activation re-reads stay in L2).
// The CTA-pair kernels (3, 4) are correct but slower in round 1.
static int gru_variant(int layer, int P) {
  static int v[2] = {-2, -2};
  if (v[0] == -2) {
    const char* e = getenv("CCSM_TC_VARIANT");
    auto dig = [](char c) { return (c >= '0' && c <= '9') ? c - '0' : (c >= 'a' && c <= 'z') ? 10 + (c - 'a') : -1; };
    v[0] = e ? dig(e[0]) : -1;
    v[1] = (e && e[0] && dig(e[1]) >= 0) ? dig(e[1]) : v[0];
  }
  const int forced = v[layer == 0 ? 0 : 1];
  if (forced >= 0) return forced;
  // Defaults: layers >= 1 -> n (CTA-pair kernel with tensor-map loads, profiles/r02_bound.md section 9); layer 0 ->
  // l (8 epilogue warps + PIPE, both directions of a row tile interleaved in one CTA: profiles/r02_bound.md section 7).
  return layer > 0 ? 23 : 21;
}

template <int P, bool F16, int NSLOT, int NBUF, int KSB, int EPIW = 1, bool MC = false, bool HS = false, bool PIPE = false,
          bool C8 = false, bool IL = false>
static int launch_gru(const GruParams& gp, int64_t tiles, int sm_count, cudaStream_t st) {
  using C = GruCfg<P, NSLOT, NBUF, KSB, EPIW, HS>;
  static bool attr = false;
  auto kern = tc_gru_layer_kernel<P, F16, NSLOT, NBUF, KSB, EPIW, MC, HS, PIPE, C8, IL>;
  if (!attr) {
    CCSM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM));
    attr = true;
  }
  const int64_t slots = (int64_t)sm_count * C::CTAS_PER_SM;
  if constexpr (MC) {
    // clusters of two CTAs; an even number of clusters so that every cluster keeps one direction
    const int64_t items = (tiles / 2) * 2;  // pairs of row tiles x 2 directions
    const int64_t max_clusters = slots / 2;
    const int clusters = (int)(items < max_clusters ? items : max_clusters) & ~1;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(2 * clusters));
    cfg.blockDim = dim3(C::THREADS + 32 * GruProd<P, NSLOT, MC, HS>::EXTRA_WARPS);
    cfg.dynamicSmemBytes = C::SMEM;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    CCSM_CUDA(cudaLaunchKernelEx(&cfg, kern, gp));
    return CCSM_OK;
  }
  if constexpr (IL) {
    const int grid = (int)(tiles < slots ? tiles : slots);  // one item = one row tile, both directions
    kern<<<grid, C::THREADS + 32 * GruProd<P, NSLOT, MC, HS>::EXTRA_WARPS, C::SMEM, st>>>(gp);
    return CCSM_OK;
  }
  const int64_t items = (tiles / NSLOT) * 2;  // groups of NSLOT row tiles x 2 directions
  const int grid = (int)(items < slots ? items : slots) & ~1;  // even: fixed direction per CTA
  kern<<<grid, C::THREADS + 32 * GruProd<P, NSLOT, MC, HS>::EXTRA_WARPS, C::SMEM, st>>>(gp);
  return CCSM_OK;
}

// variant -> (NSLOT, NBUF, KSB); 3 and 4 are the CTA-pair kernels
template <int P, bool F16, bool C8>
static int launch_gru_variant(int variant, const GruParams& gp, int64_t tiles, int sm_count, cudaStream_t st) {
  if constexpr (C8) {
    // the C8 mode is built on the pipelined-epilogue kernels only
    switch (variant) {
      case 13: return launch_gru<P, F16, 1, 2, 8, 1, false, false, true, true>(gp, tiles, sm_count, st);
      case 14: return launch_gru<P, F16, 1, 2, 8, 2, false, false, true, true>(gp, tiles, sm_count, st);
      case 15: return launch_gru<P, F16, 2, 1, 8, 1, false, false, true, true>(gp, tiles, sm_count, st);
      case 17: return launch_gru<P, F16, 1, 1, 8, 1, false, false, true, true>(gp, tiles, sm_count, st);  // two CTAs per SM
      case 20: return launch_gru<P, F16, 1, 2, 8, 1, false, false, true, true, true>(gp, tiles, sm_count, st);  // d, interleaved
      case 21: return launch_gru<P, F16, 1, 2, 8, 2, false, false, true, true, true>(gp, tiles, sm_count, st);  // e, interleaved
      case 22: return launch_gru<P, F16, 1, 2, 8, 4, false, false, true, true, true>(gp, tiles, sm_count, st);  // 16 epilogue warps
      default:
        set_error("precision fp16c8 runs GRU kernel variants d, e, f, g, h, i, j, k, l only (got %d)", variant);
        return CCSM_EINVAL;
    }
  } else
  switch (variant) {
    case 0: return launch_gru<P, F16, 1, 1, 4>(gp, tiles, sm_count, st);
    case 1: return launch_gru<P, F16, 2, 1, 8>(gp, tiles, sm_count, st);
    case 2: return launch_gru<P, F16, 1, 2, 8>(gp, tiles, sm_count, st);
    case 5: return launch_gru<P, F16, 1, 1, 8>(gp, tiles, sm_count, st);
    case 6: return launch_gru<P, F16, 1, 2, 8, 2>(gp, tiles, sm_count, st);  // 2 + 8 epilogue warps
    case 7: return launch_gru<P, F16, 1, 2, 8, 1, true>(gp, tiles, sm_count, st);  // 2 + weight multicast
    case 8: return launch_gru<P, F16, 1, 1, 4, 1, true>(gp, tiles, sm_count, st);  // 0 + weight multicast
    case 9: return launch_gru<P, F16, 1, 1, 8, 1, true>(gp, tiles, sm_count, st);  // 5 + weight multicast
    case 10:  // h_t kept in shared memory (single-pass modes; x3 falls back to variant 6)
      if constexpr (P == 1) return launch_gru<P, F16, 1, 2, 4, 2, false, true>(gp, tiles, sm_count, st);
      else return launch_gru<P, F16, 1, 2, 8, 2>(gp, tiles, sm_count, st);
    case 11:
      if constexpr (P == 1) return launch_gru<P, F16, 1, 2, 4, 1, false, true>(gp, tiles, sm_count, st);
      else return launch_gru<P, F16, 1, 2, 8, 2>(gp, tiles, sm_count, st);
    case 12: return launch_gru<P, F16, 1, 1, 4, 1, false, false, true>(gp, tiles, sm_count, st);  // 0 + pipelined epilogue
    case 13: return launch_gru<P, F16, 1, 2, 8, 1, false, false, true>(gp, tiles, sm_count, st);  // 2 + pipelined epilogue
    case 14: return launch_gru<P, F16, 1, 2, 8, 2, false, false, true>(gp, tiles, sm_count, st);  // 6 + pipelined epilogue
    case 15: return launch_gru<P, F16, 2, 1, 8, 1, false, false, true>(gp, tiles, sm_count, st);  // 1 + pipelined epilogue
    case 17: return launch_gru<P, F16, 1, 1, 8, 1, false, false, true>(gp, tiles, sm_count, st);  // 5 + pipelined epilogue
    case 18: return launch_gru<P, F16, 1, 2, 8, 1, false, false, true>(gp, tiles, sm_count, st);  // (fp16c8-only kernel: d here)
    case 19: return launch_gru<P, F16, 1, 2, 8, 2, false, false, true>(gp, tiles, sm_count, st);  // (fp16c8-only kernel: e here)
    case 20: return launch_gru<P, F16, 1, 2, 8, 1, false, false, true, false, true>(gp, tiles, sm_count, st);  // d, both directions interleaved
    case 21: return launch_gru<P, F16, 1, 2, 8, 2, false, false, true, false, true>(gp, tiles, sm_count, st);  // e, both directions interleaved
    case 22: return launch_gru<P, F16, 1, 2, 8, 4, false, false, true, false, true>(gp, tiles, sm_count, st);  // 16 epilogue warps
    default:
      set_error("unknown GRU kernel variant %d", variant);
      return CCSM_EINVAL;
  }
}

template <int P, bool F16, bool C8 = false>
static int tc_run_chunk(ccsm_model* m, int64_t sites, int64_t site0, int64_t n_total, const ccsm_strand* fwd,
                        const ccsm_strand* rev, const float* h0_f, const float* h0_r, float* logits, float* probs,
                        cudaStream_t st) {
  TcState& T = *m->tc;
  const int L = m->cfg.seq_len, NL = m->cfg.num_layers;
  int64_t tiles = (sites * 2 + TILE_ROWS - 1) / TILE_ROWS;
  tiles += tiles & 1;
  T.last_tiles = tiles;
  TcStrand s0{fwd->kmer, fwd->kpass, fwd->ipd_means, fwd->pw_means}, s1{rev->kmer, rev->kpass, rev->ipd_means, rev->pw_means};
  const int64_t rows = tiles * TILE_ROWS;
  int pid = m->prof.begin(PROF_PREP, (double)sites, st);
  tc_prep_kernel<P, F16, C8><<<(unsigned)((rows + 127) / 128), 128, 0, st>>>(
      tiles, sites, site0, n_total, L, NL, m->cfg.n_vocab, (m->cfg.feat_flags & CCSM_FEAT_NPASS) ? 1 : 0, s0, s1,
      T.embed.as<float>(), h0_f, h0_r, m->h0_mode == CCSM_H0_DEVICE_RANDOM ? 1 : 0,
      (unsigned long long)m->h0_seed, (unsigned long long)(m->h0_calls * 256), T.x0img.as<uint8_t>(),
      T.h0img.as<uint8_t>());
  m->prof.end(pid, st);
  count_launch();
  static bool attr_set = false;  // one flag per instantiation of this function template
  if (!attr_set) {
    if constexpr (!C8) {
      CCSM_CUDA(cudaFuncSetAttribute(tc_gru_pair_kernel<P, F16, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)PairCfg<P, 2>::SMEM));
      CCSM_CUDA(cudaFuncSetAttribute(tc_gru_pair_kernel<P, F16, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)PairCfg<P, 1>::SMEM));
    }
    CCSM_CUDA(cudaFuncSetAttribute(tc_att_head_kernel<P, F16, C8>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)AttCfg<P>::SMEM));
    attr_set = true;
  }
  for (int l = 0; l < NL; ++l) {
    GruParams gp;
    gp.xin = l == 0 ? T.x0img.as<uint8_t>() : T.act[(l - 1) % 3].as<uint8_t>();
    gp.h0img = T.h0img.as<uint8_t>() + (size_t)l * tiles * 2 * 4 * P * CHUNK_BYTES;
    gp.out = T.act[l % 3].as<uint8_t>();
    gp.wimg = T.wimg[l].as<uint8_t>();
    gp.bias = T.bias.as<float>() + (size_t)l * 2 * 4 * 256;
    gp.n_tiles = (int)tiles;
    gp.L = L;
    gp.kx_slabs = (int)T.kx_slabs[l];
    {
      static int hint = -1;
      if (hint < 0) {
        const char* e = getenv("CCSM_TC_L2HINT");
        hint = e ? atoi(e) : -2;
      }
      // default: hi/lo images mark the last read of x_t and of h_{t-1} evict_first (22 -> 16 GB of DRAM reads per fp16c8
      // layer launch, profiles/r02_bound.md section 8); the single-pass images already sit at the 2x-image floor
      gp.l2_hint = hint == -2 ? (P == 2 ? 4 : 0) : hint;
    }
    pid = m->prof.begin(l == 0 ? PROF_GRU_L0 : PROF_GRU_LN, (double)sites, st);
    const int variant = gru_variant(l, P);
    if (C8 && (variant == 18 || variant == 19)) {
      // fp16c8 with on-chip operand conversion (30 KB instead of 40 KB through the SM's L2 port per stage)
      static bool cv_attr = false;
      if (!cv_attr) {
        CCSM_CUDA(cudaFuncSetAttribute(tc_gru_cv_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CvCfg<1>::SMEM));
        CCSM_CUDA(cudaFuncSetAttribute(tc_gru_cv_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CvCfg<2>::SMEM));
        cv_attr = true;
      }
      const int64_t items = tiles * 2;
      const int grid = (int)(items < T.sm_count ? items : T.sm_count) & ~1;
      if (variant == 18) tc_gru_cv_kernel<1><<<grid, CvCfg<1>::THREADS, CvCfg<1>::SMEM, st>>>(gp);
      else tc_gru_cv_kernel<2><<<grid, CvCfg<2>::THREADS, CvCfg<2>::SMEM, st>>>(gp);
    } else if (variant == 23 || variant == 24 || variant == 25) {
      // pair2 kernel: CTA pairs, tensor-map loads completing on the leader's barrier (24 = both directions interleaved)
      gp.wimg = T.wpair[l].as<uint8_t>();
      const bool duo = variant != 23;
      constexpr int KSP = Pair2Cfg<P, 0>::KS;
      const int kx = (int)T.kx_slabs[l];
      static bool p2_attr = false;
      if (!p2_attr) {
        CCSM_CUDA(cudaFuncSetAttribute(tc_gru_pair2_kernel<P, F16, C8, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)Pair2Cfg<P, 0>::SMEM));
        CCSM_CUDA(cudaFuncSetAttribute(tc_gru_pair2_kernel<P, F16, C8, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)Pair2Cfg<P, 1>::SMEM));
        CCSM_CUDA(cudaFuncSetAttribute(tc_gru_pair2_kernel<P, F16, C8, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)Pair2Cfg<P, 2>::SMEM));
        p2_attr = true;
      }
      const bool x_short = kx < KSP;
      const uint64_t x_slabs = x_short ? (uint64_t)tiles * L * P * 2 : (uint64_t)tiles * L * 8 * P * 8;
      const uint64_t act_slabs = (uint64_t)tiles * L * 8 * P * 8, h0_slabs = (uint64_t)tiles * 2 * 4 * P * 8;
      const uint64_t w_slabs = (uint64_t)8 * (2 * P * kx + 2 * P * 32);
      const uint32_t xbox = x_short ? (uint32_t)kx : (uint32_t)KSP;
      CUtensorMap tm_w, tm_x, tm_o, tm_h0, tm_wx, tm_xs;
      if (make_slab_tmap(&tm_w, gp.wimg, GH_SLAB, w_slabs, KSP) || make_slab_tmap(&tm_wx, gp.wimg, GH_SLAB, w_slabs, xbox) ||
          make_slab_tmap(&tm_x, gp.xin, A_SLAB, x_slabs, xbox) || make_slab_tmap(&tm_xs, gp.xin, A_SLAB, x_slabs, xbox) ||
          make_slab_tmap(&tm_o, gp.out, A_SLAB, act_slabs, KSP) || make_slab_tmap(&tm_h0, gp.h0img, A_SLAB, h0_slabs, KSP)) {
        set_error("cuTensorMapEncodeTiled failed (pair2 GRU kernel)");
        return CCSM_ECUDA;
      }
      const int64_t max_clusters = T.sm_count / 2;
      if (duo) {
        const int64_t items = tiles / 2;
        const int clusters = (int)(items < max_clusters ? items : max_clusters);
        if (variant == 24)
          tc_gru_pair2_kernel<P, F16, C8, 1><<<2 * clusters, Pair2Cfg<P, 1>::THREADS, Pair2Cfg<P, 1>::SMEM, st>>>(
              gp, tm_w, tm_x, tm_o, tm_h0, tm_wx, tm_xs);
        else
          tc_gru_pair2_kernel<P, F16, C8, 2><<<2 * clusters, Pair2Cfg<P, 2>::THREADS, Pair2Cfg<P, 2>::SMEM, st>>>(
              gp, tm_w, tm_x, tm_o, tm_h0, tm_wx, tm_xs);
      } else {
        const int64_t items = tiles;  // (tiles / 2) pairs x 2 directions
        const int clusters = (int)(items < max_clusters ? items : max_clusters) & ~1;  // even: fixed direction per cluster
        tc_gru_pair2_kernel<P, F16, C8, 0><<<2 * clusters, Pair2Cfg<P, 0>::THREADS, Pair2Cfg<P, 0>::SMEM, st>>>(
            gp, tm_w, tm_x, tm_o, tm_h0, tm_wx, tm_xs);
      }
    } else if (variant == 16) {
      // duo kernel: CTA pairs, both directions interleaved; one cluster per TPC
      gp.wimg = T.wpair[l].as<uint8_t>();
      static bool duo_attr = false;
      if (!duo_attr) {
        CCSM_CUDA(cudaFuncSetAttribute(tc_gru_duo_kernel<P, F16, C8>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)DuoCfg<P>::SMEM));
        duo_attr = true;
      }
      const int64_t items = tiles / 2;
      const int64_t max_clusters = T.sm_count / 2;
      const int clusters = (int)(items < max_clusters ? items : max_clusters);
      tc_gru_duo_kernel<P, F16, C8><<<2 * clusters, DUO_THREADS, DuoCfg<P>::SMEM, st>>>(gp);
    } else if (!C8 && (variant == 3 || variant == 4)) {
      gp.wimg = T.wpair[l].as<uint8_t>();
      const int64_t items = tiles;  // (tiles / 2) pairs x 2 directions
      const int64_t max_clusters = (variant == 3 ? 1 : 2) * (T.sm_count / 2);
      const int clusters = (int)(items < max_clusters ? items : max_clusters) & ~1;  // even: fixed direction per cluster
      if constexpr (!C8) {
        if (variant == 3)
          tc_gru_pair_kernel<P, F16, 2><<<2 * clusters, PAIR_THREADS, PairCfg<P, 2>::SMEM, st>>>(gp);
        else
          tc_gru_pair_kernel<P, F16, 1><<<2 * clusters, PAIR_THREADS, PairCfg<P, 1>::SMEM, st>>>(gp);
      }
    } else {
      CCSM_TRY((launch_gru_variant<P, F16, C8>(variant, gp, tiles, T.sm_count, st)));
    }
    m->prof.end(pid, st);
    count_launch();
  }
  AttParams ap;
  ap.act = T.act[(NL - 1) % 3].as<uint8_t>();
  ap.wa_img = T.wa_img.as<uint8_t>();
  ap.ua_img = T.ua_img.as<uint8_t>();
  ap.va = T.va.as<float>();
  ap.fc_w = T.fc_w.as<float>();
  ap.fc_b = T.fc_b.as<float>();
  ap.logits = logits ? logits + site0 * 2 : nullptr;
  ap.probs = probs ? probs + site0 * 2 : nullptr;
  ap.n_tiles = (int)tiles;
  ap.L = L;
  ap.sites = sites;
  const int grid = (int)(tiles < T.sm_count ? tiles : T.sm_count);
  pid = m->prof.begin(PROF_ATT, (double)sites, st);
  tc_att_head_kernel<P, F16, C8><<<grid, ATT_THREADS, AttCfg<P>::SMEM, st>>>(ap);
  m->prof.end(pid, st);
  count_launch();
  CCSM_CUDA(cudaGetLastError());
  return CCSM_OK;
}

int tc_forward_att2s(ccsm_model* m, int64_t n, const ccsm_strand* fwd, const ccsm_strand* rev, const float* h0_f,
                     const float* h0_r, float* logits, float* probs, cudaStream_t st) {
  if (!m->tc || m->tc->P == 0) {
    set_error("tensor-core weights not packed");
    return CCSM_ESTATE;
  }
  TcState& T = *m->tc;
  // chunk = 8 items per SM: 2 slots x 148 SMs x 4 rounds of row tiles
  const int64_t max_tiles = (int64_t)T.sm_count * 8;
  const int64_t chunk_sites = max_tiles * (TILE_ROWS / 2);
  int64_t need_tiles = ((n < chunk_sites ? n : chunk_sites) * 2 + TILE_ROWS - 1) / TILE_ROWS;
  need_tiles += need_tiles & 1;
  CCSM_TRY(tc_reserve(m, need_tiles));
  for (int64_t s0 = 0; s0 < n; s0 += chunk_sites) {
    const int64_t sites = (n - s0) < chunk_sites ? (n - s0) : chunk_sites;
    int rc;
    if (T.c8) rc = tc_run_chunk<2, true, true>(m, sites, s0, n, fwd, rev, h0_f, h0_r, logits, probs, st);
    else if (T.P == 1 && !T.f16) rc = tc_run_chunk<1, false>(m, sites, s0, n, fwd, rev, h0_f, h0_r, logits, probs, st);
    else if (T.P == 1 && T.f16) rc = tc_run_chunk<1, true>(m, sites, s0, n, fwd, rev, h0_f, h0_r, logits, probs, st);
    else if (T.P == 2 && !T.f16) rc = tc_run_chunk<2, false>(m, sites, s0, n, fwd, rev, h0_f, h0_r, logits, probs, st);
    else rc = tc_run_chunk<2, true>(m, sites, s0, n, fwd, rev, h0_f, h0_r, logits, probs, st);
    CCSM_TRY(rc);
  }
  return CCSM_OK;
}

int tc_debug_layer_out(ccsm_model* m, int layer, float* host, int64_t cap, int64_t* written) {
  if (!m->tc || m->tc->last_tiles == 0) {
    set_error("tc debug: no forward recorded");
    return CCSM_ESTATE;
  }
  TcState& T = *m->tc;
  const int L = m->cfg.seq_len;
  const int64_t rows = T.last_tiles * TILE_ROWS;
  int64_t nfl = rows * L * 512;
  DevBuf tmp;
  CCSM_TRY(tmp.reserve((size_t)nfl * 4));
  const int64_t total = rows * L * 64;
  const unsigned blocks = (unsigned)((total + 255) / 256);
  const uint8_t* act = T.act[layer % 3].as<uint8_t>();
  CCSM_CUDA(cudaDeviceSynchronize());
  if (T.c8) tc_unpack_act_kernel<2, true, true><<<blocks, 256>>>(act, tmp.as<float>(), rows, L);
  else if (T.P == 1 && !T.f16) tc_unpack_act_kernel<1, false><<<blocks, 256>>>(act, tmp.as<float>(), rows, L);
  else if (T.P == 1 && T.f16) tc_unpack_act_kernel<1, true><<<blocks, 256>>>(act, tmp.as<float>(), rows, L);
  else if (T.P == 2 && !T.f16) tc_unpack_act_kernel<2, false><<<blocks, 256>>>(act, tmp.as<float>(), rows, L);
  else tc_unpack_act_kernel<2, true><<<blocks, 256>>>(act, tmp.as<float>(), rows, L);
  count_launch();
  CCSM_CUDA(cudaDeviceSynchronize());
  if (nfl > cap) nfl = cap;
  CCSM_CUDA(cudaMemcpy(host, tmp.p, (size_t)nfl * 4, cudaMemcpyDeviceToHost));
  tmp.release();
  *written = nfl;
  return CCSM_OK;
}

}  // namespace ccsm
