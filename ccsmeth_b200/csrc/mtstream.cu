// Device reproduction of the reference's GRU initial-state stream.
//
// Replaces: reference ccsmeth/models.py:77-87 (ModelAttRNN.init_hidden = torch.randn(2*layers, n, hidden) on the CPU
// default generator, one draw per strand per model call) seeded by call_modifications.py:479-481 (torch.manual_seed).
//
// What ATen does for that draw (aten/src/ATen/native/cpu/DistributionTemplates.h; the kernel every AVX2-capable x86
// host dispatches to -- the AVX512 slot of this stub is empty, so AVX512 hosts run it too):
//   * engine at::mt19937 = MT19937 seeded like init_genrand from the low 32 seed bits, one 32-bit output per element;
//   * uniform u = (x & (2^24 - 1)) * 2^-24;
//   * normal_fill_AVX2 on groups of 16 consecutive elements: for j < 8, u1 = 1 - u[j], u2 = u[j + 8],
//       radius = sqrt(-2 * log256_ps(u1)), theta = float(2 pi) * u2, out[j] = radius * cos, out[j + 8] = radius * sin
//     with the single-precision Cephes polynomials of avx_mathfun.h, whose multiply-adds the compiler contracted into
//     FMAs (pattern pinned bit-for-bit against torch.randn by tests/test_h0_stream_gpu.py).
// Here: mt_generate_kernel produces the 32-bit generator words (one CTA: the twist is a sequential recurrence over
// 624-word blocks, 227-wide inside a block), mt_normal_kernel turns them into the (2*layers, n, hidden) fp32 tensors the
// forward takes as explicit h0 -- same values, same order, nothing crosses PCIe.  All float arithmetic below is written
// with round-to-nearest intrinsics so that nvcc neither contracts nor reassociates it.
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <mutex>
#include <vector>

#include "ccsm_internal.h"
#include "mtjump.h"

namespace ccsm {

constexpr int MT_N = 624, MT_M = 397;

__device__ __forceinline__ uint32_t mt_tw(uint32_t u, uint32_t v) {
  const uint32_t y = (u & 0x80000000u) | (v & 0x7fffffffu);
  return (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
}
__device__ __forceinline__ uint32_t mt_temper(uint32_t y) {
  y ^= y >> 11;
  y ^= (y << 7) & 0x9d2c5680u;
  y ^= (y << 15) & 0xefc60000u;
  y ^= y >> 18;
  return y;
}

// state[0..623] = generator words (a twisted block), state[624] = position of the next output inside it (624 = the
// block is used up).  Writes the next n tempered outputs to out and leaves the generator exactly where ATen's would be.
//
// The twist is the linear recurrence x[k] = x[k - 227] ^ T(k), T(k) = tw(x[k - 624], x[k - 623]), over the output
// sequence: only 227 consecutive words are independent of each other.  Substituting the recurrence into itself twice,
//     x[k] = x[k - 681] ^ T(k) ^ T(k - 227) ^ T(k - 454),
// every operand lies at least 623 words back, so 623 consecutive words are independent: one block-wide barrier per 623
// words (one word per thread, three tw() instead of one) instead of one per 227.  The first 454 words after the state
// block are produced with the plain recurrence (two steps) to build the 1078 words of history the long form reads.
constexpr int MT_LAG = MT_N - MT_M;  // 227
constexpr int MT_WIDE = MT_N - 1;    // 623 words per step
constexpr int MT_HIST = MT_N + 2 * MT_LAG;  // 1078 words of history a wide step reads
constexpr int MT_BUF = 8192;         // linear window: slides back to the start every 11 steps, so addresses need no masking
constexpr int MT_THREADS = 640;

// Emits n UNTEMPERED words (the consumers temper on load: they run on every SM, the generator on one) starting at
// position `pos` of the 624-word window in x[0..624), x = a shared-memory buffer of MT_BUF words.  A window whose first
// word only carries its top bit -- a state vector, as a jump produces it -- is emitted with pos = 1.  If state_out is
// given, the block holding the last output and the position inside it are stored there (ATen's representation).
__device__ void mt_emit(uint32_t* x, int pos, long long n, uint32_t* __restrict__ out, uint32_t* __restrict__ state_out) {
  const int tid = threadIdx.x;
  if (n <= 0) return;
  // outputs still owed by the current block
  const long long head = (long long)(MT_N - pos) < n ? (MT_N - pos) : n;
  for (int i = tid; i < head; i += MT_THREADS) out[i] = x[pos + i];
  // stream index (from the start of the current block) one past the last output, and the block that holds that output
  const long long p_end = pos + n;
  const long long blk = (p_end - 1) / MT_N;          // >= 0
  const long long k_stop = (blk + 1) * MT_N;         // generate x[624 .. k_stop)
  long long kbase = 0;  // stream index of x[0]
  if (blk > 0) {
    // two plain steps: x[624 .. 1078)
    for (int m = 0; m < 2; ++m) {
      const int k = MT_N + m * MT_LAG + tid;
      if (tid < MT_LAG) {
        const uint32_t v = x[k - MT_LAG] ^ mt_tw(x[k - MT_N], x[k - MT_N + 1]);
        x[k] = v;
        const long long o = (long long)k - pos;
        if (o < n) out[o] = v;
      }
      __syncthreads();
    }
    // wide steps: stream words k0 + tid, k0 = 1078 + 623 q, at x[w + tid]
    const long long k_first = MT_HIST;
    const long long n_wide = k_stop > k_first ? (k_stop - k_first + MT_WIDE - 1) / MT_WIDE : 0;
    const long long o_first = k_first - pos;
    long long n_full = (n - o_first) / MT_WIDE;  // steps whose 623 words are all outputs: no bounds check inside
    n_full = n_full < 0 ? 0 : (n_full > n_wide ? n_wide : n_full);
    uint32_t* op = out + (o_first + tid);
    const bool act = tid < MT_WIDE;
    int w = MT_HIST;  // x[] index of the step's first word
    auto wide = [&](const uint32_t* c) {  // c = &x[index of this thread's word]
      return c[-3 * MT_LAG] ^ mt_tw(c[-MT_N], c[-MT_N + 1]) ^ mt_tw(c[-MT_N - MT_LAG], c[-MT_N - MT_LAG + 1]) ^
             mt_tw(c[-MT_N - 2 * MT_LAG], c[-MT_N - 2 * MT_LAG + 1]);
    };
    for (long long q = 0; q < n_wide; ++q) {
      if (w + MT_WIDE > MT_BUF) {
        // slide the last 1078 words back to the start of the window
        uint32_t keep[2];
#pragma unroll
        for (int i = 0; i < 2; ++i) keep[i] = (tid + i * MT_THREADS < MT_HIST) ? x[w - MT_HIST + tid + i * MT_THREADS] : 0u;
        __syncthreads();
#pragma unroll
        for (int i = 0; i < 2; ++i)
          if (tid + i * MT_THREADS < MT_HIST) x[tid + i * MT_THREADS] = keep[i];
        __syncthreads();
        kbase += w - MT_HIST;
        w = MT_HIST;
      }
      if (act) {
        uint32_t* c = x + w + tid;
        const uint32_t v = wide(c);
        *c = v;
        if (q < n_full || o_first + q * MT_WIDE + tid < n) *op = v;
      }
      w += MT_WIDE;
      op += MT_WIDE;
      __syncthreads();
    }
  }
  if (state_out) {
    // the block holding the last output becomes the state; its position = how far into it the stream has advanced
    const long long b0 = blk * MT_N;
    for (int i = tid; i < MT_N; i += MT_THREADS) state_out[i] = x[(int)(b0 - kbase) + i];
    if (tid == 0) state_out[MT_N] = (uint32_t)(p_end - b0);
  }
}

// state[0..623] = generator words (a twisted block), state[624] = position of the next output inside it (624 = the block
// is used up).  Writes the next n generator words to out and leaves the generator exactly where ATen's would be.
__global__ void __launch_bounds__(MT_THREADS, 1) mt_generate_kernel(uint32_t* __restrict__ state, uint32_t* __restrict__ out,
                                                                    long long n) {
  __shared__ uint32_t x[MT_BUF];
  for (int i = threadIdx.x; i < MT_N; i += MT_THREADS) x[i] = state[i];
  const int pos = (int)state[MT_N];
  __syncthreads();
  mt_emit(x, pos, n, out, state);
}

// ---- sub-streams by jump-ahead: the stream of one call cut into pieces of MT_J words, one CTA each.
// The state J words ahead is a fixed GF(2)-linear function of the state: with g(x) = x^J mod phi(x) (phi = the
// characteristic polynomial of the one-word transition, degree 19937; host code in mtjump.h), the state vector at word
// b + J is the XOR, over the set bits i of g, of the state vectors at words b + i.  CTA c applies the polynomials of
// J, 2 J, 4 J, ... selected by the bits of c: each jump lays out the next 19,937 + 623 words after its window in shared
// memory and XORs about ten thousand shifted views of them.  A state vector is (top bit of x[b], x[b + 1 .. b + 623]),
// so CTA c jumps from the word BEFORE the call's first output and emits from position 1 of the resulting window.
constexpr int MT_JLOG = 21;
constexpr long long MT_J = 1LL << MT_JLOG;   // words per sub-stream
constexpr int MT_JMAXBITS = 7;               // up to 128 sub-streams per launch
constexpr int MT_DEG = 19937;
constexpr int MT_JLIST = 12288;              // capacity of a polynomial's set-bit list (about 10,000 of 19,937 bits are set)
constexpr int MT_SEQ = 2 * MT_N + MT_DEG + MT_WIDE + 32;  // words a jump lays out: window offset < 624, + 19,937 + window

__global__ void __launch_bounds__(MT_THREADS, 1) mt_generate_multi_kernel(const uint32_t* __restrict__ state,
                                                                          uint32_t* __restrict__ out, long long n,
                                                                          const uint16_t* __restrict__ jlist,
                                                                          const int* __restrict__ jcount) {
  extern __shared__ uint32_t sm[];  // MT_SEQ words (>= MT_BUF)
  const int tid = threadIdx.x;
  const int c = blockIdx.x;
  for (int i = tid; i < MT_N; i += MT_THREADS) sm[i] = state[i];
  int pos = (int)state[MT_N];
  __syncthreads();
  if (c > 0) {
    int b = pos - 1;  // the window (state vector) starts at sm[b]
    for (int k = 0; k < MT_JMAXBITS; ++k) {
      if (!((c >> k) & 1)) continue;
      // lay out sm[624 + ..] up to b + 624 + 19,937: every word from index 624 on follows from the 624 before it
      const int stop = b + MT_N + MT_DEG;
      for (int k0 = MT_N; k0 < stop; k0 += MT_LAG) {   // plain 227-wide steps (20 k words: a few tens of microseconds)
        const int kk = k0 + tid;
        if (tid < MT_LAG && kk < stop) sm[kk] = sm[kk - MT_LAG] ^ mt_tw(sm[kk - MT_N], sm[kk - MT_N + 1]);
        __syncthreads();
      }
      uint32_t acc = 0;
      if (tid < MT_N) {
        const uint16_t* lst = jlist + (size_t)k * MT_JLIST;
        const int cnt = jcount[k];
        const uint32_t* base = sm + b + tid;
        for (int i = 0; i < cnt; ++i) acc ^= base[lst[i]];
      }
      __syncthreads();
      if (tid < MT_N) sm[tid] = acc;   // the jumped state vector; only the top bit of its first word is meaningful
      __syncthreads();
      b = 0;
    }
    pos = 1;
    if (b != 0) {  // (c > 0 always jumps at least once, so b == 0 here; kept for clarity)
      return;
    }
  }
  const long long first = (long long)c * MT_J;
  const long long len = (n - first) < MT_J ? (n - first) : MT_J;
  mt_emit(sm, pos, len, out + first, nullptr);
}

// After the sub-streams: the generator state ATen would hold, from the last 624 words written (n >= 624).
__global__ void __launch_bounds__(256, 1) mt_finalize_kernel(uint32_t* __restrict__ state, const uint32_t* __restrict__ out,
                                                             long long n) {
  __shared__ uint32_t x[2 * MT_N];
  const int tid = threadIdx.x;
  const long long e = (long long)state[MT_N] + n - 1;   // index of the last output, counted from the old block's start
  const long long blk = e / MT_N;
  const int have = (int)(e - blk * MT_N) + 1;           // words of block `blk` that were output: its first `have` words
  // x[0..624) = the 624 words ending at the last output (index e - 623 .. e)
  for (int i = tid; i < MT_N; i += 256) x[i] = out[n - MT_N + i];
  __syncthreads();
  // the rest of the block: words e + 1 .. (blk + 1) * 624 - 1 at x[624 ..]
  const int more = MT_N - have;
  for (int k0 = 0; k0 < more; k0 += MT_LAG) {
    const int kk = MT_N + k0 + tid;
    if (tid < MT_LAG && k0 + tid < more) x[kk] = x[kk - MT_LAG] ^ mt_tw(x[kk - MT_N], x[kk - MT_N + 1]);
    __syncthreads();
  }
  // block word i sits at x[624 - have + i]
  for (int i = tid; i < MT_N; i += 256) state[i] = x[MT_N - have + i];
  if (tid == 0) state[MT_N] = (uint32_t)have;
}

// Jump polynomials x^(2^k J) mod phi, k < MT_JMAXBITS, as lists of set-bit indices; computed once per process
// (Berlekamp-Massey + a few polynomial squarings, ~0.1 s) and uploaded once per device.
struct MtJumpTables {
  std::vector<uint16_t> lists;  // [MT_JMAXBITS][MT_JLIST]
  std::vector<int> counts;
  bool ok = false;
};
static const MtJumpTables& mt_jump_tables() {
  static MtJumpTables T;
  static std::once_flag once;
  std::call_once(once, [] {
    mtjump::Field F = mtjump::make_field();
    int deg = -1;
    for (int i = (int)F.phi.size() * 64 - 1; i >= 0 && deg < 0; --i)
      if (mtjump::get(F.phi, i)) deg = i;
    if (deg != MT_DEG) return;  // never expected; the callers fall back to the one-CTA generator
    T.lists.assign((size_t)MT_JMAXBITS * MT_JLIST, 0);
    T.counts.assign(MT_JMAXBITS, 0);
    mtjump::Poly g = F.xpow((uint64_t)MT_J);
    for (int k = 0; k < MT_JMAXBITS; ++k) {
      int c = 0;
      for (int i = 0; i < MT_DEG; ++i)
        if (mtjump::get(g, i)) {
          if (c >= MT_JLIST) return;
          T.lists[(size_t)k * MT_JLIST + c++] = (uint16_t)i;
        }
      T.counts[k] = c;
      if (k + 1 < MT_JMAXBITS) g = F.mulmod(g, g);
    }
    T.ok = true;
  });
  return T;
}
struct MtJumpDev {
  DevBuf lists, counts;
  bool ready = false, failed = false;
};
static MtJumpDev* mt_jump_dev() {
  static MtJumpDev dev[64];
  static std::mutex mu;
  int d = 0;
  if (cudaGetDevice(&d) != cudaSuccess || d < 0 || d >= 64) return nullptr;
  std::lock_guard<std::mutex> lk(mu);
  MtJumpDev& D = dev[d];
  if (D.failed) return nullptr;
  if (!D.ready) {
    const MtJumpTables& T = mt_jump_tables();
    if (!T.ok || D.lists.reserve(T.lists.size() * 2) != CCSM_OK || D.counts.reserve(T.counts.size() * 4) != CCSM_OK ||
        cudaMemcpy(D.lists.p, T.lists.data(), T.lists.size() * 2, cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMemcpy(D.counts.p, T.counts.data(), T.counts.size() * 4, cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaFuncSetAttribute(mt_generate_multi_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, MT_SEQ * 4) != cudaSuccess) {
      D.failed = true;
      return nullptr;
    }
    D.ready = true;
  }
  return &D;
}

// Advances the generator by n words into out.  Short draws run on one CTA; from two sub-streams' worth on, the draw is
// cut into sub-streams of MT_J words generated in parallel (up to 128 per launch), followed by the state update.
// pos_may_be_zero: the state was just set by the caller with position 0 (a state vector needs the word before the first
// output, which such a block does not hold): one-CTA path.
static void mt_launch_generate(uint32_t* state, uint32_t* out, long long n, cudaStream_t st, bool pos_may_be_zero = false) {
  static int force_serial = -1;
  if (force_serial < 0) {
    const char* e = getenv("CCSM_MT_SERIAL");
    force_serial = (e && atoi(e) != 0) ? 1 : 0;
  }
  MtJumpDev* J = (n >= 2 * MT_J && !pos_may_be_zero && !force_serial) ? mt_jump_dev() : nullptr;
  if (!J) {
    mt_generate_kernel<<<1, MT_THREADS, 0, st>>>(state, out, n);
    return;
  }
  const long long per_launch = MT_J << MT_JMAXBITS;
  for (long long o = 0; o < n; o += per_launch) {
    const long long m = (n - o) < per_launch ? (n - o) : per_launch;
    if (m < MT_N) {  // (cannot happen for the sizes that reach this path; the state update needs 624 written words)
      mt_generate_kernel<<<1, MT_THREADS, 0, st>>>(state, out + o, m);
      continue;
    }
    const int ctas = (int)((m + MT_J - 1) / MT_J);
    mt_generate_multi_kernel<<<ctas, MT_THREADS, MT_SEQ * 4, st>>>(state, out + o, m, J->lists.as<uint16_t>(), J->counts.as<int>());
    mt_finalize_kernel<<<1, 256, 0, st>>>(state, out + o, m);
  }
}

// ---- avx_mathfun.h log256_ps / sincos256_ps, one lane, with the FMA contractions of the shipped ATen binary
__device__ __forceinline__ float avx_log(float x) {
  x = fmaxf(x, 1.17549435e-38f);
  int e_i = (int)(__float_as_uint(x) >> 23) - 0x7f;
  x = __uint_as_float((__float_as_uint(x) & ~0x7f800000u) | 0x3f000000u);
  float e = __fadd_rn((float)e_i, 1.0f);
  const bool lt = x < 0.707106781186547524f;
  const float tmp = lt ? x : 0.0f;
  x = __fsub_rn(x, 1.0f);
  e = __fsub_rn(e, lt ? 1.0f : 0.0f);
  x = __fadd_rn(x, tmp);
  const float z = __fmul_rn(x, x);
  float y = 7.0376836292E-2f;
  y = __fmaf_rn(y, x, -1.1514610310E-1f);
  y = __fmaf_rn(y, x, 1.1676998740E-1f);
  y = __fmaf_rn(y, x, -1.2420140846E-1f);
  y = __fmaf_rn(y, x, 1.4249322787E-1f);
  y = __fmaf_rn(y, x, -1.6668057665E-1f);
  y = __fmaf_rn(y, x, 2.0000714765E-1f);
  y = __fmaf_rn(y, x, -2.4999993993E-1f);
  y = __fmaf_rn(y, x, 3.3333331174E-1f);
  y = __fmul_rn(y, x);
  y = __fmaf_rn(y, z, __fmul_rn(e, -2.12194440e-4f));   // (y * z) fused with the add; e * q1 rounded on its own
  y = __fmaf_rn(-z, 0.5f, y);
  x = __fadd_rn(x, y);
  return __fmaf_rn(e, 0.693359375f, x);
}
__device__ __forceinline__ void avx_sincos(float x, float& s, float& c) {
  uint32_t sign_sin = __float_as_uint(x) & 0x80000000u;
  x = fabsf(x);
  float y = __fmul_rn(x, 1.27323954473516f);
  int j = __float2int_rz(y);
  j = (j + 1) & ~1;
  y = (float)j;
  const uint32_t swap_sin = ((uint32_t)(j & 4)) << 29;
  const bool poly = (j & 2) == 0;
  x = __fmaf_rn(y, -0.78515625f, x);
  x = __fmaf_rn(y, -2.4187564849853515625e-4f, x);
  x = __fmaf_rn(y, -3.77489497744594108e-8f, x);
  const uint32_t sign_cos = ((uint32_t)(~(j - 2) & 4)) << 29;
  sign_sin ^= swap_sin;
  const float z = __fmul_rn(x, x);
  float yc = 2.443315711809948E-005f;
  yc = __fmaf_rn(yc, z, -1.388731625493765E-003f);
  yc = __fmaf_rn(yc, z, 4.166664568298827E-002f);
  yc = __fmul_rn(yc, z);
  yc = __fmaf_rn(yc, z, -__fmul_rn(z, 0.5f));
  yc = __fadd_rn(yc, 1.0f);
  float ys = -1.9515295891E-4f;
  ys = __fmaf_rn(ys, z, 8.3321608736E-3f);
  ys = __fmaf_rn(ys, z, -1.6666654611E-1f);
  ys = __fmul_rn(ys, z);
  ys = __fmaf_rn(ys, x, x);
  s = __uint_as_float(__float_as_uint(poly ? ys : yc) ^ sign_sin);
  c = __uint_as_float(__float_as_uint(poly ? yc : ys) ^ sign_cos);
}
// one Box-Muller pair of normal_fill_16: raw outputs a (-> u1) and b (-> u2)
__device__ __forceinline__ void normal_pair(uint32_t a, uint32_t b, float& n_cos, float& n_sin) {
  const float u1 = __fsub_rn(1.0f, __fmul_rn((float)(a & 0xffffffu), 5.9604644775390625e-8f));  // exact
  const float u2 = __fmul_rn((float)(b & 0xffffffu), 5.9604644775390625e-8f);
  const float radius = __fsqrt_rn(__fmul_rn(-2.0f, avx_log(u1)));
  float s, c;
  avx_sincos(__fmul_rn(6.283185307179586f, u2), s, c);
  // ATen finishes with fmadd(n, std = 1, mean = +0): exact, except that it turns a -0 (u1 == 1) into +0
  n_cos = __fmaf_rn(__fmul_rn(radius, c), 1.0f, 0.0f);
  n_sin = __fmaf_rn(__fmul_rn(radius, s), 1.0f, 0.0f);
}

// Plain stream: out[16 g + j] / out[16 g + 8 + j] for every group g of 16 words (debug / tests).
__global__ void mt_normal_flat_kernel(const uint32_t* __restrict__ words, float* __restrict__ out, long long groups) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;  // one pair per thread
  if (t >= groups * 8) return;
  const long long g = t >> 3;
  const int j = (int)(t & 7);
  float nc, ns;
  normal_pair(mt_temper(words[g * 16 + j]), mt_temper(words[g * 16 + 8 + j]), nc, ns);
  out[g * 16 + j] = nc;
  out[g * 16 + 8 + j] = ns;
}

// The reference's layout: model call k (a "segment" of n_k <= batch sites starting at site s_k) draws
// randn(LD, n_k, H) for strand 1, then for strand 2; word base w_k = sum_{k' < k} 2 LD n_k' H.
// seg_site[k] = s_k (seg_site[nseg] = n), seg_word[k] = w_k.  h0a / h0b: (LD, n, H) fp32.
__global__ void mt_normal_h0_kernel(const uint32_t* __restrict__ words, const long long* __restrict__ seg_site,
                                    const long long* __restrict__ seg_word, int nseg, long long n, int LD, int H,
                                    float* __restrict__ h0a, float* __restrict__ h0b) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;  // one pair per thread
  const long long total_pairs = n * 2 * LD * H / 2;
  if (t >= total_pairs) return;
  const long long g = t >> 3;
  const int j = (int)(t & 7);
  const long long w = g * 16;  // stream position of the group's first word
  int lo = 0, hi = nseg - 1;   // last segment whose word base is <= w
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (seg_word[mid] <= w) lo = mid;
    else hi = mid - 1;
  }
  const long long s0 = seg_site[lo], nk = seg_site[lo + 1] - s0;
  long long rel = w - seg_word[lo];
  const long long per_strand = (long long)LD * nk * H;
  const int strand = rel >= per_strand;
  rel -= strand * per_strand;
  const long long ld = rel / (nk * H);
  rel -= ld * nk * H;
  const long long site = s0 + rel / H;
  const int unit = (int)(rel % H);
  float nc, ns;
  normal_pair(mt_temper(words[w + j]), mt_temper(words[w + 8 + j]), nc, ns);
  float* dst = (strand ? h0b : h0a) + (ld * n + site) * H + unit;
  dst[j] = nc;
  dst[8 + j] = ns;
}

struct MtStream {
  DevBuf state;   // 624 words + position
  bool seeded = false;
  bool pos_zero = false;   // the state was set with position 0 and nothing has been drawn since
  uint64_t seed = 0;
  int batch = 512;                 // the reference's --batch_size: sites per model call
  std::vector<int64_t> pending;    // hole-batch site counts announced for the next forward call
  DevBuf words[2], h0[2], segtab[2];
  int next_buf = 0;
  cudaStream_t st = nullptr;
  cudaEvent_t ready[2] = {nullptr, nullptr}, freed[2] = {nullptr, nullptr};
  bool used[2] = {false, false};
};

static int mt_ensure(ccsm_model* m) {
  if (!m->mts) m->mts = new MtStream();
  MtStream& S = *m->mts;
  if (!S.st) {
    CCSM_CUDA(cudaStreamCreateWithFlags(&S.st, cudaStreamNonBlocking));
    for (int i = 0; i < 2; ++i) {
      CCSM_CUDA(cudaEventCreateWithFlags(&S.ready[i], cudaEventDisableTiming));
      CCSM_CUDA(cudaEventCreateWithFlags(&S.freed[i], cudaEventDisableTiming));
    }
  }
  return CCSM_OK;
}

int mt_seed(ccsm_model* m, uint64_t seed) {
  CCSM_TRY(mt_ensure(m));
  MtStream& S = *m->mts;
  uint32_t st[MT_N + 1];
  st[0] = (uint32_t)(seed & 0xffffffffu);
  for (int j = 1; j < MT_N; ++j) st[j] = 1812433253u * (st[j - 1] ^ (st[j - 1] >> 30)) + (uint32_t)j;
  st[MT_N] = MT_N;  // the first output twists
  CCSM_TRY(S.state.reserve(sizeof(st)));
  CCSM_CUDA(cudaStreamSynchronize(S.st));
  CCSM_CUDA(cudaMemcpy(S.state.p, st, sizeof(st), cudaMemcpyHostToDevice));
  S.seeded = true;
  S.pos_zero = false;
  S.seed = seed;
  S.pending.clear();
  return CCSM_OK;
}

int mt_set_state(ccsm_model* m, const uint32_t* words, int32_t pos) {
  CCSM_TRY(mt_ensure(m));
  MtStream& S = *m->mts;
  if (pos < 0 || pos > MT_N) {
    set_error("h0 stream: position %d outside [0, 624]", pos);
    return CCSM_EINVAL;
  }
  uint32_t st[MT_N + 1];
  for (int i = 0; i < MT_N; ++i) st[i] = words[i];
  st[MT_N] = (uint32_t)pos;
  CCSM_TRY(S.state.reserve(sizeof(st)));
  CCSM_CUDA(cudaStreamSynchronize(S.st));
  CCSM_CUDA(cudaMemcpy(S.state.p, st, sizeof(st), cudaMemcpyHostToDevice));
  S.seeded = true;
  S.pos_zero = pos == 0;
  return CCSM_OK;
}

int mt_get_state(ccsm_model* m, uint32_t* words, int32_t* pos) {
  CCSM_TRY(mt_ensure(m));
  MtStream& S = *m->mts;
  if (!S.seeded) CCSM_TRY(mt_seed(m, m->h0_seed));
  uint32_t st[MT_N + 1];
  CCSM_CUDA(cudaStreamSynchronize(S.st));
  CCSM_CUDA(cudaMemcpy(st, S.state.p, sizeof(st), cudaMemcpyDeviceToHost));
  for (int i = 0; i < MT_N; ++i) words[i] = st[i];
  *pos = (int32_t)st[MT_N];
  return CCSM_OK;
}

int mt_set_batching(ccsm_model* m, const int64_t* counts, int64_t n_counts, int32_t batch_size) {
  CCSM_TRY(mt_ensure(m));
  MtStream& S = *m->mts;
  if (batch_size <= 0) {
    set_error("h0 stream: batch_size must be positive");
    return CCSM_EINVAL;
  }
  S.batch = batch_size;
  S.pending.assign(counts, counts + n_counts);
  return CCSM_OK;
}

// The model calls ("segments") the reference would make for the next n sites: every announced hole-batch is cut
// into slices of <= batch sites (reference call_modifications.py:177-181); without an announcement the n sites are one
// hole-batch.  Consumes the announcement.
int mt_take_segments(ccsm_model* m, int64_t n, std::vector<int64_t>& segs) {
  CCSM_TRY(mt_ensure(m));
  MtStream& S = *m->mts;
  if (!S.seeded) CCSM_TRY(mt_seed(m, m->h0_seed));
  std::vector<int64_t> hb;
  hb.swap(S.pending);
  if (hb.empty()) hb.push_back(n);
  int64_t tot = 0;
  segs.clear();
  for (int64_t c : hb) {
    if (c < 0) {
      set_error("h0 stream: negative hole-batch count");
      return CCSM_EINVAL;
    }
    tot += c;
    for (int64_t s = 0; s < c; s += S.batch) segs.push_back((c - s) < S.batch ? (c - s) : S.batch);
  }
  if (tot != n) {
    set_error("h0 stream: the announced hole-batches hold %lld sites, the call has %lld", (long long)tot, (long long)n);
    return CCSM_EINVAL;
  }
  return CCSM_OK;
}

// Draws the h0 of `nseg` consecutive model calls (site counts segs[]) on the generator's own stream into one of two
// buffers; `user` (the stream the forward runs on) is made to wait for it.  *h0a / *h0b: (LD, n, H) device tensors,
// valid until the second-next call; the caller records `freed` via mt_release after launching the consumer.
int mt_fill(ccsm_model* m, const int64_t* segs, int nseg, cudaStream_t user, const float** h0a, const float** h0b, int* buf) {
  MtStream& S = *m->mts;
  const int LD = 2 * m->cfg.num_layers, H = m->cfg.hidden;
  if (H % 16 != 0) {
    set_error("h0 stream: hidden size must be a multiple of 16");
    return CCSM_EUNSUPPORTED;
  }
  std::vector<long long> tab(2 * (size_t)(nseg + 1));
  long long n = 0, w = 0;
  for (int k = 0; k < nseg; ++k) {
    tab[k] = n;
    tab[nseg + 1 + k] = w;
    n += segs[k];
    w += 2LL * LD * segs[k] * H;
  }
  tab[nseg] = n;
  tab[2 * nseg + 1] = w;
  const int b = S.next_buf;
  S.next_buf ^= 1;
  if (S.used[b]) CCSM_CUDA(cudaStreamWaitEvent(S.st, S.freed[b], 0));  // the forward that read this buffer is done
  CCSM_TRY(S.words[b].reserve((size_t)w * 4));
  CCSM_TRY(S.h0[b].reserve((size_t)w * 4));
  CCSM_TRY(S.segtab[b].reserve(tab.size() * 8));
  // pageable source: the call returns once the table sits in the driver's staging memory, so `tab` may die afterwards
  CCSM_CUDA(cudaMemcpyAsync(S.segtab[b].p, tab.data(), tab.size() * 8, cudaMemcpyHostToDevice, S.st));
  mt_launch_generate(S.state.as<uint32_t>(), S.words[b].as<uint32_t>(), w, S.st, S.pos_zero);
  if (w > 0) S.pos_zero = false;
  float* a = S.h0[b].as<float>();
  float* bb = a + (size_t)LD * n * H;
  const long long pairs = w / 2;
  mt_normal_h0_kernel<<<(unsigned)((pairs + 255) / 256), 256, 0, S.st>>>(
      S.words[b].as<uint32_t>(), S.segtab[b].as<long long>(), S.segtab[b].as<long long>() + nseg + 1, nseg, n, LD, H, a, bb);
  count_launch(2);
  CCSM_CUDA(cudaGetLastError());
  CCSM_CUDA(cudaEventRecord(S.ready[b], S.st));
  CCSM_CUDA(cudaStreamWaitEvent(user, S.ready[b], 0));
  S.used[b] = true;
  *h0a = a;
  *h0b = bb;
  *buf = b;
  return CCSM_OK;
}

int mt_release(ccsm_model* m, int buf, cudaStream_t user) {
  CCSM_CUDA(cudaEventRecord(m->mts->freed[buf], user));
  return CCSM_OK;
}

void mt_destroy(ccsm_model* m) {
  if (!m->mts) return;
  MtStream& S = *m->mts;
  S.state.release();
  for (int i = 0; i < 2; ++i) {
    S.words[i].release();
    S.h0[i].release();
    S.segtab[i].release();
    if (S.ready[i]) cudaEventDestroy(S.ready[i]);
    if (S.freed[i]) cudaEventDestroy(S.freed[i]);
  }
  if (S.st) cudaStreamDestroy(S.st);
  delete m->mts;
  m->mts = nullptr;
}

}  // namespace ccsm

using namespace ccsm;

// torch.manual_seed(seed); torch.randn(skip); torch.randn(n) -> out (host, n floats; skip and n multiples of 16).
extern "C" int ccsm_debug_torch_randn(int32_t device, uint64_t seed, int64_t skip, int64_t n, float* out) {
  if (!out || n <= 0 || n % 16 || skip < 0 || skip % 16) {
    set_error("ccsm_debug_torch_randn: n and skip must be multiples of 16");
    return CCSM_EINVAL;
  }
  CCSM_CUDA(cudaSetDevice(device));
  uint32_t st[MT_N + 1];
  st[0] = (uint32_t)(seed & 0xffffffffu);
  for (int j = 1; j < MT_N; ++j) st[j] = 1812433253u * (st[j - 1] ^ (st[j - 1] >> 30)) + (uint32_t)j;
  st[MT_N] = MT_N;
  DevBuf ds, dw, df;
  CCSM_TRY(ds.reserve(sizeof(st)));
  CCSM_TRY(dw.reserve((size_t)(skip > n ? skip : n) * 4));
  CCSM_TRY(df.reserve((size_t)n * 4));
  CCSM_CUDA(cudaMemcpy(ds.p, st, sizeof(st), cudaMemcpyHostToDevice));
  if (skip) mt_launch_generate(ds.as<uint32_t>(), dw.as<uint32_t>(), skip, nullptr);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  cudaEventRecord(e0);
  mt_launch_generate(ds.as<uint32_t>(), dw.as<uint32_t>(), n, nullptr);
  cudaEventRecord(e1);
  mt_normal_flat_kernel<<<(unsigned)((n / 2 + 255) / 256), 256>>>(dw.as<uint32_t>(), df.as<float>(), n / 16);
  count_launch(3);
  cudaError_t e = cudaDeviceSynchronize();
  if (e == cudaSuccess && getenv("CCSM_MT_TIMING")) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    fprintf(stderr, "[mt] %lld words in %.3f ms = %.2f G words/s\n", (long long)n, ms, n / (ms * 1e6));
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  if (e == cudaSuccess) e = cudaMemcpy(out, df.p, (size_t)n * 4, cudaMemcpyDeviceToHost);
  ds.release(); dw.release(); df.release();
  if (e != cudaSuccess) {
    set_error("ccsm_debug_torch_randn: %s", cudaGetErrorString(e));
    return CCSM_ECUDA;
  }
  return CCSM_OK;
}

// Host-only self-check of the jump polynomials: the state vector J_k = 2^k J words ahead of word b, computed as the XOR
// of the state vectors i words ahead over the set bits i of x^(J_k) mod phi, against plain generation.  Returns the number
// of mismatching words (0 = pass), or a negative error code.  No GPU needed.
extern "C" int ccsm_debug_mt_jump_check(uint32_t seed, int32_t k) {
  if (k < 0 || k >= MT_JMAXBITS) {
    set_error("ccsm_debug_mt_jump_check: k outside [0, %d)", MT_JMAXBITS);
    return CCSM_EINVAL;
  }
  const MtJumpTables& T = mt_jump_tables();
  if (!T.ok) {
    set_error("ccsm_debug_mt_jump_check: the characteristic polynomial did not come out with degree 19937");
    return CCSM_ESTATE;
  }
  const long long J = MT_J << k;
  mtjump::MT g(seed);
  std::vector<uint32_t> x((size_t)J + MT_DEG + 2 * MT_N);
  for (auto& v : x) v = g.next();
  const int b = 17;
  int bad = 0;
  const uint16_t* lst = T.lists.data() + (size_t)k * MT_JLIST;
  for (int j = 0; j < MT_N; ++j) {
    uint32_t acc = 0;
    for (int i = 0; i < T.counts[k]; ++i) acc ^= x[(size_t)b + lst[i] + j];
    const uint32_t want = x[(size_t)b + J + j];
    if (j == 0 ? ((acc ^ want) & 0x80000000u) != 0 : acc != want) ++bad;
  }
  return bad;
}
