// Device feature extraction for call_mods: reads (raw bytes + descriptors) -> site list -> the reference's
// 16-tensor feature layout -> forward -> per-site MM/ML values.  HBM-bound byte/integer kernels.
//
// Reference behaviour restated here (all under /root/reference/ccsmeth/):
//   extract_features.py:261-406   extract_features_from_double_strand_read (site selection and windows)
//   extract_features.py:181-199   _normalize_signals (np.mean / population np.std / np.around(6), float64; "mad" = np.median
//                                 and statsmodels 0.14 robust.scale.mad, restated from its published formula)
//   utils/process_utils.py:426-449 CodecV1 code -> frames
//   utils/process_utils.py:122-137 get_refloc_of_methysite_in_motif
//   call_modifications.py:73-123  _batch_feature_list2s (base codes, npass broadcast over the window)
//   call_modifications.py:222-224 prob_1_norm;  _bam2modbam.py:187-208 MM deltas / ML bytes
//
// Arithmetic notes.  The kinetics are small integers (<= 952), so sum and sum of squares of a read are exact in
// int64: mean = S/n and var = (n*Q - S*S)/n^2 are each one IEEE division away from the exact rational, where
// numpy's pairwise float64 sums may differ from it in the last ulps.  Every later step ((x-mean)/std, *1e6, rint,
// /1e6, cast to float32) is the same sequence of IEEE double operations numpy performs (no FMA contraction).
#include <math.h>
#include <string.h>

#include <algorithm>

#include "ccsm_internal.h"

namespace ccsm {

struct ExState {
  DevBuf blob, reads, stats, site_cnt, site_off, site_read, site_loc, site_cord;
  DevBuf feat, h0, out, tags;   // per-chunk staging: 16-tensor features, h0, logits/probs, prob1/mm/ml
  ccsm_extract_opts opts{};
  int32_t n_reads = 0;
  int64_t n_sites = -1;
  int64_t blob_bytes = 0;
  cudaStream_t st = nullptr, st_copy = nullptr;
  cudaEvent_t ev[2] = {nullptr, nullptr}, done[2] = {nullptr, nullptr};
};

struct ExParams {
  const uint8_t* blob;
  const ccsm_read* reads;
  int n_reads;
  int seq_len, nb, mod_loc, rev_offset, norm, decode, n_motifs, motif_len;
  uint32_t motif_code[8];  // motif k packed 4 bits per base code (A0 C1 G2 T3)
};

__device__ __forceinline__ int nib_to_code(int nib) {
  return nib == 1 ? 0 : nib == 2 ? 1 : nib == 4 ? 2 : nib == 8 ? 3 : 4;
}
__device__ __forceinline__ int ascii_to_code(int c) {
  return c == 'A' ? 0 : c == 'C' ? 1 : c == 'G' ? 2 : c == 'T' ? 3 : 4;
}
__host__ __device__ __forceinline__ int code_to_frames(int c) {
  // CodecV1 (process_utils.py:426-449): 64 codes each at strides 1, 2, 4, 8
  return c < 64 ? c : c < 128 ? 64 + ((c - 64) << 1) : c < 192 ? 192 + ((c - 128) << 2) : 448 + ((c - 192) << 3);
}

// base code (0..4) of base i of the FORWARD read
__device__ __forceinline__ int fwd_code(const uint8_t* blob, const ccsm_read& r, int i) {
  const int j = (r.flags & CCSM_READ_REVERSE) ? r.len - 1 - i : i;
  int c;
  if (r.flags & CCSM_READ_SEQ_4BIT) {
    const int b = blob[r.seq_off + (j >> 1)];
    c = nib_to_code((j & 1) ? (b & 15) : (b >> 4));
  } else {
    c = ascii_to_code(blob[r.seq_off + j]);
  }
  if ((r.flags & CCSM_READ_REVERSE) && c < 4) c = 3 - c;
  return c;
}

__device__ __forceinline__ long long warp_sum_ll(long long v) {
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---- sequence tiles in shared memory -----------------------------------------------------------------------
// Both per-read kernels walk the read in tiles of SEQ_TILE forward positions.  A tile (plus a halo long enough for
// the longest motif and mod_loc) is first expanded into one base code per byte in shared memory -- the only place
// that knows about 4-bit packing and reverse-strand storage -- and sites / C's are then flagged from there.
constexpr int SEQ_TILE = 4096;
constexpr int SEQ_HALO = 16;

__device__ __forceinline__ void stage_codes(const ExParams& P, const ccsm_read& r, int t0, uint8_t* s_codes) {
  const int n = min(SEQ_TILE + SEQ_HALO, r.len - t0);
  if ((r.flags & CCSM_READ_SEQ_4BIT) && !(r.flags & CCSM_READ_REVERSE)) {
    // forward 4-bit: t0 is even, so stored byte (t0 >> 1) + i holds positions t0 + 2i, t0 + 2i + 1
    const uint8_t* src = P.blob + r.seq_off + (t0 >> 1);
    const int nb = (n + 1) >> 1;
    for (int i = threadIdx.x; i < nb; i += blockDim.x) {
      const int b = src[i];
      s_codes[2 * i] = (uint8_t)nib_to_code(b >> 4);
      s_codes[2 * i + 1] = (uint8_t)nib_to_code(b & 15);  // position t0 + n may get a pad nibble: never read
    }
  } else {
    for (int i = threadIdx.x; i < n; i += blockDim.x) s_codes[i] = (uint8_t)fwd_code(P.blob, r, t0 + i);
  }
}

// is tile-relative position q (forward position t0 + q) the start of a motif whose modified base is a callable site?
// (extract_features.py:336-343; needs codes up to q + motif_len - 1 < SEQ_TILE + SEQ_HALO)
__device__ __forceinline__ bool site_in_tile(const ExParams& P, const ccsm_read& r, int t0, int q, const uint8_t* s_codes) {
  const int p = t0 + q;
  if (p + P.motif_len > r.len) return false;
  uint32_t w = 0;
  for (int k = 0; k < P.motif_len; ++k) {
    const uint32_t c = s_codes[q + k];
    if (c > 3) return false;
    w |= c << (4 * k);
  }
  bool hit = false;
  for (int k = 0; k < P.n_motifs; ++k) hit |= (w == P.motif_code[k]);
  if (!hit) return false;
  const int loc = p + P.mod_loc;
  const int rl = r.len - 1 - (loc + P.rev_offset);
  return loc >= P.nb && loc < r.len - P.nb && rl >= P.nb && rl < r.len - P.nb && loc >= r.win_lo && loc < r.win_hi;
}

// ---- kernel 1: one CTA per read -- normalisation statistics of the four kinetics arrays + site count
struct SigStat { double shift, scale; };

__device__ __forceinline__ void acc_code(int c, int decode, long long& S, long long& Q, int& mn, int& mx) {
  const int v = decode ? code_to_frames(c) : c;
  S += v;
  Q += v * v;
  mn = min(mn, v);
  mx = max(mx, v);
}

__global__ void __launch_bounds__(256) read_scan_kernel(ExParams P, SigStat* __restrict__ stats,
                                                        int* __restrict__ site_cnt) {
  const int r_idx = blockIdx.x;
  const ccsm_read r = P.reads[r_idx];
  __shared__ long long s_sum[8], s_sq[8];
  __shared__ int s_min[8], s_max[8], s_cnt[8];
  __shared__ int s_hist[256];  // --norm mad: occurrences of every kinetics code in the read (the codes are bytes)
  __shared__ __align__(16) uint8_t s_codes[SEQ_TILE + SEQ_HALO + 16];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int sig = 0; sig < 4; ++sig) {
    // order: ipd fwd, ipd rev, pw fwd, pw rev
    const int64_t off = sig == 0 ? r.fi_off : sig == 1 ? r.ri_off : sig == 2 ? r.fp_off : r.rp_off;
    long long S = 0, Q = 0;
    int mn = 1 << 30, mx = -1;
    if (P.norm == CCSM_NORM_MAD) {
      for (int i = threadIdx.x; i < 256; i += blockDim.x) s_hist[i] = 0;
      __syncthreads();
      const uint8_t* a = P.blob + off;
      for (int i = threadIdx.x; i < r.len; i += blockDim.x) atomicAdd(&s_hist[a[i]], 1);
    } else if (P.norm != CCSM_NORM_NONE) {
      // 16-byte loads over the aligned span covering [off, off + len); only the two end chunks are masked
      // (the blob allocation is 16-byte aligned and padded by 16 bytes)
      const uint8_t* a = P.blob + off;
      const int mis = (int)(reinterpret_cast<uintptr_t>(a) & 15);
      const uint4* A = reinterpret_cast<const uint4*>(a - mis);
      const int nchunk = (r.len + mis + 15) >> 4;
      for (int c = threadIdx.x; c < nchunk; c += blockDim.x) {
        const uint4 v = __ldg(A + c);
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
        if (c > 0 && c < nchunk - 1) {
#pragma unroll
          for (int k = 0; k < 16; ++k) acc_code((w[k >> 2] >> (8 * (k & 3))) & 255, P.decode, S, Q, mn, mx);
        } else {
#pragma unroll
          for (int k = 0; k < 16; ++k) {
            const int i = c * 16 + k - mis;
            if (i >= 0 && i < r.len) acc_code((w[k >> 2] >> (8 * (k & 3))) & 255, P.decode, S, Q, mn, mx);
          }
        }
      }
    }
    S = warp_sum_ll(S);
    Q = warp_sum_ll(Q);
    for (int o = 16; o; o >>= 1) {
      mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, o));
      mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    if (lane == 0) { s_sum[warp] = S; s_sq[warp] = Q; s_min[warp] = mn; s_max[warp] = mx; }
    __syncthreads();
    if (threadIdx.x == 0) {
      long long St = 0, Qt = 0;
      int mnt = 1 << 30, mxt = -1;
      for (int w = 0; w < 8; ++w) { St += s_sum[w]; Qt += s_sq[w]; mnt = min(mnt, s_min[w]); mxt = max(mxt, s_max[w]); }
      SigStat st;
      const double n = (double)r.len;
      if (P.norm == CCSM_NORM_ZSCORE) {
        st.shift = __ddiv_rn((double)St, n);
        const long long num = (long long)r.len * Qt - St * St;  // exact: n^2 * variance
        st.scale = __dsqrt_rn(__ddiv_rn((double)num, __dmul_rn(n, n)));
      } else if (P.norm == CCSM_NORM_MINMEAN) {
        st.shift = (double)mnt;
        st.scale = __ddiv_rn((double)St, n);
      } else if (P.norm == CCSM_NORM_MINMAX) {
        st.shift = (double)mnt;
        st.scale = (double)(mxt - mnt);
      } else if (P.norm == CCSM_NORM_MAD && r.len > 0) {
        // np.median and statsmodels.robust.scale.mad (0.14: median(|a - median(a)| / c), c = Gaussian.ppf(3/4), float64)
        // from the 256-bin histogram: the code -> value map is increasing, so code order is value order
        const int k1 = (r.len - 1) >> 1, k2 = r.len >> 1;  // the middle element(s), 0-based
        int v1 = -1, v2 = -1, c2 = 0, cum = 0;
        for (int c = 0; c < 256 && v2 < 0; ++c) {
          const int cnt = s_hist[c];
          if (!cnt) continue;
          const int v = P.decode ? code_to_frames(c) : c;
          if (v1 < 0 && cum + cnt > k1) v1 = v;
          if (cum + cnt > k2) { v2 = v; c2 = c; }
          cum += cnt;
        }
        const double med = v1 == v2 ? (double)v1 : __ddiv_rn((double)(v1 + v2), 2.0);
        // |value - med| is V-shaped over the codes: merge the two sorted arms (downwards from the median, upwards above it)
        int li = c2, ri = c2 + 1;
        while (li >= 0 && (double)(P.decode ? code_to_frames(li) : li) > med) --li;  // even n: c2 is the upper middle
        ri = li + 1;
        double d1 = -1.0, d2 = -1.0;
        cum = 0;
        while ((li >= 0 || ri < 256) && d2 < 0.0) {
          const double dl = li >= 0 ? med - (double)(P.decode ? code_to_frames(li) : li) : 1e300;
          const double dr = ri < 256 ? (double)(P.decode ? code_to_frames(ri) : ri) - med : 1e300;
          double d;
          int cnt;
          if (dl <= dr) { d = dl; cnt = s_hist[li]; --li; } else { d = dr; cnt = s_hist[ri]; ++ri; }
          if (!cnt) continue;
          if (d1 < 0.0 && cum + cnt > k1) d1 = d;
          if (cum + cnt > k2) d2 = d;
          cum += cnt;
        }
        const double c = 0.6744897501960817;
        const double x1 = __ddiv_rn(d1, c), x2 = __ddiv_rn(d2, c);
        st.shift = med;
        st.scale = d1 == d2 ? x1 : __ddiv_rn(__dadd_rn(x1, x2), 2.0);
      } else {
        st.shift = 0.0;
        st.scale = 1.0;
      }
      if (r.len == 0) { st.shift = 0.0; st.scale = 0.0; }
      stats[(size_t)r_idx * 4 + sig] = st;
    }
    __syncthreads();
  }
  int cnt = 0;
  for (int t0 = 0; t0 < r.len; t0 += SEQ_TILE) {
    stage_codes(P, r, t0, s_codes);
    __syncthreads();
    const int nq = min(SEQ_TILE, r.len - t0);
    for (int q = threadIdx.x; q < nq; q += blockDim.x) cnt += site_in_tile(P, r, t0, q, s_codes) ? 1 : 0;
    __syncthreads();
  }
  for (int o = 16; o; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  if (lane == 0) s_cnt[warp] = cnt;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int w = 0; w < 8; ++w) t += s_cnt[w];
    site_cnt[r_idx] = t;
  }
}

// ---- kernel 2: exclusive scan of the per-read site counts (one CTA; n_reads is small)
__global__ void __launch_bounds__(1024) site_offsets_kernel(const int* __restrict__ cnt, long long* __restrict__ off,
                                                            int n) {
  __shared__ long long s_warp[32];
  __shared__ long long s_carry;
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int base = 0; base < n; base += 1024) {
    const int i = base + threadIdx.x;
    const long long v = i < n ? cnt[i] : 0;
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
      s_warp[lane] = w;  // inclusive over warps
    }
    __syncthreads();
    const long long before = s_carry + (warp ? s_warp[warp - 1] : 0) + x - v;
    if (i < n) off[i] = before;
    __syncthreads();
    if (threadIdx.x == 1023) s_carry = before + v;
    __syncthreads();
  }
  if (threadIdx.x == 0) off[n] = s_carry;
}

// ---- kernel 3: one CTA per read -- ordered emission of (read, loc, order of the C among the read's C's)
__global__ void __launch_bounds__(256) site_emit_kernel(ExParams P, const long long* __restrict__ site_off,
                                                        int* __restrict__ site_read, int* __restrict__ site_loc,
                                                        int* __restrict__ site_cord) {
  const int r_idx = blockIdx.x;
  const ccsm_read r = P.reads[r_idx];
  const long long o0 = site_off[r_idx];
  if (site_off[r_idx + 1] == o0) return;
  __shared__ int s_ws[8], s_wc[8];
  __shared__ int s_cs, s_cc;  // running carries: sites, C's
  __shared__ __align__(16) uint8_t s_codes[SEQ_TILE + SEQ_HALO + 16];
  if (threadIdx.x == 0) {
    s_cs = 0;
    int c0 = 0;  // C's in front of the first position the loop below looks at
    for (int q = 0; q < P.mod_loc && q < r.len; ++q) c0 += fwd_code(P.blob, r, q) == 1;
    s_cc = c0;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int t0 = 0; t0 < r.len; t0 += SEQ_TILE) {
    stage_codes(P, r, t0, s_codes);
    __syncthreads();
    const int nq = min(SEQ_TILE, r.len - t0);
    for (int base = 0; base < nq; base += 256) {
      const int q = base + threadIdx.x;
      const bool in = q < nq;
      const bool is_site = in && site_in_tile(P, r, t0, q, s_codes);
      // MM deltas count C's of the forward read (_bam2modbam.py:187-203); the called base sits at p + mod_loc, so
      // flag the C at position p + mod_loc in p's own thread (mod_loc < motif_len <= 8 stays inside the halo)
      const int pc_pos = t0 + q + P.mod_loc;
      const bool is_c = in && pc_pos < r.len && s_codes[q + P.mod_loc] == 1;
      const unsigned ms = __ballot_sync(0xffffffffu, is_site), mc = __ballot_sync(0xffffffffu, is_c);
      const unsigned below = (1u << lane) - 1u;
      if (lane == 0) { s_ws[warp] = __popc(ms); s_wc[warp] = __popc(mc); }
      __syncthreads();
      int ps = s_cs, pc = s_cc;
      for (int w = 0; w < warp; ++w) { ps += s_ws[w]; pc += s_wc[w]; }
      ps += __popc(ms & below);
      pc += __popc(mc & below);
      if (is_site) {
        site_read[o0 + ps] = r_idx;
        site_loc[o0 + ps] = pc_pos;
        site_cord[o0 + ps] = pc;  // C's strictly before the called base
      }
      __syncthreads();
      if (threadIdx.x == 0) {
        int ts = 0, tc = 0;
        for (int w = 0; w < 8; ++w) { ts += s_ws[w]; tc += s_wc[w]; }
        s_cs += ts;
        s_cc += tc;
      }
      __syncthreads();
    }
  }
}

// ---- kernel 4: the 16-tensor layout of sites [s0, s0+cn)
struct FeatOut {
  float *kmer, *kpass, *ipd, *pw, *sns;
};

__device__ __forceinline__ float norm_value(int v, const SigStat& st, int norm) {
  if (norm == CCSM_NORM_NONE) return (float)v;
  if (st.scale == 0.0) return 0.f;
  const double x = __ddiv_rn(__dsub_rn((double)v, st.shift), st.scale);
  return (float)__ddiv_rn(rint(__dmul_rn(x, 1e6)), 1e6);  // np.around(x, 6): multiply, rint, divide
}

// One thread per (site, strand) computes the 21-position window into shared memory; the CTA then writes each output
// array as one contiguous, coalesced run (rows of GATHER_SITES consecutive sites).
constexpr int GATHER_SITES = 64;

__global__ void __launch_bounds__(2 * GATHER_SITES) window_gather_kernel(ExParams P, const SigStat* __restrict__ stats,
                                                                         const int* __restrict__ site_read,
                                                                         const int* __restrict__ site_loc, long long s0,
                                                                         long long cn, FeatOut f, FeatOut rv) {
  extern __shared__ float s_out[];  // [strand][array kmer, kpass, ipd, pw][site][L]
  __shared__ float s_sn[GATHER_SITES][4];
  const int L = P.seq_len;
  const int strand = threadIdx.x / GATHER_SITES, ls = threadIdx.x % GATHER_SITES;
  const int per_arr = GATHER_SITES * L;
  for (long long g0 = (long long)blockIdx.x * GATHER_SITES; g0 < cn; g0 += (long long)gridDim.x * GATHER_SITES) {
    const long long i = g0 + ls;
    if (i < cn) {
      const int r_idx = site_read[s0 + i];
      const int loc = site_loc[s0 + i];
      const ccsm_read* rp = P.reads + r_idx;
      const int len = rp->len, flags = rp->flags;
      const SigStat st_ipd = stats[(size_t)r_idx * 4 + (strand ? 1 : 0)];
      const SigStat st_pw = stats[(size_t)r_idx * 4 + (strand ? 3 : 2)];
      const float npass = (float)(strand ? rp->rn : rp->fn);
      // forward strand: window [loc - nb, loc + nb].  reverse strand: window on the reverse complement around
      // len-1-(loc+rev_offset); its kinetics are indexed in that same coordinate, unflipped
      // (extract_features.py:316,319,355-360)
      const int w0 = strand ? len - 1 - (loc + P.rev_offset) - P.nb : loc - P.nb;
      const uint8_t* ipd_a = P.blob + (strand ? rp->ri_off : rp->fi_off) + w0;
      const uint8_t* pw_a = P.blob + (strand ? rp->rp_off : rp->fp_off) + w0;
      const int64_t seq_off = rp->seq_off;
      float* o_kmer = s_out + (strand * 4 + 0) * per_arr + ls * L;
      float* o_kpass = s_out + (strand * 4 + 1) * per_arr + ls * L;
      float* o_ipd = s_out + (strand * 4 + 2) * per_arr + ls * L;
      float* o_pw = s_out + (strand * 4 + 3) * per_arr + ls * L;
      for (int t = 0; t < L; ++t) {
        // forward-read position whose base this window slot shows
        const int fpos = strand ? len - 1 - (w0 + t) : w0 + t;
        const int j = (flags & CCSM_READ_REVERSE) ? len - 1 - fpos : fpos;
        int c;
        if (flags & CCSM_READ_SEQ_4BIT) {
          const int b = P.blob[seq_off + (j >> 1)];
          c = nib_to_code((j & 1) ? (b & 15) : (b >> 4));
        } else {
          c = ascii_to_code(P.blob[seq_off + j]);
        }
        // complement once per strand flip: stored-reverse and reverse-strand window cancel each other
        const bool comp = ((flags & CCSM_READ_REVERSE) != 0) != (strand != 0);
        if (comp && c < 4) c = 3 - c;
        const int ic = ipd_a[t], pc = pw_a[t];
        o_kmer[t] = (float)c;
        o_kpass[t] = npass;
        o_ipd[t] = norm_value(P.decode ? code_to_frames(ic) : ic, st_ipd, P.norm);
        o_pw[t] = norm_value(P.decode ? code_to_frames(pc) : pc, st_pw, P.norm);
      }
      if (strand == 0 && f.sns) {
        // np.around(np.array(tag_sn, dtype=float), 6) (extract_features.py:328)
        for (int k = 0; k < 4; ++k) s_sn[ls][k] = (float)__ddiv_rn(rint(__dmul_rn((double)rp->sn[k], 1e6)), 1e6);
      }
    }
    __syncthreads();
    const int rows = (int)min((long long)GATHER_SITES, cn - g0);
    const int nval = rows * L;
    const long long obase = g0 * L;
    for (int k = threadIdx.x; k < nval; k += blockDim.x) {
      f.kmer[obase + k] = s_out[0 * per_arr + k];
      if (f.kpass) f.kpass[obase + k] = s_out[1 * per_arr + k];
      f.ipd[obase + k] = s_out[2 * per_arr + k];
      f.pw[obase + k] = s_out[3 * per_arr + k];
      rv.kmer[obase + k] = s_out[4 * per_arr + k];
      if (rv.kpass) rv.kpass[obase + k] = s_out[5 * per_arr + k];
      rv.ipd[obase + k] = s_out[6 * per_arr + k];
      rv.pw[obase + k] = s_out[7 * per_arr + k];
    }
    if (f.sns)
      for (int k = threadIdx.x; k < rows * 4; k += blockDim.x) {
        const float v = s_sn[k >> 2][k & 3];
        f.sns[g0 * 4 + k] = v;
        if (rv.sns) rv.sns[g0 * 4 + k] = v;
      }
    __syncthreads();
  }
}

// ---- kernel 5: per-site outputs for the modbam writer
__global__ void __launch_bounds__(256) site_tags_kernel(const float* __restrict__ probs, int classes,
                                                        const int* __restrict__ site_read,
                                                        const int* __restrict__ site_cord, long long s0, long long cn,
                                                        float* __restrict__ prob1, int* __restrict__ mm,
                                                        uint8_t* __restrict__ ml) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= cn) return;
  const float p0 = probs[i * classes], p1 = probs[i * classes + 1];
  // round(prob_1 / (prob_0 + prob_1), 6) on float32 scalars (call_modifications.py:222-223)
  const float p = __fdiv_rn(rintf(__fmul_rn(__fdiv_rn(p1, __fadd_rn(p0, p1)), 1e6f)), 1e6f);
  prob1[i] = p;
  ml[i] = p < 1.f ? (uint8_t)floorf(__fmul_rn(p, 256.f)) : (uint8_t)255;  // _bam2modbam.py:206-208
  const long long g = s0 + i;
  const bool first = g == 0 || site_read[g - 1] != site_read[g];
  mm[i] = first ? site_cord[g] : site_cord[g] - site_cord[g - 1] - 1;    // _bam2modbam.py:200-203
}

// ------------------------------------------------------------------------------------------------------------
static int ensure_streams(ExState* ex) {
  if (!ex->st) CCSM_CUDA(cudaStreamCreateWithFlags(&ex->st, cudaStreamNonBlocking));
  if (!ex->st_copy) CCSM_CUDA(cudaStreamCreateWithFlags(&ex->st_copy, cudaStreamNonBlocking));
  for (int i = 0; i < 2; ++i) {
    if (!ex->ev[i]) CCSM_CUDA(cudaEventCreateWithFlags(&ex->ev[i], cudaEventDisableTiming));
    if (!ex->done[i]) CCSM_CUDA(cudaEventCreateWithFlags(&ex->done[i], cudaEventDisableTiming));
  }
  return CCSM_OK;
}

static int make_params(const ccsm_model* m, const ExState* ex, ExParams& P) {
  const ccsm_extract_opts& o = ex->opts;
  P.blob = ex->blob.as<uint8_t>();
  P.reads = ex->reads.as<ccsm_read>();
  P.n_reads = ex->n_reads;
  P.seq_len = m->cfg.seq_len;
  P.nb = (m->cfg.seq_len - 1) / 2;
  P.mod_loc = o.mod_loc;
  P.rev_offset = (o.motif_len - 1 - o.mod_loc) - o.mod_loc;  // extract_features.py:333
  P.norm = o.norm;
  P.decode = o.decode;
  P.n_motifs = o.n_motifs;
  P.motif_len = o.motif_len;
  for (int k = 0; k < 8; ++k) P.motif_code[k] = 0xffffffffu;
  for (int k = 0; k < o.n_motifs; ++k) {
    uint32_t w = 0;
    for (int j = 0; j < o.motif_len; ++j) {
      const char c = o.motifs[k * o.motif_len + j];
      const int code = c == 'A' ? 0 : c == 'C' ? 1 : c == 'G' ? 2 : c == 'T' ? 3 : -1;
      if (code < 0) {
        set_error("ccsm_reads_extract_host: motif character '%c' is not one of ACGT (expand IUPAC codes first)", c);
        return CCSM_EINVAL;
      }
      w |= (uint32_t)code << (4 * j);
    }
    P.motif_code[k] = w;
  }
  return CCSM_OK;
}

void ex_release(ccsm_model* m) {
  ExState* ex = m->ex;
  if (!ex) return;
  for (DevBuf* b : {&ex->blob, &ex->reads, &ex->stats, &ex->site_cnt, &ex->site_off, &ex->site_read, &ex->site_loc,
                    &ex->site_cord, &ex->feat, &ex->h0, &ex->out, &ex->tags})
    b->release();
  if (ex->st) cudaStreamDestroy(ex->st);
  if (ex->st_copy) cudaStreamDestroy(ex->st_copy);
  for (int i = 0; i < 2; ++i) {
    if (ex->ev[i]) cudaEventDestroy(ex->ev[i]);
    if (ex->done[i]) cudaEventDestroy(ex->done[i]);
  }
  delete ex;
  m->ex = nullptr;
}

static int check_model(ccsm_model* m, const char* fn) {
  if (!m || m->cfg.kind != CCSM_KIND_ATT2S) {
    set_error("%s: not an att2s model", fn);
    return CCSM_EINVAL;
  }
  if (m->cfg.feat_flags & (CCSM_FEAT_STDS | CCSM_FEAT_MAP)) {
    set_error("%s: --is_stds / --is_map features are not produced by the device extractor", fn);
    return CCSM_EUNSUPPORTED;
  }
  return CCSM_OK;
}

static int launch_features(ccsm_model* m, int64_t s0, int64_t cn, const ccsm_strand* fo, const ccsm_strand* ro,
                           cudaStream_t st) {
  ExState* ex = m->ex;
  ExParams P;
  CCSM_TRY(make_params(m, ex, P));
  FeatOut f{const_cast<float*>(fo->kmer), const_cast<float*>(fo->kpass), const_cast<float*>(fo->ipd_means),
            const_cast<float*>(fo->pw_means), const_cast<float*>(fo->sns)};
  FeatOut r{const_cast<float*>(ro->kmer), const_cast<float*>(ro->kpass), const_cast<float*>(ro->ipd_means),
            const_cast<float*>(ro->pw_means), const_cast<float*>(ro->sns)};
  if (!f.kmer || !f.ipd || !f.pw || !r.kmer || !r.ipd || !r.pw) {
    set_error("ccsm_reads_features: kmer / ipd_means / pw_means outputs are required");
    return CCSM_EINVAL;
  }
  const size_t smem = (size_t)8 * GATHER_SITES * P.seq_len * sizeof(float);
  static bool attr = false;
  if (!attr) {
    CCSM_CUDA(cudaFuncSetAttribute(window_gather_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * GATHER_SITES * 32 * 4));
    attr = true;
  }
  const int grid = (int)std::min<long long>((cn + GATHER_SITES - 1) / GATHER_SITES, 148LL * 8);
  const int pid = m->prof.begin(PROF_EX_GATHER, (double)cn, st);
  window_gather_kernel<<<grid, 2 * GATHER_SITES, smem, st>>>(P, ex->stats.as<SigStat>(), ex->site_read.as<int>(),
                                                             ex->site_loc.as<int>(), s0, cn, f, r);
  m->prof.end(pid, st);
  count_launch();
  CCSM_CUDA(cudaGetLastError());
  return CCSM_OK;
}

}  // namespace ccsm

using namespace ccsm;

extern "C" {

int ccsm_reads_extract_host(ccsm_model* m, const ccsm_extract_opts* o, const uint8_t* blob, int64_t blob_bytes,
                            const ccsm_read* reads, int32_t n_reads, int64_t* n_sites) {
  CCSM_TRY(check_model(m, "ccsm_reads_extract_host"));
  if (!o || !n_sites || n_reads < 0 || blob_bytes < 0 || (n_reads > 0 && (!blob || !reads))) {
    set_error("ccsm_reads_extract_host: bad argument");
    return CCSM_EINVAL;
  }
  if (o->n_motifs < 1 || o->n_motifs > 8 || o->motif_len < 1 || o->motif_len > 8 || o->mod_loc < 0 ||
      o->mod_loc >= o->motif_len || o->norm < CCSM_NORM_ZSCORE || o->norm > CCSM_NORM_MAD) {
    set_error("ccsm_reads_extract_host: unsupported options (motifs=%d x %d, mod_loc=%d, norm=%d)", o->n_motifs,
              o->motif_len, o->mod_loc, o->norm);
    return CCSM_EINVAL;
  }
  for (int i = 0; i < n_reads; ++i) {
    const ccsm_read& r = reads[i];
    const int64_t seq_bytes = (r.flags & CCSM_READ_SEQ_4BIT) ? (r.len + 1) / 2 : r.len;
    if (r.len < 0 || r.seq_off < 0 || r.seq_off + seq_bytes > blob_bytes || r.fi_off < 0 || r.ri_off < 0 ||
        r.fp_off < 0 || r.rp_off < 0 || r.fi_off + r.len > blob_bytes || r.ri_off + r.len > blob_bytes ||
        r.fp_off + r.len > blob_bytes || r.rp_off + r.len > blob_bytes) {
      set_error("ccsm_reads_extract_host: read %d points outside the blob", i);
      return CCSM_EINVAL;
    }
  }
  CCSM_CUDA(cudaSetDevice(m->cfg.device));
  if (!m->ex) m->ex = new (std::nothrow) ExState();
  ExState* ex = m->ex;
  if (!ex) return CCSM_ENOMEM;
  CCSM_TRY(ensure_streams(ex));
  ex->opts = *o;
  ex->n_reads = n_reads;
  ex->n_sites = 0;
  ex->blob_bytes = blob_bytes;
  *n_sites = 0;
  if (n_reads == 0) return CCSM_OK;
  cudaStream_t st = ex->st;
  CCSM_TRY(ex->blob.reserve((size_t)blob_bytes + 16));
  CCSM_TRY(ex->reads.reserve((size_t)n_reads * sizeof(ccsm_read)));
  CCSM_TRY(ex->stats.reserve((size_t)n_reads * 4 * sizeof(SigStat)));
  CCSM_TRY(ex->site_cnt.reserve((size_t)n_reads * sizeof(int)));
  CCSM_TRY(ex->site_off.reserve((size_t)(n_reads + 1) * sizeof(long long)));
  CCSM_CUDA(cudaMemcpyAsync(ex->blob.p, blob, (size_t)blob_bytes, cudaMemcpyHostToDevice, st));
  CCSM_CUDA(cudaMemcpyAsync(ex->reads.p, reads, (size_t)n_reads * sizeof(ccsm_read), cudaMemcpyHostToDevice, st));
  ExParams P;
  CCSM_TRY(make_params(m, ex, P));
  double n_bases = 0;
  for (int i = 0; i < n_reads; ++i) n_bases += reads[i].len;
  int pid = m->prof.begin(PROF_EX_SCAN, n_bases, st);
  read_scan_kernel<<<n_reads, 256, 0, st>>>(P, ex->stats.as<SigStat>(), ex->site_cnt.as<int>());
  site_offsets_kernel<<<1, 1024, 0, st>>>(ex->site_cnt.as<int>(), ex->site_off.as<long long>(), n_reads);
  m->prof.end(pid, st);
  count_launch(2);
  CCSM_CUDA(cudaGetLastError());
  long long total = 0;
  CCSM_CUDA(cudaMemcpyAsync(&total, ex->site_off.as<long long>() + n_reads, sizeof(long long), cudaMemcpyDeviceToHost,
                            st));
  CCSM_CUDA(cudaStreamSynchronize(st));
  if (total > 0) {
    CCSM_TRY(ex->site_read.reserve((size_t)total * sizeof(int)));
    CCSM_TRY(ex->site_loc.reserve((size_t)total * sizeof(int)));
    CCSM_TRY(ex->site_cord.reserve((size_t)total * sizeof(int)));
    pid = m->prof.begin(PROF_EX_SCAN, 0, st);
    site_emit_kernel<<<n_reads, 256, 0, st>>>(P, ex->site_off.as<long long>(), ex->site_read.as<int>(),
                                              ex->site_loc.as<int>(), ex->site_cord.as<int>());
    m->prof.end(pid, st);
    count_launch();
    CCSM_CUDA(cudaGetLastError());
  }
  ex->n_sites = total;
  *n_sites = total;
  return CCSM_OK;
}

int ccsm_reads_sites(ccsm_model* m, int32_t* site_read, int32_t* site_loc) {
  CCSM_TRY(check_model(m, "ccsm_reads_sites"));
  if (!m->ex || m->ex->n_sites < 0) {
    set_error("ccsm_reads_sites: no resident read batch (call ccsm_reads_extract_host first)");
    return CCSM_ESTATE;
  }
  ExState* ex = m->ex;
  if (ex->n_sites == 0) return CCSM_OK;
  CCSM_CUDA(cudaSetDevice(m->cfg.device));
  const size_t bytes = (size_t)ex->n_sites * sizeof(int);
  if (site_read) CCSM_CUDA(cudaMemcpyAsync(site_read, ex->site_read.p, bytes, cudaMemcpyDeviceToHost, ex->st));
  if (site_loc) CCSM_CUDA(cudaMemcpyAsync(site_loc, ex->site_loc.p, bytes, cudaMemcpyDeviceToHost, ex->st));
  CCSM_CUDA(cudaStreamSynchronize(ex->st));
  return CCSM_OK;
}

int ccsm_reads_features(ccsm_model* m, int64_t s0, int64_t cn, const ccsm_strand* fwd_out, const ccsm_strand* rev_out,
                        void* stream) {
  CCSM_TRY(check_model(m, "ccsm_reads_features"));
  if (!m->ex || m->ex->n_sites < 0) {
    set_error("ccsm_reads_features: no resident read batch (call ccsm_reads_extract_host first)");
    return CCSM_ESTATE;
  }
  if (!fwd_out || !rev_out || s0 < 0 || cn < 0 || s0 + cn > m->ex->n_sites) {
    set_error("ccsm_reads_features: bad site range [%lld, %lld) of %lld", (long long)s0, (long long)(s0 + cn),
              (long long)m->ex->n_sites);
    return CCSM_EINVAL;
  }
  if (cn == 0) return CCSM_OK;
  CCSM_CUDA(cudaSetDevice(m->cfg.device));
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  // the site list was produced on the extractor's stream
  CCSM_CUDA(cudaEventRecord(m->ex->ev[0], m->ex->st));
  CCSM_CUDA(cudaStreamWaitEvent(st, m->ex->ev[0], 0));
  return launch_features(m, s0, cn, fwd_out, rev_out, st);
}

int ccsm_reads_forward_host(ccsm_model* m, const float* h0_fwd, const float* h0_rev, float* logits, float* probs,
                            float* prob1, int32_t* mm_delta, uint8_t* ml) {
  CCSM_TRY(check_model(m, "ccsm_reads_forward_host"));
  if (!m->finalized) {
    set_error("ccsm_reads_forward_host: model not finalized");
    return CCSM_ESTATE;
  }
  if (!m->ex || m->ex->n_sites < 0) {
    set_error("ccsm_reads_forward_host: no resident read batch (call ccsm_reads_extract_host first)");
    return CCSM_ESTATE;
  }
  if (m->cfg.num_classes != 2 && (prob1 || mm_delta || ml)) {
    set_error("ccsm_reads_forward_host: prob1/mm/ml need a 2-class model");
    return CCSM_EINVAL;
  }
  ExState* ex = m->ex;
  const int64_t n = ex->n_sites;
  if (n == 0) return CCSM_OK;
  CCSM_CUDA(cudaSetDevice(m->cfg.device));
  const int L = m->cfg.seq_len, H = m->cfg.hidden, NL = m->cfg.num_layers, C = m->cfg.num_classes;
  const bool has_np = m->cfg.feat_flags & CCSM_FEAT_NPASS, has_sn = m->cfg.feat_flags & CCSM_FEAT_SN;
  const int64_t kChunk = 75776;  // one tensor-core library chunk
  // torch-stream h0 mode: chunks are cut at the reference's model-call boundaries (whole randn calls per chunk)
  const bool torch_h0 = m->h0_mode == CCSM_H0_TORCH_STREAM && !(h0_fwd && h0_rev) && m->gates == 3 && !m->is_trans;
  std::vector<int64_t> segs;
  std::vector<SegChunk> chunks;
  if (torch_h0) {
    CCSM_TRY(mt_take_segments(m, n, segs));
    seg_chunks(segs, kChunk, chunks);
  } else {
    for (int64_t s0 = 0; s0 < n; s0 += kChunk) chunks.push_back(SegChunk{0, 0, s0, (n - s0) < kChunk ? (n - s0) : kChunk});
  }
  int64_t chunk = 0;
  for (const SegChunk& c : chunks) chunk = c.sites > chunk ? c.sites : chunk;
  const int64_t per_strand = (int64_t)4 * L + 4;
  const int64_t h0_floats = (int64_t)2 * NL * H;
  const bool has_h0 = h0_fwd && h0_rev;
  // two of everything: chunk c+1's h0 upload overlaps chunk c's kernels
  const size_t feat_bytes = (size_t)chunk * 2 * per_strand * sizeof(float);
  const size_t h0_bytes = has_h0 ? (size_t)chunk * 2 * h0_floats * sizeof(float) : 0;
  const size_t out_bytes = (size_t)chunk * 2 * C * sizeof(float);
  const size_t tag_bytes = (size_t)chunk * (sizeof(float) + sizeof(int) + 4);
  CCSM_TRY(ex->feat.reserve(2 * feat_bytes));
  CCSM_TRY(ex->h0.reserve(2 * h0_bytes + 16));
  CCSM_TRY(ex->out.reserve(2 * out_bytes));
  CCSM_TRY(ex->tags.reserve(2 * tag_bytes));
  cudaStream_t st = ex->st, sc = ex->st_copy;
  int rc = CCSM_OK;
  int64_t ci = -1;
  for (const SegChunk& ck : chunks) {
    if (rc != CCSM_OK) break;
    ++ci;
    const int b = (int)(ci & 1);
    const int64_t s0 = ck.site0, cn = ck.sites;
    float* fb = reinterpret_cast<float*>(ex->feat.as<char>() + b * feat_bytes);
    ccsm_strand dev[2];
    float* cur = fb;
    for (int s = 0; s < 2; ++s) {
      memset(&dev[s], 0, sizeof(ccsm_strand));
      dev[s].kmer = cur; cur += cn * L;
      if (has_np) { dev[s].kpass = cur; cur += cn * L; }
      dev[s].ipd_means = cur; cur += cn * L;
      dev[s].pw_means = cur; cur += cn * L;
      if (has_sn) { dev[s].sns = cur; cur += cn * 4; }
    }
    const float* dh0[2] = {nullptr, nullptr};
    if (has_h0) {
      if (ci >= 2) cudaStreamWaitEvent(sc, ex->done[b], 0);
      float* hb = reinterpret_cast<float*>(ex->h0.as<char>() + b * h0_bytes);
      const float* src[2] = {h0_fwd, h0_rev};
      for (int s = 0; s < 2; ++s) {
        float* d = hb + (int64_t)s * cn * h0_floats;
        cudaMemcpy2DAsync(d, (size_t)cn * H * sizeof(float), src[s] + s0 * H, (size_t)n * H * sizeof(float),
                          (size_t)cn * H * sizeof(float), 2 * NL, cudaMemcpyHostToDevice, sc);
        dh0[s] = d;
      }
      cudaEventRecord(ex->ev[b], sc);
      cudaStreamWaitEvent(st, ex->ev[b], 0);
    }
    rc = launch_features(m, s0, cn, &dev[0], &dev[1], st);
    if (rc != CCSM_OK) break;
    float* dl = reinterpret_cast<float*>(ex->out.as<char>() + b * out_bytes);
    float* dp = dl + cn * C;
    rc = forward_att2s_dev(m, cn, &dev[0], &dev[1], dh0[0], dh0[1], dl, dp, st, torch_h0 ? segs.data() + ck.seg0 : nullptr,
                           torch_h0 ? ck.nseg : 0);
    if (rc != CCSM_OK) break;
    char* tb = ex->tags.as<char>() + b * tag_bytes;
    float* d_p1 = reinterpret_cast<float*>(tb);
    int* d_mm = reinterpret_cast<int*>(tb + (size_t)chunk * sizeof(float));
    uint8_t* d_ml = reinterpret_cast<uint8_t*>(tb + (size_t)chunk * (sizeof(float) + sizeof(int)));
    if (prob1 || mm_delta || ml) {
      site_tags_kernel<<<(unsigned)((cn + 255) / 256), 256, 0, st>>>(dp, C, ex->site_read.as<int>(),
                                                                     ex->site_cord.as<int>(), s0, cn, d_p1, d_mm, d_ml);
      count_launch();
    }
    if (logits) cudaMemcpyAsync(logits + s0 * C, dl, (size_t)cn * C * sizeof(float), cudaMemcpyDeviceToHost, st);
    if (probs) cudaMemcpyAsync(probs + s0 * C, dp, (size_t)cn * C * sizeof(float), cudaMemcpyDeviceToHost, st);
    if (prob1) cudaMemcpyAsync(prob1 + s0, d_p1, (size_t)cn * sizeof(float), cudaMemcpyDeviceToHost, st);
    if (mm_delta) cudaMemcpyAsync(mm_delta + s0, d_mm, (size_t)cn * sizeof(int), cudaMemcpyDeviceToHost, st);
    if (ml) cudaMemcpyAsync(ml + s0, d_ml, (size_t)cn, cudaMemcpyDeviceToHost, st);
    cudaEventRecord(ex->done[b], st);
  }
  cudaError_t e1 = cudaStreamSynchronize(sc);
  cudaError_t e2 = cudaStreamSynchronize(st);
  if (rc != CCSM_OK) return rc;
  if (e1 != cudaSuccess || e2 != cudaSuccess) {
    set_error("ccsm_reads_forward_host: %s", cudaGetErrorString(e1 != cudaSuccess ? e1 : e2));
    return CCSM_ECUDA;
  }
  return CCSM_OK;
}

}  // extern "C"
