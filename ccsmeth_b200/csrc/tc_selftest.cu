// Single-CTA tcgen05 GEMM used by tests to pin the UMMA descriptor conventions (LBO/SBO meaning,
// instruction descriptor bits, TMEM lane/column mapping of tcgen05.ld) that tc_path.cu relies on.
#include <vector>

#include "ccsm_internal.h"
#include <string.h>

#include "tc_common.cuh"
#include "tmap.h"

namespace ccsm {
using namespace tc;

// A image: K/8 slabs of (128 rows x 16 B); B image: K/8 slabs of (N rows x 16 B).
template <bool F16>
__global__ void __launch_bounds__(128, 1) umma_selftest_kernel(const uint8_t* __restrict__ a_img,
                                                               const uint8_t* __restrict__ b_img, float* __restrict__ D,
                                                               int N, int K, int swap_lbo_sbo) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bars[2];
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t a_bytes = (uint32_t)(K / 8) * 2048u, b_bytes = (uint32_t)(K / 8) * (uint32_t)N * 16u;
  uint8_t* sA = smem;
  uint8_t* sB = smem + a_bytes;
  const uint32_t bar_full = smem_u32(&bars[0]), bar_done = smem_u32(&bars[1]);
  if (threadIdx.x == 0) {
    mbar_init(bar_full, 1);
    mbar_init(bar_done, 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    tmem_alloc(smem_u32(&tmem_base_s), 256);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  if (warp == 0) {
    if (elect_one()) {
      mbar_expect_tx(bar_full, a_bytes + b_bytes);
      bulk_g2s(smem_u32(sA), a_img, a_bytes, bar_full);
      bulk_g2s(smem_u32(sB), b_img, b_bytes, bar_full);
      mbar_wait(bar_full, 0);
      tc_fence_after();
      const uint32_t idesc = make_idesc(128, N, F16);
      const uint32_t a_lbo = 2048, b_lbo = (uint32_t)N * 16u, sbo = 128;
      for (int ks = 0; ks < K / 16; ++ks) {
        uint64_t ad, bd;
        if (!swap_lbo_sbo) {
          ad = make_smem_desc(smem_u32(sA) + ks * 2 * a_lbo, a_lbo, sbo);
          bd = make_smem_desc(smem_u32(sB) + ks * 2 * b_lbo, b_lbo, sbo);
        } else {
          ad = make_smem_desc(smem_u32(sA) + ks * 2 * a_lbo, sbo, a_lbo);
          bd = make_smem_desc(smem_u32(sB) + ks * 2 * b_lbo, sbo, b_lbo);
        }
        umma_f16(tmem, ad, bd, idesc, ks > 0 ? 1u : 0u);
      }
      umma_commit(bar_done);
    }
    __syncwarp();
  }
  mbar_wait(bar_done, 0);
  tc_fence_after();
  const int row = warp * 32 + lane;
  for (int c0 = 0; c0 < N; c0 += 16) {
    uint32_t v[16];
    tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 16; ++i) D[(size_t)row * N + c0 + i] = __uint_as_float(v[i]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 256);
}

// CTA-pair variant: D(256, N) = A(256, K) . B(N, K)^T with tcgen05.mma.cta_group::2.  CTA c holds A rows
// [128c, 128c+128) and B rows [c*N/2, (c+1)*N/2).  Also exercises tcgen05.st (zeroing columns) -> Z.
template <bool F16>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
    umma_pair_selftest_kernel(const uint8_t* __restrict__ a_img, const uint8_t* __restrict__ b_img,
                              float* __restrict__ D, float* __restrict__ Z, int N, int K, int direct_signal,
                              const __grid_constant__ CUtensorMap ta, const __grid_constant__ CUtensorMap tb) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bars[3];  // full (local), peer_full (used in CTA 0), done
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int NH = N / 2;
  const uint32_t a_bytes = (uint32_t)(K / 8) * 2048u, b_bytes = (uint32_t)(K / 8) * (uint32_t)NH * 16u;
  uint8_t* sA = smem;
  uint8_t* sB = smem + a_bytes;
  const uint32_t bar_full = smem_u32(&bars[0]), bar_peer = smem_u32(&bars[1]), bar_done = smem_u32(&bars[2]);
  if (threadIdx.x == 0) {
    mbar_init(bar_full, 1);
    mbar_init(bar_peer, 1);
    mbar_init(bar_done, 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    tmem_alloc2(smem_u32(&tmem_base_s), 256);
    tmem_relinquish2();
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  if (warp == 0) {
    if (elect_one()) {
      if (direct_signal == 2) {
        // tensor-map loads of both CTAs complete on the LEADER's mbarrier (cp.async.bulk.tensor ... .cta_group::2)
        if (rank == 0) mbar_expect_tx(bar_full, 2 * (a_bytes + b_bytes));
        const uint32_t bar = mapa_u32(bar_full, 0);
        tma2d_pair(smem_u32(sA), &ta, 0, (int)rank * (K / 8), bar);
        tma2d_pair(smem_u32(sB), &tb, 0, (int)rank * (K / 8), bar);
        if (rank == 0) mbar_wait(bar_full, 0);
      } else if (direct_signal) {
        // experiment: the peer's bulk copies complete_tx directly on the LEADER's mbarrier (no relay hop)
        if (rank == 0) mbar_expect_tx(bar_full, 2 * (a_bytes + b_bytes));
        const uint32_t bar = rank == 0 ? bar_full : mapa_u32(bar_full, 0);
        bulk_g2s(smem_u32(sA), a_img + (size_t)rank * a_bytes, a_bytes, bar);
        bulk_g2s(smem_u32(sB), b_img + (size_t)rank * b_bytes, b_bytes, bar);
        if (rank == 0) mbar_wait(bar_full, 0);
      } else {
        mbar_expect_tx(bar_full, a_bytes + b_bytes);
        bulk_g2s(smem_u32(sA), a_img + (size_t)rank * a_bytes, a_bytes, bar_full);
        bulk_g2s(smem_u32(sB), b_img + (size_t)rank * b_bytes, b_bytes, bar_full);
        mbar_wait(bar_full, 0);
      }
      if (rank == 1) {
        if (!direct_signal) mbar_arrive_remote(mapa_u32(bar_peer, 0));  // tell the leader our half has landed
      } else {
        if (!direct_signal) mbar_wait(bar_peer, 0);
        tc_fence_after();
        const uint32_t idesc = make_idesc(256, N, F16);
        const uint32_t a_lbo = 2048, b_lbo = (uint32_t)NH * 16u, sbo = 128;
        for (int ks = 0; ks < K / 16; ++ks) {
          const uint64_t ad = make_smem_desc(smem_u32(sA) + ks * 2 * a_lbo, a_lbo, sbo);
          const uint64_t bd = make_smem_desc(smem_u32(sB) + ks * 2 * b_lbo, b_lbo, sbo);
          umma_f16_pair(tmem, ad, bd, idesc, ks > 0 ? 1u : 0u);
        }
        umma_commit_pair(bar_done, 0x3);
      }
    }
    __syncwarp();
  }
  mbar_wait(bar_done, 0);
  tc_fence_after();
  const int row = (int)rank * 128 + warp * 32 + lane;
  const uint32_t trow = tmem + ((uint32_t)(warp * 32) << 16);
  for (int c0 = 0; c0 < N; c0 += 16) {
    uint32_t v[16];
    tmem_ld16(trow + (uint32_t)c0, v);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 16; ++i) D[(size_t)row * N + c0 + i] = __uint_as_float(v[i]);
  }
  {
    uint32_t z[16], v[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) z[i] = 0u;
    tmem_st16(trow + 16, z);  // zero columns [16, 32)
    tmem_st_wait();
    tmem_ld16(trow + 16, v);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 16; ++i) Z[(size_t)row * 32 + i] = __uint_as_float(v[i]);
    tmem_ld16(trow + 0, v);  // neighbours must be untouched
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 16; ++i) Z[(size_t)row * 32 + 16 + i] = __uint_as_float(v[i]);
  }
  tc_fence_before();
  cluster_sync_all();
  if (warp == 0) tmem_dealloc2(tmem, 256);
}

static void pack_rows(const float* src, int rows, int K, bool f16, std::vector<uint16_t>& img) {
  img.assign((size_t)rows * K, 0);
  for (int r = 0; r < rows; ++r)
    for (int k = 0; k < K; ++k) {
      size_t off = (size_t)(k / 8) * rows * 8 + (size_t)r * 8 + (k % 8);
      float v = src[(size_t)r * K + k];
      uint16_t bits;
      if (f16) {
        __half h = __float2half_rn(v);
        bits = *reinterpret_cast<uint16_t*>(&h);
      } else {
        __nv_bfloat16 h = __float2bfloat16_rn(v);
        bits = *reinterpret_cast<uint16_t*>(&h);
      }
      img[off] = bits;
    }
}


// Mixed-kind accumulation: D = fp16(A) . fp16(B)^T [kind::f16, K = 16 per MMA] + e4m3(A) . e4m3(B)^T [kind::f8f6f4,
// K = 32 per MMA] into the same fp32 TMEM columns.  Pins the 8-bit operand layout (slab = 16 K elements, rows 16 B
// apart, LBO = slab bytes) and that the two kinds may share an accumulator -- what the fp16c8 mode relies on.
__global__ void __launch_bounds__(128, 1) umma_mixed_selftest_kernel(const uint8_t* __restrict__ a16, const uint8_t* __restrict__ b16,
                                                                     const uint8_t* __restrict__ a8, const uint8_t* __restrict__ b8,
                                                                     float* __restrict__ D, int N, int K) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bars[2];
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t a16_bytes = (uint32_t)(K / 8) * 2048u, b16_bytes = (uint32_t)(K / 8) * (uint32_t)N * 16u;
  const uint32_t a8_bytes = a16_bytes / 2, b8_bytes = b16_bytes / 2;
  uint8_t* sA = smem;
  uint8_t* sB = sA + a16_bytes;
  uint8_t* sA8 = sB + b16_bytes;
  uint8_t* sB8 = sA8 + a8_bytes;
  const uint32_t bar_full = smem_u32(&bars[0]), bar_done = smem_u32(&bars[1]);
  if (threadIdx.x == 0) {
    mbar_init(bar_full, 1);
    mbar_init(bar_done, 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    tmem_alloc(smem_u32(&tmem_base_s), 256);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  if (warp == 0) {
    if (elect_one()) {
      mbar_expect_tx(bar_full, a16_bytes + b16_bytes + a8_bytes + b8_bytes);
      bulk_g2s(smem_u32(sA), a16, a16_bytes, bar_full);
      bulk_g2s(smem_u32(sB), b16, b16_bytes, bar_full);
      bulk_g2s(smem_u32(sA8), a8, a8_bytes, bar_full);
      bulk_g2s(smem_u32(sB8), b8, b8_bytes, bar_full);
      mbar_wait(bar_full, 0);
      tc_fence_after();
      const uint32_t idesc = make_idesc(128, N, true), idesc8 = make_idesc_e4m3(128, N);
      const uint32_t a_lbo = 2048, b_lbo = (uint32_t)N * 16u, sbo = 128;
      for (int g = 0; g < K / 32; ++g) {  // interleaved like the layer kernel: 2 x f16, then 1 x e4m3 per 32 K elements
        for (int ks = 2 * g; ks < 2 * g + 2; ++ks)
          umma_f16(tmem, make_smem_desc(smem_u32(sA) + ks * 2 * a_lbo, a_lbo, sbo),
                   make_smem_desc(smem_u32(sB) + ks * 2 * b_lbo, b_lbo, sbo), idesc, ks > 0 ? 1u : 0u);
        umma_f8(tmem, make_smem_desc(smem_u32(sA8) + g * 2 * a_lbo, a_lbo, sbo),
                make_smem_desc(smem_u32(sB8) + g * 2 * b_lbo, b_lbo, sbo), idesc8, 1u);
      }
      umma_commit(bar_done);
    }
    __syncwarp();
  }
  mbar_wait(bar_done, 0);
  tc_fence_after();
  const int row = warp * 32 + lane;
  for (int c0 = 0; c0 < N; c0 += 16) {
    uint32_t v[16];
    tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 16; ++i) D[(size_t)row * N + c0 + i] = __uint_as_float(v[i]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 256);
}


// Tensor-pipe rate probe: one CTA issues `iters` rounds of MMAs (M = 128, N columns) on fixed shared-memory operands and
// reports the cycles until the commit lands.  mode 0: 4 x kind::f16 per round; 1: 4 x kind::f8f6f4 (e4m3);
// 2: 2 x f16 then 2 x e4m3 (the fp16c8 stage pattern); 3: f16, e4m3, f16, e4m3; 4: rounds alternate f16 x4 / e4m3 x4.
__global__ void __launch_bounds__(128, 1) umma_rate_kernel(int N, int mode, int iters, long long* __restrict__ cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bars[1];
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 65536 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  const uint32_t bar_done = smem_u32(&bars[0]);
  if (threadIdx.x == 0) {
    mbar_init(bar_done, 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    tmem_alloc(smem_u32(&tmem_base_s), 256);
    tmem_relinquish();
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  if (warp == 0) {
    if (elect_one()) {
      const uint32_t idesc = make_idesc(128, N, true), idesc8 = make_idesc_e4m3(128, N);
      const uint32_t sA = smem_u32(smem), sB = smem_u32(smem) + 16384;
      const uint32_t a_lbo = 2048, b_lbo = (uint32_t)N * 16u, sbo = 128;
      const long long t0 = clock64();
      for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          bool f8;
          if (mode == 0) f8 = false;
          else if (mode == 1) f8 = true;
          else if (mode == 2) f8 = q >= 2;
          else if (mode == 3) f8 = q & 1;
          else f8 = it & 1;
          const uint64_t ad = make_smem_desc(sA + q * 2 * a_lbo, a_lbo, sbo), bd = make_smem_desc(sB + q * 2 * b_lbo, b_lbo, sbo);
          if (f8) umma_f8(tmem, ad, bd, idesc8, 1u);
          else umma_f16(tmem, ad, bd, idesc, 1u);
        }
      }
      umma_commit(bar_done);
      mbar_wait(bar_done, 0);
      *cycles = clock64() - t0;
    }
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 256);
}


// Same probe for a CTA pair (tcgen05.mma.cta_group::2, M = 256, each CTA holds N/2 rows of B): mode 0 = kind::f16.
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) umma_rate_pair_kernel(int N, int mode, int iters,
                                                                                        long long* __restrict__ cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bars[1];
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5;
  const uint32_t rank = cluster_ctarank();
  for (int i = threadIdx.x; i < 65536 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  const uint32_t bar_done = smem_u32(&bars[0]);
  if (threadIdx.x == 0) {
    mbar_init(bar_done, 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    tmem_alloc2(smem_u32(&tmem_base_s), 256);
    tmem_relinquish2();
  }
  fence_proxy_async_smem();
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  if (warp == 0 && rank == 0) {
    if (elect_one()) {
      const uint32_t idesc = make_idesc(256, N, true), idesc8 = make_idesc_e4m3(256, N);
      const uint32_t sA = smem_u32(smem), sB = smem_u32(smem) + 16384;
      const uint32_t a_lbo = 2048, b_lbo = (uint32_t)(N / 2) * 16u, sbo = 128;
      const long long t0 = clock64();
      for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const bool f8 = mode == 1 || (mode == 3 && (q & 1));
          const uint64_t ad = make_smem_desc(sA + q * 2 * a_lbo, a_lbo, sbo), bd = make_smem_desc(sB + q * 2 * b_lbo, b_lbo, sbo);
          if (f8) umma_f8_pair(tmem, ad, bd, idesc8, 1u);
          else umma_f16_pair(tmem, ad, bd, idesc, 1u);
        }
      }
      umma_commit_pair(bar_done, 0x3);
      mbar_wait(bar_done, 0);
      *cycles = clock64() - t0;
    }
    __syncwarp();
  } else if (warp == 0) {
    mbar_wait(bar_done, 0);
  }
  tc_fence_before();
  cluster_sync_all();
  if (warp == 0) tmem_dealloc2(tmem, 256);
}

// rows x K floats -> e4m3 image: K/16 slabs of (rows x 16 B)
static void pack_rows_e4m3(const float* src, int rows, int K, std::vector<uint8_t>& img) {
  img.assign((size_t)rows * K, 0);
  for (int r = 0; r < rows; ++r)
    for (int k = 0; k < K; ++k)
      img[(size_t)(k / 16) * rows * 16 + (size_t)r * 16 + (k % 16)] =
          (uint8_t)__nv_cvt_float_to_fp8(src[(size_t)r * K + k], __NV_SATFINITE, __NV_E4M3);
}

}  // namespace ccsm

using namespace ccsm;

extern "C" int ccsm_debug_umma_mixed_gemm(int32_t device, int32_t N, int32_t K, const float* A, const float* B, float* D) {
  if (N < 16 || N > 256 || N % 16 || K < 32 || K % 32 || !A || !B || !D) {
    set_error("ccsm_debug_umma_mixed_gemm: bad shape N=%d K=%d", N, K);
    return CCSM_EINVAL;
  }
  const size_t smem = (size_t)(K / 8) * (2048 + (size_t)N * 16) * 3 / 2;
  if (smem > 200 * 1024) {
    set_error("ccsm_debug_umma_mixed_gemm: K too large for one stage");
    return CCSM_EINVAL;
  }
  CCSM_CUDA(cudaSetDevice(device));
  std::vector<uint16_t> ai, bi;
  std::vector<uint8_t> a8, b8;
  pack_rows(A, 128, K, true, ai);
  pack_rows(B, N, K, true, bi);
  pack_rows_e4m3(A, 128, K, a8);
  pack_rows_e4m3(B, N, K, b8);
  DevBuf da, db, da8, db8, dd;
  CCSM_TRY(da.reserve(ai.size() * 2));
  CCSM_TRY(db.reserve(bi.size() * 2));
  CCSM_TRY(da8.reserve(a8.size()));
  CCSM_TRY(db8.reserve(b8.size()));
  CCSM_TRY(dd.reserve((size_t)128 * N * 4));
  CCSM_CUDA(cudaMemcpy(da.p, ai.data(), ai.size() * 2, cudaMemcpyHostToDevice));
  CCSM_CUDA(cudaMemcpy(db.p, bi.data(), bi.size() * 2, cudaMemcpyHostToDevice));
  CCSM_CUDA(cudaMemcpy(da8.p, a8.data(), a8.size(), cudaMemcpyHostToDevice));
  CCSM_CUDA(cudaMemcpy(db8.p, b8.data(), b8.size(), cudaMemcpyHostToDevice));
  CCSM_CUDA(cudaFuncSetAttribute(umma_mixed_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  umma_mixed_selftest_kernel<<<1, 128, smem>>>(da.as<uint8_t>(), db.as<uint8_t>(), da8.as<uint8_t>(), db8.as<uint8_t>(),
                                               dd.as<float>(), N, K);
  count_launch();
  cudaError_t e = cudaDeviceSynchronize();
  if (e == cudaSuccess) e = cudaMemcpy(D, dd.p, (size_t)128 * N * 4, cudaMemcpyDeviceToHost);
  da.release(); db.release(); da8.release(); db8.release(); dd.release();
  if (e != cudaSuccess) {
    set_error("umma mixed selftest kernel failed: %s", cudaGetErrorString(e));
    return CCSM_ECUDA;
  }
  return CCSM_OK;
}

extern "C" int ccsm_debug_umma_pair_gemm(int32_t device, int32_t N, int32_t K, int32_t is_f16, const float* A,
                                         const float* B, float* D, float* Z) {
  // is_f16 bit 1 selects the "peer signals the leader's mbarrier directly" experiment with plain bulk copies (fails: a bulk
  // copy can only complete on a barrier of the destination CTA), bit 2 the same with cta_group::2 tensor-map loads
  const int direct_signal = (is_f16 & 4) ? 2 : (is_f16 >> 1) & 1;
  is_f16 &= 1;
  if (N < 32 || N > 256 || N % 32 || K < 16 || K % 16 || !A || !B || !D || !Z) {
    set_error("ccsm_debug_umma_pair_gemm: bad shape N=%d K=%d", N, K);
    return CCSM_EINVAL;
  }
  const int NH = N / 2;
  size_t smem = (size_t)(K / 8) * (2048 + (size_t)NH * 16);
  if (smem > 200 * 1024) {
    set_error("ccsm_debug_umma_pair_gemm: K too large");
    return CCSM_EINVAL;
  }
  CCSM_CUDA(cudaSetDevice(device));
  std::vector<uint16_t> ai, a0, a1, b0, b1;
  pack_rows(A, 128, K, is_f16 != 0, a0);
  pack_rows(A + (size_t)128 * K, 128, K, is_f16 != 0, a1);
  pack_rows(B, NH, K, is_f16 != 0, b0);
  pack_rows(B + (size_t)NH * K, NH, K, is_f16 != 0, b1);
  ai = a0; ai.insert(ai.end(), a1.begin(), a1.end());
  std::vector<uint16_t> bi = b0; bi.insert(bi.end(), b1.begin(), b1.end());
  DevBuf da, db, dd, dz;
  CCSM_TRY(da.reserve(ai.size() * 2));
  CCSM_TRY(db.reserve(bi.size() * 2));
  CCSM_TRY(dd.reserve((size_t)256 * N * 4));
  CCSM_TRY(dz.reserve((size_t)256 * 32 * 4));
  CCSM_CUDA(cudaMemcpy(da.p, ai.data(), ai.size() * 2, cudaMemcpyHostToDevice));
  CCSM_CUDA(cudaMemcpy(db.p, bi.data(), bi.size() * 2, cudaMemcpyHostToDevice));
  CUtensorMap ta, tb;
  memset(&ta, 0, sizeof(ta));
  memset(&tb, 0, sizeof(tb));
  if (direct_signal == 2) {
    if (K / 8 > 256 || make_slab_tmap(&ta, da.p, 2048, (uint64_t)2 * (K / 8), (uint32_t)(K / 8)) ||
        make_slab_tmap(&tb, db.p, (uint32_t)NH * 16u, (uint64_t)2 * (K / 8), (uint32_t)(K / 8))) {
      set_error("ccsm_debug_umma_pair_gemm: cuTensorMapEncodeTiled failed");
      return CCSM_ECUDA;
    }
  }
  if (is_f16) {
    CCSM_CUDA(cudaFuncSetAttribute(umma_pair_selftest_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    umma_pair_selftest_kernel<true><<<2, 128, smem>>>(da.as<uint8_t>(), db.as<uint8_t>(), dd.as<float>(), dz.as<float>(), N, K, direct_signal, ta, tb);
  } else {
    CCSM_CUDA(cudaFuncSetAttribute(umma_pair_selftest_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    umma_pair_selftest_kernel<false><<<2, 128, smem>>>(da.as<uint8_t>(), db.as<uint8_t>(), dd.as<float>(), dz.as<float>(), N, K, direct_signal, ta, tb);
  }
  count_launch();
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    set_error("umma pair selftest kernel failed: %s", cudaGetErrorString(e));
    return CCSM_ECUDA;
  }
  CCSM_CUDA(cudaMemcpy(D, dd.p, (size_t)256 * N * 4, cudaMemcpyDeviceToHost));
  CCSM_CUDA(cudaMemcpy(Z, dz.p, (size_t)256 * 32 * 4, cudaMemcpyDeviceToHost));
  da.release(); db.release(); dd.release(); dz.release();
  return CCSM_OK;
}

extern "C" int ccsm_debug_umma_gemm(int32_t device, int32_t N, int32_t K, int32_t is_f16, int32_t swap_lbo_sbo,
                                    const float* A, const float* B, float* D) {
  if (N < 16 || N > 256 || N % 16 || K < 16 || K % 16 || !A || !B || !D) {
    set_error("ccsm_debug_umma_gemm: bad shape N=%d K=%d", N, K);
    return CCSM_EINVAL;
  }
  size_t smem = (size_t)(K / 8) * (2048 + (size_t)N * 16);
  if (smem > 200 * 1024) {
    set_error("ccsm_debug_umma_gemm: K too large for one stage");
    return CCSM_EINVAL;
  }
  CCSM_CUDA(cudaSetDevice(device));
  std::vector<uint16_t> ai, bi;
  pack_rows(A, 128, K, is_f16 != 0, ai);
  pack_rows(B, N, K, is_f16 != 0, bi);
  DevBuf da, db, dd;
  CCSM_TRY(da.reserve(ai.size() * 2));
  CCSM_TRY(db.reserve(bi.size() * 2));
  CCSM_TRY(dd.reserve((size_t)128 * N * 4));
  CCSM_CUDA(cudaMemcpy(da.p, ai.data(), ai.size() * 2, cudaMemcpyHostToDevice));
  CCSM_CUDA(cudaMemcpy(db.p, bi.data(), bi.size() * 2, cudaMemcpyHostToDevice));
  if (is_f16) {
    CCSM_CUDA(cudaFuncSetAttribute(umma_selftest_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    umma_selftest_kernel<true><<<1, 128, smem>>>(da.as<uint8_t>(), db.as<uint8_t>(), dd.as<float>(), N, K, swap_lbo_sbo);
  } else {
    CCSM_CUDA(cudaFuncSetAttribute(umma_selftest_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    umma_selftest_kernel<false><<<1, 128, smem>>>(da.as<uint8_t>(), db.as<uint8_t>(), dd.as<float>(), N, K, swap_lbo_sbo);
  }
  count_launch();
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    set_error("umma selftest kernel failed: %s", cudaGetErrorString(e));
    da.release(); db.release(); dd.release();
    return CCSM_ECUDA;
  }
  CCSM_CUDA(cudaMemcpy(D, dd.p, (size_t)128 * N * 4, cudaMemcpyDeviceToHost));
  da.release(); db.release(); dd.release();
  return CCSM_OK;
}

extern "C" int ccsm_debug_umma_rate(int32_t device, int32_t N, int32_t mode, int32_t iters, int64_t* cycles) {
  if (N < 16 || N > 256 || N % 32 || mode < 0 || (mode > 4 && mode < 16) || mode > 19 || iters < 1 || !cycles) {
    set_error("ccsm_debug_umma_rate: bad argument");
    return CCSM_EINVAL;
  }
  CCSM_CUDA(cudaSetDevice(device));
  DevBuf dc;
  CCSM_TRY(dc.reserve(8));
  const int smem = 65536;
  CCSM_CUDA(cudaFuncSetAttribute(umma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  if (mode >= 16) {  // bit 4: CTA-pair probe (modes 0, 1, 3)
    CCSM_CUDA(cudaFuncSetAttribute(umma_rate_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    umma_rate_pair_kernel<<<2, 128, smem>>>(N, mode & 15, iters, dc.as<long long>());
  } else
  umma_rate_kernel<<<1, 128, smem>>>(N, mode, iters, dc.as<long long>());
  count_launch();
  cudaError_t e = cudaDeviceSynchronize();
  long long c = 0;
  if (e == cudaSuccess) e = cudaMemcpy(&c, dc.p, 8, cudaMemcpyDeviceToHost);
  dc.release();
  if (e != cudaSuccess) {
    set_error("umma rate kernel failed: %s", cudaGetErrorString(e));
    return CCSM_ECUDA;
  }
  *cycles = c;
  return CCSM_OK;
}
