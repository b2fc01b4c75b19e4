// Tensor maps over the byte images (host side): cuTensorMapEncodeTiled through the runtime's driver entry point, so that the
// library needs no -lcuda.  The images are plain byte arrays of fixed-size slabs; a map views one as a 2-D array of 8-byte
// elements [rows = slabs][inner = slab bytes / 8] (inner <= 256 elements = 2048 B, the box limit), no swizzle, no interleave:
// a box of (inner, n) lands in shared memory as n * slab bytes, contiguous -- the same bytes a 1-D bulk copy would bring.
// Used by the CTA-pair kernels, whose loads must complete on the LEADER CTA's mbarrier (cp.async.bulk.tensor ... .cta_group::2;
// plain bulk copies can only signal a barrier in the destination CTA).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace ccsm {

typedef CUresult (*tmap_encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline tmap_encode_fn tmap_encoder() {
  static tmap_encode_fn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<tmap_encode_fn>(p);
  }
  return fn;
}

// base: 16-byte aligned; slab_bytes: multiple of 16, <= 2048; box_slabs <= 256.  Returns 0 on success.
inline int make_slab_tmap(CUtensorMap* m, const void* base, uint32_t slab_bytes, uint64_t n_slabs, uint32_t box_slabs) {
  tmap_encode_fn enc = tmap_encoder();
  if (!enc || slab_bytes % 16 || slab_bytes > 2048 || box_slabs == 0 || box_slabs > 256 || n_slabs == 0) return -1;
  const cuuint64_t dims[2] = {slab_bytes / 8, n_slabs};
  const cuuint64_t strides[1] = {slab_bytes};
  const cuuint32_t box[2] = {slab_bytes / 8, box_slabs};
  const cuuint32_t estr[2] = {1, 1};
  // 8-byte elements: no 64-bit integer type among the tensor-map data types; FLOAT64 moves the same bytes (no arithmetic,
  // no OOB fill is ever generated: every box lies inside the array)
  const CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<void*>(base), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -2;
}

}  // namespace ccsm
