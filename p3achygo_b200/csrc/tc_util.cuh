// Helpers shared by the tcgen05 kernels: TMA tensor-map construction (driver entry point looked up at run time, so
// libp3b200 links only the CUDA runtime) and operand packing.
#pragma once
#include <cuda.h>

#include <string>

#include "common.cuh"

namespace p3 {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn tc_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess ||
      qres != cudaDriverEntryPointSuccess)
    return nullptr;
  fn = reinterpret_cast<EncodeTiledFn>(p);
  return fn;
}

// 2-D row-major [dim1, dim0] tensor of `elem_bytes` elements, box [box1, box0], zero OOB fill.
inline int tc_make_map_2d(CUtensorMap* map, const void* base, CUtensorMapDataType dt, int elem_bytes, uint64_t dim0,
                          uint64_t dim1, uint32_t box0, uint32_t box1, CUtensorMapSwizzle swz) {
  EncodeTiledFn fn = tc_encode_fn();
  if (!fn) return fail(P3_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
  cuuint64_t gdim[2] = {dim0, dim1};
  cuuint64_t gstride[1] = {dim0 * elem_bytes};
  cuuint32_t box[2] = {box0, box1};
  cuuint32_t estride[2] = {1, 1};
  CUresult r = fn(map, dt, 2, const_cast<void*>(base), gdim, gstride, box, estride, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(P3_ERR_CUDA, "cuTensorMapEncodeTiled failed: " + std::to_string(r));
  return P3_OK;
}

__device__ __forceinline__ uint32_t tc_pack_bf16(float a, float b) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}
// two floats -> packed IEEE fp16 pair, saturating to +-65504 (the residual stream never overflows to inf)
__device__ __forceinline__ uint32_t tc_pack_f16(float a, float b) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
  return r;
}

}  // namespace p3
