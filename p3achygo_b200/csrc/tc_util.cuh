// Helpers shared by the tcgen05 kernels: TMA tensor-map construction (driver entry point looked up at run time, so
// libp3b200 links only the CUDA runtime) and operand packing.
#pragma once
#include <cuda.h>

#include <cstdlib>
#include <string>

#include "common.cuh"
#include "math.cuh"
#include "ptx.cuh"

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

// Launch `kern` with programmatic stream serialization (see ptx::griddep_wait); P3_PDL=0 falls back to a plain launch.
template <typename... KArgs, typename... Args>
inline cudaError_t tc_launch_pdl(void (*kern)(KArgs...), int grid, int block, size_t smem, cudaStream_t stream, Args&&... args) {
  const char* pdl_env = std::getenv("P3_PDL");  // read per launch (launches are captured into a graph once per engine)
  const bool enabled = !(pdl_env && std::atoi(pdl_env) == 0);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(static_cast<unsigned>(grid));
  cfg.blockDim = dim3(static_cast<unsigned>(block));
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = enabled ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
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

// activated operand pair in the engine's operand format (bf16, or IEEE fp16 when f16 != 0; both saturate instead of overflowing)
__device__ __forceinline__ uint32_t tc_pack_act(float a, float b, int f16) { return f16 ? tc_pack_f16(a, b) : tc_pack_bf16(a, b); }

// 8 activations a[i] = mish(t[i]), t = (x * scale + shift), with the constants pre-multiplied by log2(e): z = t log2(e) comes
// straight out of the BN FFMA, e^t = ex2(z), and mish(t) = t (1 - 2/d) = z * (ln2 - 2 ln2 / d), d = e^t (e^t + 2) + 2.  Pairs
// share one reciprocal (1/d0 = d1 / (d0 d1)).  Written phase by phase over the 8 values so that the 8 dependency chains are
// interleaved (the SFU latency of one is covered by the others) instead of running back to back.
__device__ __forceinline__ void bn_mish8(const float* x, float* a, uint32_t sc, uint32_t sh, int col0) {
  const float4 s0 = ptx::lds_f4_const(sc + static_cast<uint32_t>(col0) * 4u), s1 = ptx::lds_f4_const(sc + static_cast<uint32_t>(col0 + 4) * 4u);
  const float4 h0 = ptx::lds_f4_const(sh + static_cast<uint32_t>(col0) * 4u), h1 = ptx::lds_f4_const(sh + static_cast<uint32_t>(col0 + 4) * 4u);
  const float scv[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
  const float shv[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
  float z[8], d[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) z[i] = fmaf(x[i], scv[i], shv[i]);
#pragma unroll
  for (int i = 0; i < 8; ++i) d[i] = ex2_approx_ftz(fminf(z[i], 28.853900817779268f));  // e^t, t clamped to 20
#pragma unroll
  for (int i = 0; i < 8; ++i) d[i] = fmaf(d[i], d[i] + 2.0f, 2.0f);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float r = rcp_approx_ftz(d[2 * i] * d[2 * i + 1]);
    const float q0 = d[2 * i + 1] * r, q1 = d[2 * i] * r;  // 1/d0, 1/d1
    a[2 * i] = z[2 * i] * fmaf(q0, -1.3862943611198906f, 0.6931471805599453f);
    a[2 * i + 1] = z[2 * i + 1] * fmaf(q1, -1.3862943611198906f, 0.6931471805599453f);
  }
}

}  // namespace p3
