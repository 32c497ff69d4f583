// Activation math shared by the kernels.
#pragma once
#include <cuda_runtime.h>

namespace p3 {

// keras.activations.mish: x * tanh(softplus(x))  (python/model.py:269-281).
//   kAccurate  : libm-grade expf / log1pf / tanhf — the fp32 parity path.
//   !kAccurate : tanh(ln(1+e^x)) = n / (n + 2) with n = e^x (e^x + 2): one ex2 + one rcp on the
//                SFU; |rel err| ~1e-6, far below bf16 operand rounding.  Branch-free and flush-to-zero
//                (ex2/rcp.approx.ftz): the non-ftz __expf / __fdividef forms expand to predicated
//                denormal fix-ups and, with the x > 20 early-out, to a branch per element — measured
//                3700 cycles per 16-element epilogue chunk on B200 against ~1000 for this form.
//                x is clamped to 20 inside the exponential (n / (n + 2) == 1.0f there), so no overflow.
__device__ __forceinline__ float ex2_approx_ftz(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_approx_ftz(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
template <bool kAccurate>
__device__ __forceinline__ float mish_f32(float x) {
  if (kAccurate) {
    const float sp = x > 20.0f ? x : log1pf(expf(x));
    return x * tanhf(sp);
  } else {
    const float e = ex2_approx_ftz(fminf(x, 20.0f) * 1.4426950408889634f);
    const float n = fmaf(e, e, e + e);
    return x * (n * rcp_approx_ftz(n + 2.0f));
  }
}

// Two fast mishes sharing ONE reciprocal: 1/d1 = d2 / (d1 d2), 1/d2 = d1 / (d1 d2).  The SFU (16 lanes / clk / SM) is the
// scarce unit in the conv epilogues (2 MUFU ops per element above); this form needs 1.5.  d = n + 2 is in [2, e^40 + 2],
// so d1 * d2 <= 5.6e34 stays finite.
// tanh(softplus(x)) = n / (n + 2) = 1 - 2 / d with d = e^x (e^x + 2) + 2, so mish(x) = x - 2 x / d (absolute error of the
// 1 - 2/d cancellation is ~1e-7 |x|, irrelevant next to the bf16 rounding of the result).
__device__ __forceinline__ void mish2_f32(float x1, float x2, float& y1, float& y2) {
  const float e1 = ex2_approx_ftz(fminf(x1, 20.0f) * 1.4426950408889634f);
  const float e2 = ex2_approx_ftz(fminf(x2, 20.0f) * 1.4426950408889634f);
  const float d1 = fmaf(e1, e1 + 2.0f, 2.0f), d2 = fmaf(e2, e2 + 2.0f, 2.0f);
  const float r = rcp_approx_ftz(d1 * d2);
  y1 = fmaf(-2.0f * x1, d2 * r, x1);
  y2 = fmaf(-2.0f * x2, d1 * r, x2);
}

__device__ __forceinline__ float softplus_f32(float x) { return x > 20.0f ? x : log1pf(expf(x)); }
__device__ __forceinline__ float sigmoid_f32(float x) { return 1.0f / (1.0f + expf(-x)); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

}  // namespace p3
