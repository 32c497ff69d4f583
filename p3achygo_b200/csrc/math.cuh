// Activation math shared by the kernels.
#pragma once
#include <cuda_runtime.h>

namespace p3 {

// keras.activations.mish: x * tanh(softplus(x))  (python/model.py:269-281).
//   kAccurate  : libm-grade expf / log1pf / tanhf — the fp32 parity path.
//   !kAccurate : tanh(ln(1+e^x)) = n / (n + 2) with n = e^x (e^x + 2): one ex2 + one rcp on the
//                SFU; |rel err| ~1e-6, far below bf16 operand rounding.
template <bool kAccurate>
__device__ __forceinline__ float mish_f32(float x) {
  if (kAccurate) {
    const float sp = x > 20.0f ? x : log1pf(expf(x));
    return x * tanhf(sp);
  } else {
    if (x > 20.0f) return x;
    const float e = __expf(x);
    const float n = e * (e + 2.0f);
    return x * __fdividef(n, n + 2.0f);
  }
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
