// BroadcastPreAct of the trunk's broadcast residual block (python/model.py:570-581):
//     per channel c:  y[b, :, c] = Dense_361->361( mish(t[b, :, c]) )      one shared [361,361] kernel + bias
// In NHWC this needs no transpose: per position it is the GEMM  Y_b[q, c] = sum_p W[p, q] * X_b[p, c] + bias[q]
// with the shared matrix as the stationary operand.  X = mish(t) was written by the
// preceding conv's epilogue (kActMish); this kernel folds the following conv's BN + mish into its own
// epilogue, and writes the zero halo rows of the padded board-row layout.
//
// v0: CUDA-core tiles (64 q x 64 c per CTA, fp32 accumulate) for both precision modes.
#include "common.cuh"
#include "math.cuh"

namespace p3 {
namespace {

constexpr int BQ = 64, BC = 64, BK = 16;

template <typename T>
__device__ __forceinline__ float to_f32(T v);
template <>
__device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <>
__device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T>
__device__ __forceinline__ T from_f32(float v);
template <>
__device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <>
__device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16(v); }

template <typename T, bool kAccurate>
__global__ void __launch_bounds__(256)
broadcast_kernel(const T* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias, int C,
                 T* __restrict__ act_out, const float* __restrict__ scale, const float* __restrict__ shift) {
  __shared__ float Ws[BK][BQ + 4];
  __shared__ float Xs[BK][BC + 4];
  const int tid = threadIdx.x;
  const int q0 = blockIdx.x * BQ, c0 = blockIdx.y * BC, b = blockIdx.z;
  const size_t row0 = static_cast<size_t>(b) * kRowsPerPos;
  const int ty = tid / 16, tx = tid % 16;
  float acc[4][4] = {};
  const int l_k = tid / 16, l_n = (tid % 16) * 4;

  for (int p0 = 0; p0 < P3_NUM_BOARD_LOCS; p0 += BK) {
    const int p = p0 + l_k;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int q = q0 + l_n + i;
      Ws[l_k][l_n + i] = (p < P3_NUM_BOARD_LOCS && q < P3_NUM_BOARD_LOCS) ? w[p * P3_NUM_BOARD_LOCS + q] : 0.0f;
      const int c = c0 + l_n + i;
      Xs[l_k][l_n + i] = (p < P3_NUM_BOARD_LOCS && c < C) ? to_f32<T>(x[(row0 + board_row(p)) * C + c]) : 0.0f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float a[4], bb[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = Ws[k][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) bb[j] = Xs[k][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
    }
    __syncthreads();
  }

#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int q = q0 + ty * 4 + i;
    if (q >= P3_NUM_BOARD_LOCS) continue;
    const float bq = bias[q];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = c0 + tx * 4 + j;
      if (c >= C) continue;
      const float y = acc[i][j] + bq;
      act_out[(row0 + board_row(q)) * C + c] = from_f32<T>(mish_f32<kAccurate>(fmaf(y, scale[c], shift[c])));
    }
  }
  if (blockIdx.x == 0) {  // zero halo rows of this channel chunk
    for (int e = tid; e < kRowsPerPos * BC; e += blockDim.x) {
      const int q = e / BC, c = c0 + e % BC;
      if (!row_is_live(q) && c < C) act_out[(row0 + q) * C + c] = from_f32<T>(0.0f);
    }
  }
}

}  // namespace

int broadcast_launch(const void* x, const float* w, const float* bias, int n, int C, void* act_out, bool bf16,
                     const float* scale, const float* shift, cudaStream_t stream) {
  dim3 grid((P3_NUM_BOARD_LOCS + BQ - 1) / BQ, (C + BC - 1) / BC, n);
  if (bf16)
    broadcast_kernel<__nv_bfloat16, false><<<grid, 256, 0, stream>>>(
        reinterpret_cast<const __nv_bfloat16*>(x), w, bias, C, reinterpret_cast<__nv_bfloat16*>(act_out), scale, shift);
  else
    broadcast_kernel<float, true><<<grid, 256, 0, stream>>>(reinterpret_cast<const float*>(x), w, bias, C,
                                                            reinterpret_cast<float*>(act_out), scale, shift);
  P3_CUDA(cudaGetLastError());
  return P3_OK;
}

}  // namespace p3
