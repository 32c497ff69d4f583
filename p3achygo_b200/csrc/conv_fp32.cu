// fp32 CUDA-core convolution over the padded board-row layout — the PARITY path
// (P3_PRECISION_FP32): fp32 operands, fp32 FFMA accumulation, accurate mish, so results sit within
// max-abs 1e-3 of the fp32 oracle.  It also serves nets the tcgen05 path cannot tile (`tiny`, C=16).
//
// conv(mish(BN(x))) of python/model.py:276-281 with the BN+mish of the NEXT layer folded into this
// layer's epilogue (ConvEpilogue, common.cuh).  A k x k "same" conv is `taps` row-shifted GEMMs:
//     acc[m, :] = sum_t  in[m + tap_off[t], :] @ w[t]            (rows outside [0, rows) read 0)
// 64x64 output tile per CTA, 16-deep K slabs through shared memory, 4x4 register tile per thread.
#include "common.cuh"
#include "math.cuh"

namespace p3 {
namespace {

constexpr int BM = 64, BN = 64, BK = 16;
constexpr int kMaxTaps = 25;

struct Taps {
  int off[kMaxTaps];
};

__global__ void __launch_bounds__(256)
conv_fp32_kernel(const float* __restrict__ in, const float* __restrict__ w, int rows, int cin, int cout, int taps,
                 Taps tap, const float* residual, float* raw_out, float* act_out, const float* __restrict__ scale,
                 const float* __restrict__ shift, int act_mode, int raw_transposed) {
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const int ty = tid / 16, tx = tid % 16;
  float acc[4][4] = {};

  const int a_row = tid / 4, a_k = (tid % 4) * 4;
  const int b_k = tid / 16, b_n = (tid % 16) * 4;

  for (int t = 0; t < taps; ++t) {
    const long src_row = static_cast<long>(m0) + a_row + tap.off[t];
    const bool row_ok = src_row >= 0 && src_row < rows;
    const float* a_src = in + src_row * cin;
    const float* w_t = w + static_cast<size_t>(t) * cin * cout;
    for (int k0 = 0; k0 < cin; k0 += BK) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int k = k0 + a_k + i;
        As[a_k + i][a_row] = (row_ok && k < cin) ? a_src[k] : 0.0f;
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int k = k0 + b_k, n = n0 + b_n + i;
        Bs[b_k][b_n + i] = (k < cin && n < cout) ? w_t[static_cast<size_t>(k) * cout + n] : 0.0f;
      }
      __syncthreads();
#pragma unroll
      for (int k = 0; k < BK; ++k) {
        float a[4], bb[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) a[i] = As[k][ty * 4 + i];
#pragma unroll
        for (int j = 0; j < 4; ++j) bb[j] = Bs[k][tx * 4 + j];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
      }
      __syncthreads();
    }
  }

#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= rows) continue;
    const bool live = row_is_live(m % kRowsPerPos);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= cout) continue;
      const size_t idx = static_cast<size_t>(m) * cout + n;
      float v = 0.0f;
      if (live) {
        v = acc[i][j];
        if (residual) v += residual[idx];
      }
      if (raw_out) raw_out[raw_transposed ? static_cast<size_t>(n) * rows + m : idx] = v;
      if (act_out) {
        float a = 0.0f;
        if (live) {
          if (act_mode == kActMishBN) a = mish_f32<true>(fmaf(v, scale[n], shift[n]));
          else if (act_mode == kActMish) a = mish_f32<true>(v);
          else a = v;
        }
        act_out[idx] = a;
      }
    }
  }
}

}  // namespace

int conv_fp32_launch(const float* in, const float* w, int rows, int cin, int cout, int taps, const int* tap_off_host,
                     const ConvEpilogue& ep, cudaStream_t stream) {
  if (taps > kMaxTaps) return fail(P3_ERR_INVALID_ARG, "conv_fp32: too many taps");
  Taps tap{};
  for (int t = 0; t < taps; ++t) tap.off[t] = tap_off_host[t];
  dim3 grid((rows + BM - 1) / BM, (cout + BN - 1) / BN);
  conv_fp32_kernel<<<grid, 256, 0, stream>>>(in, w, rows, cin, cout, taps, tap, reinterpret_cast<const float*>(ep.residual),
                                             reinterpret_cast<float*>(ep.raw_out),
                                             reinterpret_cast<float*>(ep.act_out), ep.scale, ep.shift, ep.act_mode,
                                             ep.raw_transposed ? 1 : 0);
  P3_CUDA(cudaGetLastError());
  return P3_OK;
}

}  // namespace p3
