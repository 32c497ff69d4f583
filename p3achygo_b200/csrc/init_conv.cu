// First layer of the tower (python/model.py:1230-1237):
//     x = conv5x5(planes, 15 -> C, same, no bias) + dense(game_state, 8 -> C)   (broadcast over HW)
// The planes are exactly {0,1}, so the 375-deep contraction is a sparse gather-add: for every
// in-board tap, add the weight rows of the set plane bits.  Reading the 15-bit masks the encode
// kernel produced (722 B / position) instead of fp32 planes (21 660 B) keeps this layer off HBM;
// a typical point touches ~15 of the 375 rows.  fp32 accumulation in both precision modes.
//
// One CTA per position, one warp per point (strided), lanes over channels.  Writes the raw fp32
// residual stream and the activated copy  mish(BN_0(x))  that the first trunk conv consumes, both in
// the padded board-row layout (common.cuh), including the zero halo rows.
#include <algorithm>

#include "common.cuh"
#include "math.cuh"

namespace p3 {
namespace {

constexpr int kMaxCPerLane = 12;  // C <= 384

template <bool kBf16>
__global__ void __launch_bounds__(256)
init_conv_kernel(const uint16_t* __restrict__ masks, const float* __restrict__ scalars, int nplanes, int nscalars,
                 int C, const float* __restrict__ wt, const float* __restrict__ gs_w, const float* __restrict__ gs_b,
                 void* __restrict__ raw_out, void* __restrict__ act_out, const float* __restrict__ scale,
                 const float* __restrict__ shift) {
  extern __shared__ float s_gs[];  // [C] game-state bias of this position
  __shared__ uint16_t s_mask[P3_NUM_BOARD_LOCS];
  const int b = blockIdx.x;
  for (int p = threadIdx.x; p < P3_NUM_BOARD_LOCS; p += blockDim.x)
    s_mask[p] = masks[static_cast<size_t>(b) * P3_NUM_BOARD_LOCS + p];
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float acc = gs_b[c];
    for (int s = 0; s < nscalars; ++s) acc = fmaf(scalars[b * nscalars + s], gs_w[s * C + c], acc);
    s_gs[c] = acc;
  }
  __syncthreads();

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const int per_lane = (C + 31) / 32;
  const size_t row0 = static_cast<size_t>(b) * kRowsPerPos;
  __nv_bfloat16* act_bf = reinterpret_cast<__nv_bfloat16*>(act_out);
  float* act_f = reinterpret_cast<float*>(act_out);
  __half* raw_h = reinterpret_cast<__half*>(raw_out);  // bf16 mode: the residual stream is fp16
  float* raw_f = reinterpret_cast<float*>(raw_out);

  for (int q = warp; q < kRowsPerPos; q += nwarps) {
    const size_t row = row0 + q;
    if (!row_is_live(q)) {  // zero halo rows
      for (int k = 0; k < per_lane; ++k) {
        const int c = lane + 32 * k;
        if (c < C) {
          if (kBf16) raw_h[row * C + c] = __float2half_rn(0.0f);
          else raw_f[row * C + c] = 0.0f;
          if (kBf16) act_bf[row * C + c] = __float2bfloat16(0.0f);
          else act_f[row * C + c] = 0.0f;
        }
      }
      continue;
    }
    const int p = row_point(q), r = p / 19, cc = p % 19;
    float acc[kMaxCPerLane];
#pragma unroll
    for (int k = 0; k < kMaxCPerLane; ++k) acc[k] = 0.0f;
    for (int di = 0; di < 5; ++di) {
      const int rr = r + di - 2;
      if (rr < 0 || rr >= 19) continue;
      for (int dj = 0; dj < 5; ++dj) {
        const int c2 = cc + dj - 2;
        if (c2 < 0 || c2 >= 19) continue;
        uint32_t m = s_mask[rr * 19 + c2];
        const float* wtap = wt + static_cast<size_t>(di * 5 + dj) * nplanes * C;
        while (m) {
          const int ch = __ffs(m) - 1;
          m &= m - 1;
          const float* wr = wtap + static_cast<size_t>(ch) * C;
#pragma unroll
          for (int k = 0; k < kMaxCPerLane; ++k) {
            const int c = lane + 32 * k;
            if (k < per_lane && c < C) acc[k] += __ldg(wr + c);
          }
        }
      }
    }
#pragma unroll
    for (int k = 0; k < kMaxCPerLane; ++k) {
      const int c = lane + 32 * k;
      if (k < per_lane && c < C) {
        const float x = acc[k] + s_gs[c];
        if (kBf16) raw_h[row * C + c] = __float2half_rn(x);
        else raw_f[row * C + c] = x;
        const float a = mish_f32<!kBf16>(fmaf(x, scale[c], shift[c]));
        if (kBf16) act_bf[row * C + c] = __float2bfloat16(a);
        else act_f[row * C + c] = a;
      }
    }
  }
}

// ---- bf16 mode, C <= 256: persistent CTAs with the whole [25][P][C] weight table resident in shared memory -----------
// The gather-add of v1 streamed ~15 weight rows of C floats per point from L2 (5.7 TB/s, L2-bound: 0.96 ms at B=1024).
// Rounding the table to bf16 (the operand precision of every other layer in this mode) makes it fit in smem
// (375 * C * 2 B = 187.5 KB at C = 256), so each CTA loads it once and then only reads masks and writes outputs.
constexpr int kIcThreads = 512;

__global__ void __launch_bounds__(kIcThreads, 1)
init_conv_smem_kernel(const uint16_t* __restrict__ masks, const float* __restrict__ scalars, int n, int nplanes,
                      int nscalars, int C, const __nv_bfloat16* __restrict__ wt_bf16, const float* __restrict__ gs_w,
                      const float* __restrict__ gs_b, __half* __restrict__ raw_out, __nv_bfloat16* __restrict__ act_out,
                      const float* __restrict__ scale, const float* __restrict__ shift) {
  extern __shared__ __align__(16) uint8_t ic_smem[];
  const int rows_w = 25 * nplanes;
  __nv_bfloat162* s_w = reinterpret_cast<__nv_bfloat162*>(ic_smem);                        // [rows_w][C/2]
  float* s_gs = reinterpret_cast<float*>(ic_smem + static_cast<size_t>(rows_w) * C * 2);  // [C]
  float* s_sc = s_gs + C;
  float* s_sh = s_sc + C;
  uint16_t* s_mask = reinterpret_cast<uint16_t*>(s_sh + C);                                // [361]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, nwarps = kIcThreads / 32;
  {
    const uint4* src = reinterpret_cast<const uint4*>(wt_bf16);
    uint4* dst = reinterpret_cast<uint4*>(ic_smem);
    const int n16 = rows_w * C * 2 / 16;
    for (int i = tid; i < n16; i += kIcThreads) dst[i] = src[i];
    for (int c = tid; c < C; c += kIcThreads) {
      s_sc[c] = scale[c];
      s_sh[c] = shift[c];
    }
  }
  const int half_c = C / 2;          // bf16x2 pairs per weight row
  const int pairs = half_c / 32;     // pairs per lane (C = 64 * pairs); <= 4
  for (int b = blockIdx.x; b < n; b += gridDim.x) {
    __syncthreads();  // previous position done with s_mask / s_gs (also covers the table load on the first pass)
    for (int p = tid; p < P3_NUM_BOARD_LOCS; p += kIcThreads) s_mask[p] = masks[static_cast<size_t>(b) * P3_NUM_BOARD_LOCS + p];
    for (int c = tid; c < C; c += kIcThreads) {
      float acc = gs_b[c];
      for (int s = 0; s < nscalars; ++s) acc = fmaf(scalars[b * nscalars + s], gs_w[s * C + c], acc);
      s_gs[c] = acc;
    }
    __syncthreads();
    const size_t row0 = static_cast<size_t>(b) * kRowsPerPos;
    for (int q = warp; q < kRowsPerPos; q += nwarps) {
      const size_t row = row0 + q;
      __half2* raw2 = reinterpret_cast<__half2*>(raw_out + row * C);
      __nv_bfloat162* act2 = reinterpret_cast<__nv_bfloat162*>(act_out + row * C);
      if (!row_is_live(q)) {
        for (int k = 0; k < pairs; ++k) {
          raw2[lane + 32 * k] = __floats2half2_rn(0.0f, 0.0f);
          act2[lane + 32 * k] = __floats2bfloat162_rn(0.0f, 0.0f);
        }
        continue;
      }
      const int p = row_point(q), r = p / 19, cc = p % 19;
      float2 acc[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) acc[k] = make_float2(0.0f, 0.0f);
      for (int di = 0; di < 5; ++di) {
        const int rr = r + di - 2;
        if (rr < 0 || rr >= 19) continue;
#pragma unroll
        for (int dj = 0; dj < 5; ++dj) {
          const int c2 = cc + dj - 2;
          if (c2 < 0 || c2 >= 19) continue;
          uint32_t m = s_mask[rr * 19 + c2];
          const __nv_bfloat162* wtap = s_w + static_cast<size_t>((di * 5 + dj) * nplanes) * half_c;
          while (m) {
            const int ch = __ffs(m) - 1;
            m &= m - 1;
            const __nv_bfloat162* wr = wtap + ch * half_c;
#pragma unroll
            for (int k = 0; k < 4; ++k)
              if (k < pairs) {
                const float2 w2 = __bfloat1622float2(wr[lane + 32 * k]);
                acc[k].x += w2.x;
                acc[k].y += w2.y;
              }
          }
        }
      }
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (k < pairs) {
          const int c = 2 * (lane + 32 * k);
          const float x0 = acc[k].x + s_gs[c], x1 = acc[k].y + s_gs[c + 1];
          raw2[lane + 32 * k] = __floats2half2_rn(x0, x1);
          act2[lane + 32 * k] = __floats2bfloat162_rn(mish_f32<false>(fmaf(x0, s_sc[c], s_sh[c])),
                                                      mish_f32<false>(fmaf(x1, s_sc[c + 1], s_sh[c + 1])));
        }
    }
  }
}

}  // namespace

size_t init_conv_smem_bytes(int nplanes, int C) {
  return static_cast<size_t>(25) * nplanes * C * 2 + 3 * C * sizeof(float) + 368 * sizeof(uint16_t) + 16;
}
bool init_conv_smem_supported(int nplanes, int C) {
  return C % 64 == 0 && C <= 256 && init_conv_smem_bytes(nplanes, C) <= 224 * 1024;
}

int init_conv_smem_launch(const uint16_t* masks, const float* scalars, int n, int nplanes, int nscalars, int C,
                          const __nv_bfloat16* wt_bf16, const float* gs_w, const float* gs_b, __half* raw_out,
                          __nv_bfloat16* act_out, const float* scale, const float* shift, cudaStream_t stream) {
  const size_t smem = init_conv_smem_bytes(nplanes, C);
  cudaError_t e = cudaFuncSetAttribute(init_conv_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024);
  if (e != cudaSuccess) return fail(P3_ERR_CUDA, std::string("init_conv smem attribute: ") + cudaGetErrorString(e));
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  init_conv_smem_kernel<<<std::min(sms, n), kIcThreads, smem, stream>>>(masks, scalars, n, nplanes, nscalars, C, wt_bf16, gs_w,
                                                                        gs_b, raw_out, act_out, scale, shift);
  P3_CUDA(cudaGetLastError());
  return P3_OK;
}

int init_conv_launch(const uint16_t* masks, const float* scalars, int n, int nplanes, int nscalars, int C,
                     const float* wt, const float* gs_w, const float* gs_b, void* raw_out, void* act_out,
                     bool act_bf16, const float* scale, const float* shift, cudaStream_t stream) {
  if (C > 32 * kMaxCPerLane) return fail(P3_ERR_UNSUPPORTED, "init_conv: C > 384");
  const size_t smem = sizeof(float) * C;
  if (act_bf16)
    init_conv_kernel<true><<<n, 256, smem, stream>>>(masks, scalars, nplanes, nscalars, C, wt, gs_w, gs_b, raw_out,
                                                     act_out, scale, shift);
  else
    init_conv_kernel<false><<<n, 256, smem, stream>>>(masks, scalars, nplanes, nscalars, C, wt, gs_w, gs_b, raw_out,
                                                      act_out, scale, shift);
  P3_CUDA(cudaGetLastError());
  return P3_OK;
}

}  // namespace p3
