// tcgen05 / TMEM / TMA implicit-GEMM convolution for the residual tower (P3_PRECISION_BF16).
//
// conv(mish(BN(x))) of python/model.py:276-281 over the padded board-row layout (common.cuh): the
// activated bf16 input is a 2-D matrix [rows, cin]; a k x k "same" conv is `taps` row-shifted GEMMs
//     acc[m, n] = sum_t sum_k  A[m + tap_off[t], k] * W[t][n][k]
// whose zero padding is the layout's zero rows plus TMA's out-of-bounds zero fill (negative and
// past-the-end row coordinates).  No im2col matrix is materialised.
//
// Warp-specialised persistent kernel, one CTA per SM:
//   warp 0    TMA producer: per (tap, 64-channel slab) one 128 x 64 bf16 A box and one N x 64 bf16
//             weight box, 128B-swizzled, into a multi-stage shared-memory ring (mbarrier full/empty)
//   warp 1    MMA issuer: one thread issues tcgen05.mma (M=128, N=tile width, K=16) x 4 per slab into
//             one of two fp32 accumulators in TMEM; tcgen05.commit releases the ring slot / publishes
//             the accumulator
//   warps 2-5 epilogue: tcgen05.ld the accumulator (32 lanes x 32 columns per warp and step), add the
//             fp32 residual, store the raw fp32 stream and/or the bf16 activated copy with the NEXT
//             layer's BN + mish folded in (ConvEpilogue), zeros on halo rows.  Runs concurrently with
//             the next tile's MMAs through the second accumulator.
#include <cuda.h>

#include <vector>

#include "common.cuh"
#include "math.cuh"
#include "ptx.cuh"

namespace p3 {

constexpr int kTileM = 128;
constexpr int kSlabK = 64;                 // bf16 elements per 128-byte swizzled row
constexpr int kUmmaK = 16;
constexpr int kMaxTapsTc = 9;
constexpr int kABytes = kTileM * kSlabK * 2;  // 16 KB
constexpr int kNumThreads = 192;
constexpr int kSmemBudget = 220 * 1024;

struct TcTaps {
  int off[kMaxTapsTc];
};

struct TcConvPlan {
  CUtensorMap map_a;
  CUtensorMap map_w;
  int rows, cin, cout, taps, n_tile, stages, tmem_cols, grid;
  size_t smem_bytes;
  TcTaps tap;
};

namespace {

__global__ void __launch_bounds__(kNumThreads, 1)
tc_conv_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w, int rows,
               int cin, int cout, int taps, TcTaps tap, int n_tile, int stages, int tmem_cols,
               const float* residual, float* raw_out, __nv_bfloat16* act_out, const float* __restrict__ scale,
               const float* __restrict__ shift, int act_mode) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // carve: [stages][A 16 KB | B n_tile*128 B] | barriers | tmem ptr
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int b_bytes = n_tile * kSlabK * 2;
  const int stage_bytes = kABytes + b_bytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + static_cast<size_t>(stages) * stage_bytes);
  uint64_t* empty_bar = full_bar + stages;
  uint64_t* tmem_full = empty_bar + stages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m_tiles = (rows + kTileM - 1) / kTileM;
  const int n_tiles = cout / n_tile;
  const int total_tiles = m_tiles * n_tiles;
  const int k_slabs = cin / kSlabK;
  const int k_steps = taps * k_slabs;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&map_a);
    ptx::prefetch_tensormap(&map_w);
    for (int s = 0; s < stages; ++s) {
      ptx::mbar_init(&full_bar[s], 1);
      ptx::mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      ptx::mbar_init(&tmem_full[s], 1);
      ptx::mbar_init(&tmem_empty[s], 4);  // one arrive per epilogue warp
    }
    ptx::fence_mbar_init();
  } else if (warp == 1) {
    ptx::tmem_alloc(tmem_ptr, static_cast<uint32_t>(tmem_cols));
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int m0 = (tile / n_tiles) * kTileM, n0 = (tile % n_tiles) * n_tile;
        for (int t = 0; t < taps; ++t) {
          for (int ks = 0; ks < k_slabs; ++ks) {
            ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
            uint8_t* sa = smem + static_cast<size_t>(stage) * stage_bytes;
            ptx::mbar_arrive_expect_tx(&full_bar[stage], static_cast<uint32_t>(stage_bytes));
            ptx::tma_load_2d(sa, &map_a, &full_bar[stage], ks * kSlabK, m0 + tap.off[t]);
            ptx::tma_load_2d(sa + kABytes, &map_w, &full_bar[stage], ks * kSlabK, t * cout + n0);
            if (++stage == stages) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      const uint32_t idesc = ptx::make_idesc_bf16(kTileM, n_tile);
      int stage = 0;
      uint32_t phase = 0;
      int iter = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++iter) {
        const int acc = iter & 1;
        const uint32_t acc_phase = (iter >> 1) & 1;
        ptx::mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        ptx::tc_fence_after_sync();
        const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(acc * n_tile);
        for (int step = 0; step < k_steps; ++step) {
          ptx::mbar_wait(&full_bar[stage], phase);
          ptx::tc_fence_after_sync();
          const uint32_t sa = ptx::smem_u32(smem + static_cast<size_t>(stage) * stage_bytes);
          const uint32_t sb = sa + kABytes;
#pragma unroll
          for (int k = 0; k < kSlabK / kUmmaK; ++k) {
            // advance 16 bf16 = 32 bytes along K inside the 128-byte swizzled row
            const uint64_t da = ptx::make_desc_sw128(sa + k * kUmmaK * 2);
            const uint64_t db = ptx::make_desc_sw128(sb + k * kUmmaK * 2);
            ptx::umma_f16(tmem_d, da, db, idesc, (step > 0 || k > 0) ? 1u : 0u);
          }
          ptx::umma_commit(&empty_bar[stage]);  // ring slot reusable once these MMAs have read it
          if (++stage == stages) {
            stage = 0;
            phase ^= 1;
          }
        }
        ptx::umma_commit(&tmem_full[acc]);  // accumulator complete
      }
    }
  } else {
    // ===== epilogue warps 2..5: TMEM lane quarter = warp % 4 =====
    const int quarter = warp & 3;
    int iter = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++iter) {
      const int m0 = (tile / n_tiles) * kTileM, n0 = (tile % n_tiles) * n_tile;
      const int acc = iter & 1;
      const uint32_t acc_phase = (iter >> 1) & 1;
      ptx::mbar_wait(&tmem_full[acc], acc_phase);
      ptx::tc_fence_after_sync();
      const int m = m0 + quarter * 32 + lane;
      const bool in_range = m < rows;
      const bool live = in_range && row_is_live(m % kRowsPerPos);
      const size_t row_off = static_cast<size_t>(m) * cout + n0;
      for (int c0 = 0; c0 < n_tile; c0 += 32) {
        uint32_t v[32];
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) +
                               static_cast<uint32_t>(acc * n_tile + c0);
        ptx::tmem_ld_32x32(taddr, v);
        float r[32];
        if (residual != nullptr && live) {
          const float4* rp = reinterpret_cast<const float4*>(residual + row_off + c0);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 t4 = rp[i];
            r[4 * i] = t4.x; r[4 * i + 1] = t4.y; r[4 * i + 2] = t4.z; r[4 * i + 3] = t4.w;
          }
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) r[i] = 0.0f;
        }
        ptx::tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) r[i] = live ? (__uint_as_float(v[i]) + r[i]) : 0.0f;
        if (in_range) {
          if (raw_out != nullptr) {
            float4* op = reinterpret_cast<float4*>(raw_out + row_off + c0);
#pragma unroll
            for (int i = 0; i < 8; ++i) op[i] = make_float4(r[4 * i], r[4 * i + 1], r[4 * i + 2], r[4 * i + 3]);
          }
          if (act_out != nullptr) {
            uint32_t packed[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              float a0 = r[2 * i], a1 = r[2 * i + 1];
              if (act_mode == kActMishBN) {
                const int n = n0 + c0 + 2 * i;
                a0 = mish_f32<false>(fmaf(a0, __ldg(scale + n), __ldg(shift + n)));
                a1 = mish_f32<false>(fmaf(a1, __ldg(scale + n + 1), __ldg(shift + n + 1)));
              } else if (act_mode == kActMish) {
                a0 = mish_f32<false>(a0);
                a1 = mish_f32<false>(a1);
              }
              if (!live) a0 = a1 = 0.0f;
              const __nv_bfloat162 h = __floats2bfloat162_rn(a0, a1);
              packed[i] = *reinterpret_cast<const uint32_t*>(&h);
            }
            uint4* ap = reinterpret_cast<uint4*>(act_out + row_off + c0);
#pragma unroll
            for (int i = 0; i < 4; ++i)
              ap[i] = make_uint4(packed[4 * i], packed[4 * i + 1], packed[4 * i + 2], packed[4 * i + 3]);
          }
        }
      }
      // accumulator drained -> hand it back to the MMA warp
      ptx::tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&tmem_empty[acc]);
    }
  }

  ptx::tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after_sync();
    ptx::tmem_dealloc(tmem_base, static_cast<uint32_t>(tmem_cols));
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
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

// 2-D bf16 row-major [dim1, dim0] tensor, box [box1, 64], 128-byte swizzle, zero OOB fill.
int make_map_2d(CUtensorMap* map, const void* base, uint64_t dim0, uint64_t dim1, uint32_t box1) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return fail(P3_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
  cuuint64_t gdim[2] = {dim0, dim1};
  cuuint64_t gstride[1] = {dim0 * 2};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(kSlabK), box1};
  cuuint32_t estride[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estride,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(P3_ERR_CUDA, "cuTensorMapEncodeTiled failed: " + std::to_string(r));
  return P3_OK;
}

int pick_n_tile(int cout) {
  if (cout <= 256) return cout;
  if (cout % 2 == 0 && cout / 2 <= 256) return cout / 2;
  if (cout % 3 == 0 && cout / 3 <= 256) return cout / 3;
  return 0;
}

}  // namespace

bool tc_conv_supported(int cin, int cout) {
  if (cin <= 0 || cin % kSlabK != 0) return false;
  const int n = pick_n_tile(cout);
  return n >= 32 && n % 32 == 0;
}

int tc_conv_plan_create(const __nv_bfloat16* in, const __nv_bfloat16* w, int rows, int cin, int cout, int taps,
                        const int* tap_off_host, TcConvPlan** out) {
  if (!tc_conv_supported(cin, cout)) return fail(P3_ERR_UNSUPPORTED, "tc_conv: cin % 64 != 0 or cout not tileable");
  if (taps > kMaxTapsTc) return fail(P3_ERR_INVALID_ARG, "tc_conv: too many taps");
  TcConvPlan* p = new TcConvPlan();
  p->rows = rows; p->cin = cin; p->cout = cout; p->taps = taps;
  p->n_tile = pick_n_tile(cout);
  for (int t = 0; t < taps; ++t) p->tap.off[t] = tap_off_host[t];
  const int stage_bytes = kABytes + p->n_tile * kSlabK * 2;
  p->stages = std::min(8, (kSmemBudget - 1024 - 256) / stage_bytes);
  p->smem_bytes = static_cast<size_t>(p->stages) * stage_bytes + 1024 /*align slack*/ + 256 /*barriers*/;
  int cols = 32;
  while (cols < 2 * p->n_tile) cols *= 2;
  p->tmem_cols = cols;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int total_tiles = ((rows + kTileM - 1) / kTileM) * (cout / p->n_tile);
  p->grid = std::min(total_tiles, sms);
  int rc = make_map_2d(&p->map_a, in, cin, rows, kTileM);
  if (rc == P3_OK) rc = make_map_2d(&p->map_w, w, cin, static_cast<uint64_t>(taps) * cout, p->n_tile);
  if (rc != P3_OK) {
    delete p;
    return rc;
  }
  {  // per-device attribute; cheap, so set it for every plan
    cudaError_t e = cudaFuncSetAttribute(tc_conv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget);
    if (e != cudaSuccess) {
      delete p;
      return fail(P3_ERR_CUDA, std::string("cudaFuncSetAttribute(smem): ") + cudaGetErrorString(e));
    }
  }
  *out = p;
  return P3_OK;
}

void tc_conv_plan_destroy(TcConvPlan* plan) { delete plan; }

int tc_conv_launch(const TcConvPlan* p, const ConvEpilogue& ep, cudaStream_t stream) {
  tc_conv_kernel<<<p->grid, kNumThreads, p->smem_bytes, stream>>>(
      p->map_a, p->map_w, p->rows, p->cin, p->cout, p->taps, p->tap, p->n_tile, p->stages, p->tmem_cols, ep.residual,
      ep.raw_out, reinterpret_cast<__nv_bfloat16*>(ep.act_out), ep.scale, ep.shift, ep.act_mode);
  P3_CUDA(cudaGetLastError());
  return P3_OK;
}

}  // namespace p3
