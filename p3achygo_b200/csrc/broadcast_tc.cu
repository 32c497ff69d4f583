// BroadcastPreAct (python/model.py:570-581) on tcgen05: per position b the board-mixing Dense(361 -> 361) is
//     Y_b[q, c] = sum_p W[p, q] * X_b[p, c] + bias[q]
// a GEMM whose A operand is the shared matrix W^T (K-major, padded to the board-row layout: [512 q-rows, 448 p])
// and whose B operand is the NHWC activation itself, consumed MN-major (channels contiguous) straight from the
// padded layout — no transpose of the activation is ever materialised.
//
//   M = 128 padded output rows q (4 tiles cover the 400 rows of a position; rows >= 400 are clipped by the 3-D
//       TMA store), N = up to 256 channels, K = 448 padded input rows p in 7 slabs of 64.
//   Halo rows/cols of W^T are zero, halo rows of X are zero, and rows p >= 400 are zero-filled by the 3-D TMA
//   load (per-position bounds), so the padding contributes nothing.
// Epilogue: + bias[q], then the following conv's BN + mish (python/model.py:276-281), bf16, zero halo rows,
// written through a 64B-swizzled staging tile and cp.async.bulk.tensor stores.
// Same warp-specialised structure as conv_tc.cu (TMA producer / MMA issuer / 8 epilogue warps).
#include <cuda.h>

#include <algorithm>
#include <vector>

#include "common.cuh"
#include "math.cuh"
#include "ptx.cuh"

namespace p3 {

namespace {
constexpr int kBM = 128;
constexpr int kBK = 64;
constexpr int kBKTotal = 448;    // 400 padded rows -> 7 slabs of 64
constexpr int kBMTotal = 512;    // 4 M tiles
constexpr int kBThreads = 64 + 256;
constexpr int kBSmemBudget = 224 * 1024;
constexpr int kBActStage = kBM * 32 * 2;  // 8 KB
constexpr int kBNumOut = 2;               // double-buffered output staging
constexpr int kBXBoxBytes = kBK * 128;    // [64 p-rows][64 channels] bf16 = 8 KB
}  // namespace

struct TcBcastPlan {
  CUtensorMap map_wt, map_x, map_act;
  int B, C, n_tile, stages, tmem_cols, grid;
  size_t smem_bytes;
  const float *bias_pad, *scale, *shift;
  void* wt_dev = nullptr;
  void* bias_dev = nullptr;
};

namespace {

__device__ __forceinline__ uint32_t pack2(float a, float b) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}

// smem descriptor of an MN-major (channel-contiguous) operand: 64-channel x 8-row atoms of 1024 B (128B swizzle);
// LBO = byte stride between 64-channel blocks, SBO = byte stride between 8-row K groups.
__device__ __forceinline__ uint64_t make_desc_mn_sw128(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr >> 4) & 0x3FFF);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

__device__ __forceinline__ void tma_load_3d(void* smem_dst, const void* desc, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(ptx::smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(desc)), "r"(ptx::smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_store_3d(const void* desc, const void* smem_src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(desc)), "r"(ptx::smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}

__global__ void __launch_bounds__(kBThreads, 1)
tc_broadcast_kernel(const __grid_constant__ CUtensorMap map_wt, const __grid_constant__ CUtensorMap map_x,
                    const __grid_constant__ CUtensorMap map_act, int B, int C, int n_tile, int stages, int tmem_cols,
                    const float* __restrict__ bias_pad, const float* __restrict__ scale,
                    const float* __restrict__ shift) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* st_act = smem;  // 8 KB
  uint8_t* ring = smem + kBNumOut * kBActStage;
  const int a_bytes = kBM * kBK * 2;  // 16 KB
  const int n_boxes = n_tile / 64;
  const int stage_bytes = a_bytes + n_boxes * kBXBoxBytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(ring + static_cast<size_t>(stages) * stage_bytes);
  uint64_t* empty_bar = full_bar + stages;
  uint64_t* tmem_full = empty_bar + stages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tiles = C / n_tile;
  const int m_tiles = kBMTotal / kBM;  // 4
  const int total_tiles = B * n_tiles * m_tiles;
  const int k_steps = kBKTotal / kBK;  // 7

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&map_wt);
    ptx::prefetch_tensormap(&map_x);
    ptx::prefetch_tensormap(&map_act);
    for (int s = 0; s < stages; ++s) {
      ptx::mbar_init(&full_bar[s], 1);
      ptx::mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      ptx::mbar_init(&tmem_full[s], 1);
      ptx::mbar_init(&tmem_empty[s], 8);
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

  // tile -> (position b, channel tile nt, row tile mt); mt fastest so neighbouring CTAs share the X_b slab in L2
  auto decode = [&](int tile, int& b, int& nt, int& mt) {
    mt = tile % m_tiles;
    const int rest = tile / m_tiles;
    nt = rest % n_tiles;
    b = rest / n_tiles;
  };

  if (warp == 0) {
    {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        int b, nt, mt;
        decode(tile, b, nt, mt);
        for (int ks = 0; ks < k_steps; ++ks) {
          ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
          if (ptx::elect_one()) {
            uint8_t* sa = ring + static_cast<size_t>(stage) * stage_bytes;
            ptx::mbar_arrive_expect_tx(&full_bar[stage], static_cast<uint32_t>(stage_bytes));
            ptx::tma_load_2d(sa, &map_wt, &full_bar[stage], ks * kBK, mt * kBM);
            for (int j = 0; j < n_boxes; ++j)
              tma_load_3d(sa + a_bytes + j * kBXBoxBytes, &map_x, &full_bar[stage], nt * n_tile + j * 64, ks * kBK, b);
          }
          __syncwarp();
          if (++stage == stages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    {
      // A K-major, B MN-major (bit 16)
      const uint32_t idesc = ptx::make_idesc_bf16(kBM, n_tile) | (1u << 16);
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
          const uint32_t sa = ptx::smem_u32(ring + static_cast<size_t>(stage) * stage_bytes);
          const uint32_t a_lo = ptx::desc_lo_sw128(sa);
          const uint64_t db0 = make_desc_mn_sw128(sa + a_bytes, kBXBoxBytes, 1024);
          const uint32_t b_lo = static_cast<uint32_t>(db0), b_hi = static_cast<uint32_t>(db0 >> 32);
          if (ptx::elect_one()) {
#pragma unroll
            for (int k = 0; k < kBK / 16; ++k)
              // A: +32 B along K; B: 16 K rows = two 8-row groups = 2048 B (128 x 16 B) further down the [rows][128 B] box
              ptx::umma_f16_lohi(tmem_d, a_lo + 2 * k, ptx::desc_hi_sw128(), b_lo + 128 * k, b_hi, idesc,
                                 (step > 0 || k > 0) ? 1u : 0u);
            ptx::umma_commit(&empty_bar[stage]);
          }
          __syncwarp();
          if (++stage == stages) {
            stage = 0;
            phase ^= 1;
          }
        }
        if (ptx::elect_one()) ptx::umma_commit(&tmem_full[acc]);
        __syncwarp();
      }
    }
  } else {
    const int ew = warp - 2;
    const int quarter = warp & 3;
    const int half = (ew >> 2) & 1;
    const int r = quarter * 32 + lane;
    const bool leader = (threadIdx.x == 64);
    const uint32_t bf_row = static_cast<uint32_t>(r) * 64u;
    const int n_chunks = n_tile / 32;
    uint32_t g = 0;
    int iter = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++iter) {
      int b, nt, mt;
      decode(tile, b, nt, mt);
      const int acc = iter & 1;
      const uint32_t acc_phase = (iter >> 1) & 1;
      ptx::mbar_wait(&tmem_full[acc], acc_phase);
      ptx::tc_fence_after_sync();
      const int q = mt * kBM + r;
      const bool live = q < kRowsPerPos && row_is_live(q);
      const float bq = bias_pad[q];
      for (int c = 0; c < n_chunks; ++c, ++g) {
        uint8_t* st_buf = st_act + (g % kBNumOut) * kBActStage;
        uint32_t v[16];
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) +
                               static_cast<uint32_t>(acc * n_tile + c * 32 + half * 16);
        ptx::tmem_ld_32x16(taddr, v);
        const int nb = nt * n_tile + c * 32 + half * 16;
        float sc[16], sh[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          sc[j] = __ldg(scale + nb + j);
          sh[j] = __ldg(shift + nb + j);
        }
        ptx::tmem_ld_wait();
        float a[16];
#pragma unroll
        for (int j = 0; j < 16; ++j)
          a[j] = live ? mish_f32<false>(fmaf(__uint_as_float(v[j]) + bq, sc[j], sh[j])) : 0.0f;
        if (leader) ptx::bulk_wait_read<kBNumOut - 1>();
        ptx::named_bar_sync(1, 256);
        uint8_t* wp = st_buf + bf_row;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int o = 8 * j;
          const uint4 pk = make_uint4(pack2(a[o], a[o + 1]), pack2(a[o + 2], a[o + 3]), pack2(a[o + 4], a[o + 5]),
                                      pack2(a[o + 6], a[o + 7]));
          *reinterpret_cast<uint4*>(wp + (((half * 2 + j) ^ ((r >> 1) & 3)) << 4)) = pk;
        }
        ptx::fence_proxy_async();
        ptx::named_bar_sync(2, 256);
        if (leader) {
          tma_store_3d(&map_act, st_buf, nt * n_tile + c * 32, mt * kBM, b);  // rows >= 400 are clipped
          ptx::bulk_commit();
        }
      }
      ptx::tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&tmem_empty[acc]);
    }
    if (leader) ptx::bulk_wait_all();
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
EncodeTiledFn encode_fn() {
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

int pick_bcast_n(int C) {
  for (int n = 256; n >= 64; n -= 64)
    if (C % n == 0) return n;
  return 0;
}

}  // namespace

bool tc_broadcast_supported(int C) { return C % 64 == 0 && pick_bcast_n(C) > 0; }

// w [361][361] fp32 (Keras Dense kernel: [in p][out q]), bias [361]: HOST pointers. x / act_out: device bf16 [B*400, C].
int tc_broadcast_plan_create(const float* w_host, const float* bias_host, const void* x, void* act_out, int B, int C,
                             const float* scale, const float* shift, TcBcastPlan** out) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return fail(P3_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
  if (!tc_broadcast_supported(C)) return fail(P3_ERR_UNSUPPORTED, "tc_broadcast: C % 64 != 0");
  TcBcastPlan* p = new TcBcastPlan();
  p->B = B;
  p->C = C;
  p->scale = scale;
  p->shift = shift;
  p->n_tile = pick_bcast_n(C);
  // W^T in the padded board-row space: wt[q_pad][p_pad] = w[p][q]; halo rows / cols and padding are zero
  std::vector<__nv_bfloat16> wt(static_cast<size_t>(kBMTotal) * kBKTotal, __float2bfloat16(0.0f));
  std::vector<float> bias_pad(kBMTotal, 0.0f);
  for (int q = 0; q < 361; ++q) {
    const int qp = board_row(q);
    bias_pad[qp] = bias_host[q];
    for (int pt = 0; pt < 361; ++pt) wt[static_cast<size_t>(qp) * kBKTotal + board_row(pt)] = __float2bfloat16(w_host[pt * 361 + q]);
  }
  if (cudaMalloc(&p->wt_dev, wt.size() * 2) != cudaSuccess || cudaMalloc(&p->bias_dev, bias_pad.size() * 4) != cudaSuccess) {
    delete p;
    return fail(P3_ERR_CUDA, "tc_broadcast: cudaMalloc failed");
  }
  cudaMemcpy(p->wt_dev, wt.data(), wt.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(p->bias_dev, bias_pad.data(), bias_pad.size() * 4, cudaMemcpyHostToDevice);
  p->bias_pad = reinterpret_cast<const float*>(p->bias_dev);

  const int a_bytes = kBM * kBK * 2;
  const int stage_bytes = a_bytes + (p->n_tile / 64) * kBXBoxBytes;
  p->stages = std::min(6, (kBSmemBudget - 1024 - 512 - kBNumOut * kBActStage) / stage_bytes);
  p->smem_bytes = static_cast<size_t>(p->stages) * stage_bytes + 1024 + 512 + kBNumOut * kBActStage;
  int cols = 32;
  while (cols < 2 * p->n_tile) cols *= 2;
  p->tmem_cols = cols;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  p->grid = std::min(sms, B * (C / p->n_tile) * (kBMTotal / kBM));

  CUresult r;
  {
    cuuint64_t gdim[2] = {static_cast<cuuint64_t>(kBKTotal), static_cast<cuuint64_t>(kBMTotal)};
    cuuint64_t gstride[1] = {static_cast<cuuint64_t>(kBKTotal) * 2};
    cuuint32_t box[2] = {kBK, kBM};
    cuuint32_t es[2] = {1, 1};
    r = fn(&p->map_wt, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, p->wt_dev, gdim, gstride, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
           CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  }
  if (r == CUDA_SUCCESS) {  // activation as [B][400][C]: per-position bounds -> rows >= 400 load as zeros
    cuuint64_t gdim[3] = {static_cast<cuuint64_t>(C), static_cast<cuuint64_t>(kRowsPerPos), static_cast<cuuint64_t>(B)};
    cuuint64_t gstride[2] = {static_cast<cuuint64_t>(C) * 2, static_cast<cuuint64_t>(C) * 2 * kRowsPerPos};
    cuuint32_t box[3] = {64, kBK, 1};
    cuuint32_t es[3] = {1, 1, 1};
    r = fn(&p->map_x, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(x), gdim, gstride, box, es,
           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  }
  if (r == CUDA_SUCCESS) {  // output, same 3-D view: stores of rows >= 400 are clipped
    cuuint64_t gdim[3] = {static_cast<cuuint64_t>(C), static_cast<cuuint64_t>(kRowsPerPos), static_cast<cuuint64_t>(B)};
    cuuint64_t gstride[2] = {static_cast<cuuint64_t>(C) * 2, static_cast<cuuint64_t>(C) * 2 * kRowsPerPos};
    cuuint32_t box[3] = {32, kBM, 1};
    cuuint32_t es[3] = {1, 1, 1};
    r = fn(&p->map_act, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, act_out, gdim, gstride, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
           CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  }
  cudaError_t e = cudaSuccess;
  if (r == CUDA_SUCCESS)
    e = cudaFuncSetAttribute(tc_broadcast_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kBSmemBudget);
  if (r != CUDA_SUCCESS || e != cudaSuccess) {
    cudaFree(p->wt_dev);
    cudaFree(p->bias_dev);
    delete p;
    return fail(P3_ERR_CUDA, "tc_broadcast: tensor map / attribute setup failed");
  }
  *out = p;
  return P3_OK;
}

void tc_broadcast_plan_destroy(TcBcastPlan* p) {
  if (!p) return;
  cudaFree(p->wt_dev);
  cudaFree(p->bias_dev);
  delete p;
}

int tc_broadcast_launch(const TcBcastPlan* p, cudaStream_t stream) {
  tc_broadcast_kernel<<<p->grid, kBThreads, p->smem_bytes, stream>>>(p->map_wt, p->map_x, p->map_act, p->B, p->C, p->n_tile,
                                                                     p->stages, p->tmem_cols, p->bias_pad, p->scale, p->shift);
  P3_CUDA(cudaGetLastError());
  return P3_OK;
}

}  // namespace p3
