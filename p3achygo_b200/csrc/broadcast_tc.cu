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
// Epilogue: + bias[q], then the following conv's BN + mish (python/model.py:276-281), bf16, zero halo rows.  16 epilogue
// warps (4 per TMEM lane quarter); a quarter's 4 warps fill one 32-row x 64-channel box per step, which the quarter's
// I/O warp stores with cp.async.bulk.tensor (same box-pool scheme as chain_tc.cu: the epilogue warps never issue TMA,
// never wait for a bulk group and never meet at a CTA barrier).  M tiles cover rows 20 .. 403 of a position (3 tiles
// instead of 4 over 0 .. 511); the leading 20 halo rows are written as zeros by quarter 0 of the first tile.
#include <cuda.h>

#include <algorithm>
#include <vector>

#include "common.cuh"
#include "math.cuh"
#include "ptx.cuh"
#include "tc_util.cuh"

namespace p3 {

namespace {
constexpr int kBM = 128;
constexpr int kBK = 64;
constexpr int kBKTotal = 448;    // 400 padded rows -> 7 slabs of 64
constexpr int kBMTotal = 512;    // 4 M tiles
constexpr int kBMTiles = 3;      // output rows 20 .. 403 of a position (rows 0..19 are halo, rows >= 400 are clipped)
constexpr int kBRow0 = 20;
constexpr int kBEpiWarps = 16;
constexpr int kBThreads = (2 + kBEpiWarps + 4) * 32;  // producer, MMA, 16 epilogue, 4 I/O (one per TMEM lane quarter)
constexpr int kBSmemBudget = 224 * 1024;
constexpr int kBBoxBytes = 32 * 128;      // a quarter's 32 rows x 64 bf16 output box (128B swizzle)
constexpr int kBBoxes = 3;                // boxes per quarter
constexpr int kBXBoxBytes = kBK * 128;    // [64 p-rows][64 channels] bf16 = 8 KB
constexpr int kBMaxC = 512;
constexpr int kBFixedSmem = (4 * kBBoxes + 1) * kBBoxBytes + 1024 /*align*/ + 512 /*barriers*/ + 2 * kBMaxC * 4;
}  // namespace

struct TcBcastPlan {
  CUtensorMap map_wt, map_x, map_act, map_halo;
  int B, C, n_tile, stages, tmem_cols, grid, f16 = 0, reverse = 0;
  size_t smem_bytes;
  const float *bias_pad, *scale, *shift;
  void* wt_dev = nullptr;
  void* bias_dev = nullptr;
};

namespace {

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
__device__ __forceinline__ void tma_store_3d(const void* desc, uint32_t smem_src_addr, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(desc)), "r"(smem_src_addr), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}

__global__ void __launch_bounds__(kBThreads, 1)
tc_broadcast_kernel(const __grid_constant__ CUtensorMap map_wt, const __grid_constant__ CUtensorMap map_x,
                    const __grid_constant__ CUtensorMap map_act, const __grid_constant__ CUtensorMap map_halo, int B, int C,
                    int n_tile, int stages, int tmem_cols,
                    const float* __restrict__ bias_pad, const float* __restrict__ scale,
                    const float* __restrict__ shift, int f16, int reverse) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_box = smem;                                  // [4 quarters][kBBoxes] output boxes, then one zero box
  uint8_t* zero_box = smem_box + 4 * kBBoxes * kBBoxBytes;
  uint8_t* ring = zero_box + kBBoxBytes;
  const int a_bytes = kBM * kBK * 2;  // 16 KB
  const int n_boxes = n_tile / 64;
  const int stage_bytes = a_bytes + n_boxes * kBXBoxBytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(ring + static_cast<size_t>(stages) * stage_bytes);
  uint64_t* empty_bar = full_bar + stages;
  uint64_t* tmem_full = empty_bar + stages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint64_t* box_ready = tmem_empty + 2;             // [4][kBBoxes] the box's previous store has read it
  uint64_t* box_written = box_ready + 4 * kBBoxes;  // [4][kBBoxes] the quarter's 4 warps have filled it
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(box_written + 4 * kBBoxes);
  float* s_scale = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(full_bar) + 512);  // x log2(e), see bn_mish8
  float* s_shift = s_scale + kBMaxC;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tiles = C / n_tile;
  const int total_tiles = B * n_tiles * kBMTiles;
  const int k_steps = kBKTotal / kBK;  // 7

  constexpr float kLog2e = 1.4426950408889634f;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    s_scale[c] = scale[c] * kLog2e;
    s_shift[c] = shift[c] * kLog2e;
  }
  for (int i = threadIdx.x; i < kBBoxBytes / 16; i += blockDim.x) reinterpret_cast<uint4*>(zero_box)[i] = make_uint4(0, 0, 0, 0);
  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&map_wt);
    ptx::prefetch_tensormap(&map_x);
    ptx::prefetch_tensormap(&map_act);
    ptx::prefetch_tensormap(&map_halo);
    for (int s = 0; s < stages; ++s) {
      ptx::mbar_init(&full_bar[s], 1);
      ptx::mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      ptx::mbar_init(&tmem_full[s], 1);
      ptx::mbar_init(&tmem_empty[s], kBEpiWarps);
    }
    for (int s = 0; s < 4 * kBBoxes; ++s) {
      ptx::mbar_init(&box_ready[s], 1);
      ptx::mbar_init(&box_written[s], 4);
    }
    ptx::fence_mbar_init();
  } else if (warp == 1) {
    ptx::tmem_alloc(tmem_ptr, static_cast<uint32_t>(tmem_cols));
    ptx::tmem_relinquish();
  }
  ptx::fence_proxy_async();  // zero box: generic-proxy writes -> TMA
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_ptr;

  // tile -> (position b, channel tile nt, row tile mt); mt fastest so neighbouring CTAs share the X_b slab in L2
  auto decode = [&](int tile, int& b, int& nt, int& mt) {
    mt = tile % kBMTiles;
    const int rest = tile / kBMTiles;
    nt = rest % n_tiles;
    b = rest / n_tiles;
    if (reverse) b = B - 1 - b;  // from the last position to the first (tile order of a launch: common.cuh, pair_tile_row0)
  };

  if (warp == 0) {
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
          ptx::tma_load_2d(sa, &map_wt, &full_bar[stage], ks * kBK, kBRow0 + mt * kBM);
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
  } else if (warp == 1) {
    // A K-major, B MN-major (bit 16)
    const uint32_t idesc = ptx::make_idesc_op(kBM, n_tile, f16) | (1u << 16);
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
  } else if (warp >= 2 + kBEpiWarps) {
    // ===== I/O warp of quarter q: stores the quarter's boxes as the epilogue warps fill them =====
    const int q = warp - (2 + kBEpiWarps);
    uint64_t* my_ready = box_ready + kBBoxes * q;
    uint64_t* my_written = box_written + kBBoxes * q;
    const uint32_t my_box = ptx::smem_u32(smem_box) + static_cast<uint32_t>(q) * (kBBoxes * kBBoxBytes);
    uint32_t s = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      int b, nt, mt;
      decode(tile, b, nt, mt);
      for (int j = 0; j < n_boxes; ++j, ++s) {
        const uint32_t ob = s % kBBoxes;
        ptx::mbar_wait(&my_written[ob], (s / kBBoxes) & 1u);
        if (lane == 0) {
          tma_store_3d(&map_act, my_box + ob * kBBoxBytes, nt * n_tile + j * 64, kBRow0 + mt * kBM + q * 32, b);
          // the 20 leading halo rows of the position: a zero box of 20 rows at row 0
          if (mt == 0 && q == 0) tma_store_3d(&map_halo, ptx::smem_u32(zero_box), nt * n_tile + j * 64, 0, b);
          ptx::bulk_commit();
          ptx::bulk_wait_read<1>();  // the previous step's store has read its box
          if (s > 0) ptx::mbar_arrive(&my_ready[(s - 1) % kBBoxes]);
        }
        __syncwarp();
      }
    }
    if (lane == 0) ptx::bulk_wait_all();
  } else {
    // ===== epilogue: 4 warps per TMEM lane quarter, thread = one output row x 16 of a step's 64 channels =====
    const int ew = warp - 2;
    const int q = warp & 3;
    const int cg = ew >> 2;
    const uint32_t sw = static_cast<uint32_t>(lane & 7);
    const uint32_t ch0 = ((2u * cg) ^ sw) << 4, ch1 = ((2u * cg + 1u) ^ sw) << 4;
    const uint32_t box_base = ptx::smem_u32(smem_box) + static_cast<uint32_t>(q) * (kBBoxes * kBBoxBytes) + static_cast<uint32_t>(lane) * 128u;
    uint64_t* my_ready = box_ready + kBBoxes * q;
    uint64_t* my_written = box_written + kBBoxes * q;
    const uint32_t sc = ptx::smem_u32(s_scale), sh = ptx::smem_u32(s_shift);
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    uint32_t s = 0;
    int iter = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++iter) {
      int b, nt, mt;
      decode(tile, b, nt, mt);
      const int acc = iter & 1;
      const int qrow = kBRow0 + mt * kBM + q * 32 + lane;
      const bool live = qrow < kRowsPerPos && row_is_live(qrow);
      const float bq = bias_pad[qrow];
      ptx::mbar_wait(&tmem_full[acc], (static_cast<uint32_t>(iter) >> 1) & 1u);
      ptx::tc_fence_after_sync();
      for (int j = 0; j < n_boxes; ++j, ++s) {
        const int col = nt * n_tile + j * 64 + cg * 16;
        uint32_t v[16];
        ptx::tmem_ld_32x16(lane_addr + static_cast<uint32_t>(acc * n_tile + j * 64 + cg * 16), v);
        ptx::tmem_ld_wait();
        if (j == n_boxes - 1) {  // the accumulator is in registers: hand it back to the MMA warp
          ptx::tc_fence_before_sync();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(&tmem_empty[acc]);
        }
        float x[16], a[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) x[i] = __uint_as_float(v[i]) + bq;
        bn_mish8(x, a, sc, sh, col);
        bn_mish8(x + 8, a + 8, sc, sh, col + 8);
        uint4 p0 = make_uint4(tc_pack_act(a[0], a[1], f16), tc_pack_act(a[2], a[3], f16), tc_pack_act(a[4], a[5], f16), tc_pack_act(a[6], a[7], f16));
        uint4 p1 = make_uint4(tc_pack_act(a[8], a[9], f16), tc_pack_act(a[10], a[11], f16), tc_pack_act(a[12], a[13], f16), tc_pack_act(a[14], a[15], f16));
        if (!live) p0 = p1 = make_uint4(0, 0, 0, 0);  // halo rows / columns of the layout stay zero
        const uint32_t ob = s % kBBoxes;
        ptx::mbar_wait(&my_ready[ob], ((s / kBBoxes) & 1u) ^ 1u);
        const uint32_t obuf = box_base + ob * kBBoxBytes;
        ptx::sts_u4(obuf + ch0, p0);
        ptx::sts_u4(obuf + ch1, p1);
        ptx::fence_proxy_async();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&my_written[ob]);
      }
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
                             const float* scale, const float* shift, TcBcastPlan** out, bool op_f16) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return fail(P3_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
  if (!tc_broadcast_supported(C)) return fail(P3_ERR_UNSUPPORTED, "tc_broadcast: C % 64 != 0");
  TcBcastPlan* p = new TcBcastPlan();
  p->f16 = op_f16 ? 1 : 0;
  p->B = B;
  p->C = C;
  p->scale = scale;
  p->shift = shift;
  p->n_tile = pick_bcast_n(C);
  // W^T in the padded board-row space: wt[q_pad][p_pad] = w[p][q]; halo rows / cols and padding are zero
  std::vector<uint16_t> wt(static_cast<size_t>(kBMTotal) * kBKTotal, 0);  // 16-bit operand words (bf16 or fp16; zero is zero in both)
  auto to_op = [&](float v) -> uint16_t {
    if (op_f16) { const __half h = __float2half_rn(v); return *reinterpret_cast<const uint16_t*>(&h); }
    const __nv_bfloat16 h = __float2bfloat16(v); return *reinterpret_cast<const uint16_t*>(&h);
  };
  std::vector<float> bias_pad(kBMTotal, 0.0f);
  for (int q = 0; q < 361; ++q) {
    const int qp = board_row(q);
    bias_pad[qp] = bias_host[q];
    for (int pt = 0; pt < 361; ++pt) wt[static_cast<size_t>(qp) * kBKTotal + board_row(pt)] = to_op(w_host[pt * 361 + q]);
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
  p->stages = std::min(6, (kBSmemBudget - kBFixedSmem) / stage_bytes);
  p->smem_bytes = static_cast<size_t>(p->stages) * stage_bytes + kBFixedSmem;
  int cols = 32;
  while (cols < 2 * p->n_tile) cols *= 2;
  p->tmem_cols = cols;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  p->grid = std::min(sms, B * (C / p->n_tile) * kBMTiles);

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
    cuuint32_t box[3] = {64, 32, 1};
    cuuint32_t es[3] = {1, 1, 1};
    r = fn(&p->map_act, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, act_out, gdim, gstride, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
           CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  }
  if (r == CUDA_SUCCESS) {  // the 20 leading halo rows of a position (zero box)
    cuuint64_t gdim[3] = {static_cast<cuuint64_t>(C), static_cast<cuuint64_t>(kRowsPerPos), static_cast<cuuint64_t>(B)};
    cuuint64_t gstride[2] = {static_cast<cuuint64_t>(C) * 2, static_cast<cuuint64_t>(C) * 2 * kRowsPerPos};
    cuuint32_t box[3] = {64, kBRow0, 1};
    cuuint32_t es[3] = {1, 1, 1};
    r = fn(&p->map_halo, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, act_out, gdim, gstride, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
           CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
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

void tc_broadcast_plan_set_reverse(TcBcastPlan* p, bool reverse) {
  if (p) p->reverse = reverse ? 1 : 0;
}

void tc_broadcast_plan_destroy(TcBcastPlan* p) {
  if (!p) return;
  cudaFree(p->wt_dev);
  cudaFree(p->bias_dev);
  delete p;
}

int tc_broadcast_launch(const TcBcastPlan* p, cudaStream_t stream) {
  tc_broadcast_kernel<<<p->grid, kBThreads, p->smem_bytes, stream>>>(p->map_wt, p->map_x, p->map_act, p->map_halo, p->B, p->C, p->n_tile,
                                                                     p->stages, p->tmem_cols, p->bias_pad, p->scale, p->shift, p->f16, p->reverse);
  P3_CUDA(cudaGetLastError());
  return P3_OK;
}

}  // namespace p3
