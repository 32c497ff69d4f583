// First layer of the tower on the tensor cores (bf16 engine; python/model.py:1230-1237):
//     x = conv5x5(planes, 15 -> C, same, no bias) + dense(game_state, 8 -> C)      (broadcast over HW)
// The planes are exactly {0, 1} (cc/nn/engine/go_features.cc:10-68), so they are exact in bf16 and the layer is an
// implicit GEMM  [rows, 25 taps x 16 channels] x [400, C]  whose A operand never exists in HBM: every CTA expands the
// 15-bit plane masks the encode kernel produced (722 B / position) into the K-major A tile in shared memory, one
// 32-byte (16-channel) piece per (row, tap), through a 256-entry byte -> 8 x bf16 table.
//
// Operand layout: K-major WITHOUT swizzle, i.e. 8-row x 16-byte core matrices; element (row r, k) lives at
//     (r / 8) * SBO + (k / 8) * LBO + (r % 8) * 16 + (k % 8) * 2     with LBO = 128 B, SBO = 50 * 128 B
// so one tap (K = 16) is two adjacent core matrices and the 25 taps are 25 tcgen05.mma steps with the start address
// advanced by 256 B.  The weights of the CTA's N slice are repacked into the same layout on the host and stay resident.
//
// Persistent CTAs (one per SM), each owning one 64-channel N slice (weights 50 KB + A tile 100 KB + 32 KB of output
// staging fit in shared memory; the A tile is rebuilt per slice, which is cheap next to the stores):
//   warp 0        stages the plane masks and the game-state bias of the next tile
//   warp 1        MMA issuer, owns TMEM (two accumulator stages)
//   warps 2-9     builders: expand the masks of tile it + 1 into the A tile as soon as the MMAs of tile it have read it
//   warps 10-17   epilogue of tile it, concurrently with the builders and the tensor core
//                 (+ game-state bias, fp16 residual stream and bf16 mish(BN_0(x)) copy; a TMEM lane quarter's 4 warps fill
//                 one 32-row x 64-column box per output, which leaves as a TMA bulk store — per-thread 32-byte stores
//                 ran at 1.6 TB/s) while the tensor core works on tile it + 1.
#include <algorithm>
#include <cstdlib>
#include <vector>

#include "common.cuh"
#include "math.cuh"
#include "ptx.cuh"
#include "tc_util.cuh"

namespace p3 {

namespace {

constexpr int kItTaps = 25;
constexpr int kItK = kItTaps * 16;         // 400
constexpr int kItKc = kItK / 8;            // 50 core matrices along K
constexpr int kItSbo = kItKc * 128;        // 6400 B between 8-row groups
constexpr int kItABytes = 16 * kItSbo;     // 128 rows: 102 400 B
constexpr int kItBuildWarps = 8, kItEpiWarps = 8;
constexpr int kItMaskStages = 6;           // plane-mask grids in flight (the bulk copies are latency-, not bandwidth-bound)
constexpr int kItBuilders = kItBuildWarps * 32;  // 256
constexpr int kItThreads = 64 + (kItBuildWarps + kItEpiWarps) * 32;  // 576
constexpr int kItPadW = kMaskPadW, kItPadH = kMaskPadH;  // zero-bordered mask grid: (r + 2) * 24 + (c + 2)
constexpr int kItNw = 64;                  // output channels per CTA (N slice)
constexpr int kItBoxBytes = 32 * 128;      // 32 rows x 64 two-byte elements, 128B swizzle

// K-major, no swizzle: [0,14) start >> 4, [16,30) LBO >> 4, [32,46) SBO >> 4, [46,48) version 1, layout type 0
__device__ __forceinline__ uint32_t desc_lo_nosw(uint32_t smem_addr) { return ((smem_addr >> 4) & 0x3FFFu) | ((128u >> 4) << 16); }
__host__ __device__ constexpr uint32_t desc_hi_nosw() { return (static_cast<uint32_t>(kItSbo) >> 4) | (1u << 14); }

__global__ void __launch_bounds__(kItThreads, 1)
init_tc_kernel(const __grid_constant__ CUtensorMap map_raw, const __grid_constant__ CUtensorMap map_act,
               const uint16_t* __restrict__ masks_padded, const float* __restrict__ gs, int n, int C, int n_w,
               const __nv_bfloat16* __restrict__ w_packed, const float* __restrict__ scale, const float* __restrict__ shift,
               int debug, int f16) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int w_bytes = (n_w / 8) * kItSbo;
  uint8_t* smem_w = smem;
  uint8_t* smem_a = smem_w + w_bytes;
  uint8_t* smem_stage = smem_a + kItABytes;                                        // [4 quarters][raw, act][32 rows x 128 B]
  uint4* s_lut = reinterpret_cast<uint4*>(smem_stage + 4 * 2 * kItBoxBytes);       // [256] byte -> 8 bf16 {0, 1}
  uint16_t* s_mask = reinterpret_cast<uint16_t*>(s_lut + 256);                     // [kItMaskStages tiles][2 positions][23 * 24] zero-bordered
  float* s_gs = reinterpret_cast<float*>(s_mask + kItMaskStages * 2 * kItPadH * kItPadW);  // [kItMaskStages][2 positions][n_w] game-state bias
  float* s_sc = s_gs + kItMaskStages * 2 * kItNw;                                          // [n_w] (x log2 e)
  float* s_sh = s_sc + 128;
  uint64_t* a_full = reinterpret_cast<uint64_t*>(s_sh + 128);
  uint64_t* a_empty = a_full + 1;
  uint64_t* acc_full = a_empty + 1;   // [2]
  uint64_t* acc_empty = acc_full + 2; // [2]
  uint64_t* stage_ready = acc_empty + 2;              // [kItMaskStages] masks of a tile staged by warp 0
  uint64_t* stage_free = stage_ready + kItMaskStages; // [kItMaskStages] the builders have read the masks
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(stage_free + kItMaskStages);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n_slices = C / n_w;
  const int slice = blockIdx.x % n_slices;
  const int n0 = slice * n_w;
  const int cta_in_slice = blockIdx.x / n_slices, ctas_per_slice = gridDim.x / n_slices;
  const int rows = n * kRowsPerPos;
  const int m_tiles = (rows + 127) / 128;
  const int n_it = cta_in_slice < m_tiles ? (m_tiles - cta_in_slice + ctas_per_slice - 1) / ctas_per_slice : 0;

  // ---- setup: resident weights of this slice (already in core-matrix order), LUT, zeroed mask borders, BN constants
  {
    const uint4* src = reinterpret_cast<const uint4*>(w_packed) + static_cast<size_t>(slice) * (w_bytes / 16);
    uint4* dst = reinterpret_cast<uint4*>(smem_w);
    for (int i = tid; i < w_bytes / 16; i += kItThreads) dst[i] = src[i];
    for (int i = tid; i < 256; i += kItThreads) {
      uint32_t w[4];
      const uint32_t one = f16 ? 0x3C00u : 0x3F80u;  // 1.0 in the operand format
#pragma unroll
      for (int j = 0; j < 4; ++j) w[j] = (((i >> (2 * j)) & 1) ? one : 0u) | (((i >> (2 * j + 1)) & 1) ? (one << 16) : 0u);
      s_lut[i] = make_uint4(w[0], w[1], w[2], w[3]);
    }
    constexpr float kLog2e = 1.4426950408889634f;
    for (int c = tid; c < n_w; c += kItThreads) {
      s_sc[c] = scale[n0 + c] * kLog2e;
      s_sh[c] = shift[n0 + c] * kLog2e;
    }
  }
  if (warp == 0 && lane == 0) {
    ptx::mbar_init(a_full, 1);
    ptx::mbar_init(a_empty, 1);
    for (int s = 0; s < 2; ++s) {
      ptx::mbar_init(&acc_full[s], 1);
      ptx::mbar_init(&acc_empty[s], kItEpiWarps);
    }
    for (int s = 0; s < kItMaskStages; ++s) {
      ptx::mbar_init(&stage_ready[s], 1);
      ptx::mbar_init(&stage_free[s], 1 + kItEpiWarps);  // the builders (masks) and every epilogue warp (bias) are done with it
    }
    ptx::fence_mbar_init();
  } else if (warp == 1) {
    ptx::tmem_alloc(tmem_ptr, static_cast<uint32_t>(2 * n_w));
    ptx::tmem_relinquish();
  }
  ptx::fence_proxy_async();  // the weight tile was written through the generic proxy
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ===== stager: bulk-copies the zero-bordered plane-mask grids of the (at most two) positions a tile touches, one tile
    // ahead of the workers =====
    for (int it = 0; it < n_it; ++it) {
      const int sb = it % kItMaskStages;
      ptx::mbar_wait(&stage_free[sb], ((static_cast<uint32_t>(it / kItMaskStages)) & 1u) ^ 1u);
      const int m0 = (cta_in_slice + it * ctas_per_slice) * 128;
      const int b0 = m0 / kRowsPerPos;
      const int npos = (b0 + 1 < n) ? 2 : 1;
      if (lane == 0) {
        ptx::mbar_arrive_expect_tx(&stage_ready[sb], static_cast<uint32_t>(npos * (kItPadH * kItPadW * 2 + n_w * 4)));
        ptx::bulk_load_1d(s_mask + sb * 2 * kItPadH * kItPadW, masks_padded + static_cast<size_t>(b0) * kItPadH * kItPadW,
                          static_cast<uint32_t>(npos * kItPadH * kItPadW * 2), &stage_ready[sb]);
        for (int pb = 0; pb < npos; ++pb)  // this slice's game-state bias of the tile's position(s)
          ptx::bulk_load_1d(s_gs + (sb * 2 + pb) * kItNw, gs + static_cast<size_t>(b0 + pb) * C + n0, static_cast<uint32_t>(n_w * 4),
                            &stage_ready[sb]);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    const uint32_t idesc = ptx::make_idesc_op(128, n_w, f16);
    const uint32_t a_lo = desc_lo_nosw(ptx::smem_u32(smem_a)), w_lo = desc_lo_nosw(ptx::smem_u32(smem_w));
    for (int it = 0; it < n_it; ++it) {
      const int as = it & 1;
      ptx::mbar_wait(&acc_empty[as], ((static_cast<uint32_t>(it) >> 1) & 1u) ^ 1u);
      ptx::mbar_wait(a_full, static_cast<uint32_t>(it) & 1u);
      ptx::tc_fence_after_sync();
      if (ptx::elect_one()) {
#pragma unroll 5
        for (int t = 0; t < ((debug & 1) ? 0 : kItTaps); ++t)  // one tap = K 16 = two core matrices = 256 B further along K
          ptx::umma_f16_lohi(tmem_base + static_cast<uint32_t>(as * n_w), a_lo + 16u * t, desc_hi_nosw(), w_lo + 16u * t,
                             desc_hi_nosw(), idesc, t > 0 ? 1u : 0u);
        ptx::umma_commit(a_empty);
        ptx::umma_commit(&acc_full[as]);
      }
      __syncwarp();
    }
  } else if (warp < 2 + kItBuildWarps) {
    // ===== builders =====
    const int bt = tid - 64;               // 0..255
    const int brow = bt & 127;             // row of the tile this thread builds (fixed), taps bt / 128 + 2 k
    const uint32_t a_row = ptx::smem_u32(smem_a) + static_cast<uint32_t>(brow >> 3) * kItSbo + static_cast<uint32_t>(brow & 7) * 16u;
    const uint32_t lut = ptx::smem_u32(s_lut);
    for (int it = 0; it < n_it; ++it) {
      const int m0 = (cta_in_slice + it * ctas_per_slice) * 128;
      const int b0 = m0 / kRowsPerPos;
      const int sb = it % kItMaskStages;
      ptx::mbar_wait(&stage_ready[sb], static_cast<uint32_t>(it / kItMaskStages) & 1u);
      ptx::mbar_wait(a_empty, (static_cast<uint32_t>(it) & 1u) ^ 1u);  // the MMAs of the previous tile have read A
      const int m = m0 + brow;
      const int pb = m / kRowsPerPos - b0, qq = m % kRowsPerPos;
      const bool live = m < rows && row_is_live(qq);
      const int r = (qq - kRowBase) / kRowPitch, c = (qq - kRowBase) % kRowPitch;
      const uint16_t* mg = s_mask + (sb * 2 + pb) * kItPadH * kItPadW + r * kItPadW + c;  // (r + dy + 2, c + dx + 2) = mg[(dy+2)*24 + dx+2]
      constexpr int kPer = (kItTaps + 1) / 2;  // 13
      uint32_t mk[kPer];
#pragma unroll
      for (int k = 0; k < kPer; ++k) {  // this thread's taps: bt / 128 + 2 k
        const int t = (bt >> 7) + 2 * k;
        mk[k] = (live && t < kItTaps) ? mg[(t / 5) * kItPadW + (t % 5)] : 0u;
      }
#pragma unroll
      for (int k = 0; k < kPer; ++k) {
        const int t = (bt >> 7) + 2 * k;
        if (t < kItTaps && !(debug & 2)) {
          const float4 lo = ptx::lds_f4_const(lut + (mk[k] & 0xffu) * 16u), hi = ptx::lds_f4_const(lut + (mk[k] >> 8) * 16u);
          ptx::sts_f4(a_row + static_cast<uint32_t>(t) * 256u, lo);
          ptx::sts_f4(a_row + static_cast<uint32_t>(t) * 256u + 128u, hi);
        }
      }
      ptx::fence_proxy_async();
      ptx::named_bar_sync(1, kItBuilders);
      if (bt == 0) {
        ptx::mbar_arrive(a_full);
        ptx::mbar_arrive(&stage_free[sb]);
      }
    }
  } else {
    // ===== epilogue: 2 warps per TMEM lane quarter, thread = one row x 32 of the slice's 64 columns =====
    const int ew = warp - 2 - kItBuildWarps;  // 0..7
    const int q = warp & 3;                   // TMEM lane quarter
    const int cg = ew >> 2;                   // which half of the slice's columns
    const uint32_t sc = ptx::smem_u32(s_sc), sh = ptx::smem_u32(s_sh);
    const bool qleader = cg == 0 && lane == 0;
    const uint32_t sw = static_cast<uint32_t>(lane & 7);
    // this thread's 4 chunks of the 128-byte box row: columns cg * 32 + h * 16 + {0..7, 8..15}
    const uint32_t box_raw = ptx::smem_u32(smem_stage) + static_cast<uint32_t>(q) * (2 * kItBoxBytes);
    const uint32_t box_act = box_raw + kItBoxBytes;
    auto epilogue = [&](int it) {
      const int as = it & 1;
      const int m0 = (cta_in_slice + it * ctas_per_slice) * 128;
      const int m = m0 + q * 32 + lane;
      const bool live = m < rows && row_is_live(m % kRowsPerPos);
      // game-state bias of this row's position: staged in shared memory with the tile's masks (the builders waited for it)
      const int sb = it % kItMaskStages;
      const uint32_t gp = ptx::smem_u32(s_gs + (sb * 2 + (m < rows ? m / kRowsPerPos - m0 / kRowsPerPos : 0)) * kItNw + cg * 32);
      ptx::mbar_wait(&acc_full[as], (static_cast<uint32_t>(it) >> 1) & 1u);
      ptx::tc_fence_after_sync();
      const uint32_t ro = static_cast<uint32_t>(lane) * 128u;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
      const int col = cg * 32 + h * 16;
      uint32_t v[16];
      ptx::tmem_ld_32x16(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(as * n_w + col), v);
      ptx::tmem_ld_wait();
      float x[16], a[16];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float4 g4 = ptx::lds_f4(gp + static_cast<uint32_t>(4 * h + i) * 16u);
        x[4 * i] = __uint_as_float(v[4 * i]) + g4.x;
        x[4 * i + 1] = __uint_as_float(v[4 * i + 1]) + g4.y;
        x[4 * i + 2] = __uint_as_float(v[4 * i + 2]) + g4.z;
        x[4 * i + 3] = __uint_as_float(v[4 * i + 3]) + g4.w;
      }
      if (h == 1) {  // the accumulator and the bias are in registers: hand both stages back
        ptx::tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) {
          ptx::mbar_arrive(&acc_empty[as]);
          ptx::mbar_arrive(&stage_free[sb]);
        }
      }
#pragma unroll
      for (int g = 0; g < 2; ++g) {
        const float4 s0 = ptx::lds_f4_const(sc + static_cast<uint32_t>(col + 8 * g) * 4u), s1 = ptx::lds_f4_const(sc + static_cast<uint32_t>(col + 8 * g + 4) * 4u);
        const float4 h0 = ptx::lds_f4_const(sh + static_cast<uint32_t>(col + 8 * g) * 4u), h1 = ptx::lds_f4_const(sh + static_cast<uint32_t>(col + 8 * g + 4) * 4u);
        const float scv[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
        const float shv[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
        float z[8], d[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) z[i] = fmaf(x[8 * g + i], scv[i], shv[i]);
#pragma unroll
        for (int i = 0; i < 8; ++i) d[i] = ex2_approx_ftz(fminf(z[i], 28.853900817779268f));
#pragma unroll
        for (int i = 0; i < 8; ++i) d[i] = fmaf(d[i], d[i] + 2.0f, 2.0f);
#pragma unroll
        for (int i = 0; i < 4; ++i) {  // mish(t) = z (ln2 - 2 ln2 / d), pairs share one reciprocal (see chain_tc.cu)
          const float r = rcp_approx_ftz(d[2 * i] * d[2 * i + 1]);
          a[8 * g + 2 * i] = z[2 * i] * fmaf(d[2 * i + 1] * r, -1.3862943611198906f, 0.6931471805599453f);
          a[8 * g + 2 * i + 1] = z[2 * i + 1] * fmaf(d[2 * i] * r, -1.3862943611198906f, 0.6931471805599453f);
        }
      }
      uint4 r0 = make_uint4(tc_pack_f16(x[0], x[1]), tc_pack_f16(x[2], x[3]), tc_pack_f16(x[4], x[5]), tc_pack_f16(x[6], x[7]));
      uint4 r1 = make_uint4(tc_pack_f16(x[8], x[9]), tc_pack_f16(x[10], x[11]), tc_pack_f16(x[12], x[13]), tc_pack_f16(x[14], x[15]));
      uint4 p0 = make_uint4(tc_pack_act(a[0], a[1], f16), tc_pack_act(a[2], a[3], f16), tc_pack_act(a[4], a[5], f16), tc_pack_act(a[6], a[7], f16));
      uint4 p1 = make_uint4(tc_pack_act(a[8], a[9], f16), tc_pack_act(a[10], a[11], f16), tc_pack_act(a[12], a[13], f16), tc_pack_act(a[14], a[15], f16));
      if (!live) r0 = r1 = p0 = p1 = make_uint4(0, 0, 0, 0);  // padding rows of the layout are zeros
      if (h == 0) {
        // the quarter's previous stores have read the boxes: waited for by the quarter leader AFTER the first pass's math,
        // which so overlaps the TMA engine's reads
        if (qleader) ptx::bulk_wait_read<0>();
        ptx::named_bar_sync(2 + q, 64);
      }
      if (!(debug & 4)) {
        const uint32_t c0 = ((4u * cg + 2u * h) ^ sw) << 4, c1 = ((4u * cg + 2u * h + 1u) ^ sw) << 4;
        ptx::sts_u4(box_raw + ro + c0, r0);
        ptx::sts_u4(box_raw + ro + c1, r1);
        ptx::sts_u4(box_act + ro + c0, p0);
        ptx::sts_u4(box_act + ro + c1, p1);
      }
      }
      ptx::fence_proxy_async();
      ptx::named_bar_sync(2 + q, 64);
      if (qleader) {
        ptx::tma_store_2d(&map_raw, nullptr, 0, 0, box_raw, n0, m0 + q * 32);
        ptx::tma_store_2d(&map_act, nullptr, 0, 0, box_act, n0, m0 + q * 32);
        ptx::bulk_commit();
      }
    };

    for (int it = 0; it < n_it; ++it) epilogue(it);
    if (qleader) ptx::bulk_wait_all();
  }

  ptx::tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after_sync();
    ptx::tmem_dealloc(tmem_base, static_cast<uint32_t>(2 * n_w));
  }
}

int init_tc_slice_width(int) { return kItNw; }
size_t init_tc_smem_bytes(int n_w) {
  return static_cast<size_t>(n_w / 8) * kItSbo + kItABytes + 4 * 2 * kItBoxBytes + 256 * 16 + kItMaskStages * 2 * (kItPadH * kItPadW * 2 + kItNw * 4) + 2 * 128 * 4 + 256 + 1024;
}

}  // namespace

bool init_tc_supported(int nplanes, int nscalars, int C) {
  (void)nscalars;
  return nplanes <= 16 && C % 64 == 0 && C >= 64;
}

// [25][nplanes][C] fp32 tap-major table -> per N slice, K-major core-matrix order [n / 8][k / 8][n % 8][k % 8], k = tap * 16 + plane
int init_tc_pack_weights(const float* wt, int nplanes, int C, std::vector<__nv_bfloat16>& out, bool op_f16) {
  const int n_w = init_tc_slice_width(C);
  out.assign(static_cast<size_t>(C) * kItK, __float2bfloat16(0.0f));
  auto to_op = [&](float v) {  // the 16-bit operand word, carried in the bf16-typed vector
    if (!op_f16) return __float2bfloat16(v);
    const __half h = __float2half_rn(v);
    return *reinterpret_cast<const __nv_bfloat16*>(&h);
  };
  for (int c = 0; c < C; ++c) {
    const int slice = c / n_w, nl = c % n_w;
    for (int t = 0; t < kItTaps; ++t)
      for (int p = 0; p < nplanes; ++p) {
        const int k = t * 16 + p;
        const size_t idx = static_cast<size_t>(slice) * n_w * kItK + (static_cast<size_t>(nl / 8) * kItKc + k / 8) * 64 + (nl % 8) * 8 + (k % 8);
        out[idx] = to_op(wt[(static_cast<size_t>(t) * nplanes + p) * C + c]);
      }
  }
  return P3_OK;
}

struct InitTcPlan {
  CUtensorMap map_raw, map_act;  // 32-row x 64-column output boxes of the [rows, C] fp16 / bf16 matrices
  int n = 0, C = 0, grid = 0, debug = 0, f16 = 0;
  size_t smem = 0;
  const uint16_t* masks_padded = nullptr;
  const float *gs = nullptr, *scale = nullptr, *shift = nullptr;
  const __nv_bfloat16* w_packed = nullptr;
  __half* raw_out = nullptr;
  __nv_bfloat16* act_out = nullptr;
};

int init_tc_plan_create(const uint16_t* masks_padded, const float* gs, int n, int C, const __nv_bfloat16* w_packed,
                        __half* raw_out, __nv_bfloat16* act_out, const float* scale, const float* shift, InitTcPlan** out,
                        bool op_f16) {
  InitTcPlan* p = new InitTcPlan();
  p->f16 = op_f16 ? 1 : 0;
  p->masks_padded = masks_padded; p->gs = gs; p->n = n; p->C = C; p->w_packed = w_packed;
  p->raw_out = raw_out; p->act_out = act_out; p->scale = scale; p->shift = shift;
  const int n_w = init_tc_slice_width(C);
  p->smem = init_tc_smem_bytes(n_w);
  int rc = tc_make_map_2d(&p->map_raw, raw_out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, C, static_cast<uint64_t>(n) * kRowsPerPos, 64, 32,
                          CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc == P3_OK)
    rc = tc_make_map_2d(&p->map_act, act_out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, C, static_cast<uint64_t>(n) * kRowsPerPos, 64, 32,
                        CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc == P3_OK) {
    cudaError_t e = cudaFuncSetAttribute(init_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) rc = fail(P3_ERR_CUDA, std::string("init_tc smem attribute: ") + cudaGetErrorString(e));
  }
  if (rc != P3_OK) {
    delete p;
    return rc;
  }
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int n_slices = C / n_w;
  const int m_tiles = (n * kRowsPerPos + 127) / 128;
  p->grid = std::max(n_slices, std::min(sms, m_tiles * n_slices) / n_slices * n_slices);
  if (const char* d = std::getenv("P3_INIT_TC_DEBUG")) p->debug = std::atoi(d);  // perf ablations (results are wrong when set)
  *out = p;
  return P3_OK;
}

void init_tc_plan_destroy(InitTcPlan* p) { delete p; }

int init_tc_launch(const InitTcPlan* p, cudaStream_t stream) {
  init_tc_kernel<<<p->grid, kItThreads, p->smem, stream>>>(p->map_raw, p->map_act, p->masks_padded, p->gs, p->n, p->C,
                                                           init_tc_slice_width(p->C), p->w_packed, p->scale, p->shift, p->debug, p->f16);
  P3_CUDA(cudaGetLastError());
  return P3_OK;
}

}  // namespace p3
