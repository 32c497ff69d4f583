// First layer of the tower on the tensor cores, second form (round 2; python/model.py:1230-1237):
//     x = conv5x5(planes, 15 -> C, same, no bias) + dense(game_state, 8 -> C)      (broadcast over HW)
//
// init_tc.cu builds the im2col A tile of the implicit GEMM explicitly - 25 taps x 16 channels per output row, 100 KB of shared
// memory writes per 128-row tile, and once per 64-channel slice: 1.25 MB per position, which is what bounds it (0.24 ms for a
// layer whose outputs take 0.064 ms to store).  Here the planes of a position are expanded ONCE, one 32-byte (16-channel) record
// per cell of the zero-bordered 23 x 24 mask grid the encode kernel writes, and the 25 taps are 25 row-SHIFTED VIEWS of that one
// array (the trick of the 3x3 kernel): with GEMM rows = grid cells, tap (dy, dx) of cell g is cell g + 24 dy + dx, i.e. the
// operand descriptor's start address moved by 16 bytes per cell.
//
// Operand layout: K-major, no swizzle: element (row r, k) at (r / 8) * SBO + (k / 8) * LBO + (r % 8) * 16 + (k % 8) * 2.
//   A: two planes (channels 0-7, 8-15) of 16-byte cell records, SBO = 128 (cells contiguous), LBO = plane stride;
//      descriptor start = plane 0 + (front pad + g0 + 24 dy + dx) * 16.
//   B: the slice's weights [n_w x 400], SBO = 50 * 128, LBO = 128 (as init_tc.cu), resident; tap t = start + 256 t.
// M tile = 5 grid rows = 120 cells (+ 8 unused lanes), 4 tiles per position (board rows 0-4, 5-9, 10-14, 15-18).
//
// Output: the epilogue writes each grid row of a tile (24 cells x 64 channels) as one 128B-swizzled staging box, and a 4-D TMA
// map over the [B x 400, C] output viewed as (channel, column c < 19, board row r < 19, position) turns the 24-pitch cells into
// the 20-pitch padded board-row layout: one 19-cell store per grid row, from 2 cells into the staging box.  The layout's zero
// column and zero rows are never written (they keep the zeros the buffers start with).  (A 24-cell box at c = -2 that relies on
// clipping at both ends faults as an illegal instruction on B200: negative start coordinates are not for stores.)
//
// Persistent CTAs, one N slice (128 channels, or 64) per CTA:
//   warp 0        stager: bulk-copies a position's mask grid + game-state bias slice
//   warp 1        MMA issuer (owns TMEM: 4 accumulator stages)
//   warps 2-5     builders: masks -> cell records (byte -> 8 x 16-bit table), double-buffered per position
//   warps 6-21    epilogue: + bias, fp16 residual stream and activated copy mish(BN_0(x)), staging, TMA stores
#include <algorithm>
#include <cstdlib>
#include <vector>

#include "common.cuh"
#include "math.cuh"
#include "ptx.cuh"
#include "tc_util.cuh"

namespace p3 {

namespace {

constexpr int kI2Taps = 25;
constexpr int kI2K = kI2Taps * 16;             // 400
constexpr int kI2Kc = kI2K / 8;                // 50 core matrices along K
constexpr int kI2WSbo = kI2Kc * 128;           // 6400 B between 8-row groups of the weights
constexpr int kI2Front = 56;                   // cells in front of the grid (tap shifts reach -50)
constexpr int kI2Cells = 672;                  // cell records per plane: 56 + 552 + slack for +50 shifts of the unused lanes (tile 3 ends at 535 + 8)
constexpr int kI2PlaneBytes = kI2Cells * 16;   // 10 752 B = LBO of the A operand
constexpr int kI2EBytes = 2 * kI2PlaneBytes;   // one position
constexpr int kI2RowBoxBytes = 24 * 128;       // one grid row of a tile x 64 channels (3072 B, a multiple of 1024)
constexpr int kI2TileRows = 5;                 // grid rows per M tile
constexpr int kI2BuildWarps = 4, kI2EpiWarps = 16;  // the epilogue (BN + mish + packs) is the long pole: 4 warps per scheduler
constexpr int kI2Builders = kI2BuildWarps * 32;
constexpr int kI2Threads = 64 + (kI2BuildWarps + kI2EpiWarps) * 32;  // 704
#ifndef I2_MASK_STAGES
#define I2_MASK_STAGES 4
#endif
#ifndef I2_ACC_STAGES
#define I2_ACC_STAGES 4
#endif
#ifndef I2_E_BUFS
#define I2_E_BUFS 2
#endif
constexpr int kI2MaskStages = I2_MASK_STAGES;
constexpr int kI2AccStages = I2_ACC_STAGES;
constexpr int kI2EBufs = I2_E_BUFS;
constexpr int kI2MaskElems = kMaskPadElems;    // 552

__device__ __forceinline__ uint32_t desc_lo(uint32_t smem_addr, uint32_t lbo) { return ((smem_addr >> 4) & 0x3FFFu) | ((lbo >> 4) << 16); }
__host__ __device__ constexpr uint32_t desc_hi(uint32_t sbo) { return (sbo >> 4) | (1u << 14); }

__device__ __forceinline__ void tma_store_4d(const void* desc, uint32_t smem_src, int32_t c0, int32_t c1, int32_t c2, int32_t c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(reinterpret_cast<uint64_t>(desc)),
               "r"(smem_src), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}

__global__ void __launch_bounds__(kI2Threads, 1)
init_tc2_kernel(const __grid_constant__ CUtensorMap map_raw, const __grid_constant__ CUtensorMap map_act,
                const uint16_t* __restrict__ masks_padded, const float* __restrict__ gs, int n, int C, int n_w,
                const __nv_bfloat16* __restrict__ w_packed, const float* __restrict__ scale, const float* __restrict__ shift, int f16,
                int debug) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int n_slabs = n_w / 64;
  const int w_bytes = (n_w / 8) * kI2WSbo;
  const int stage_bytes = 2 * n_slabs * kI2TileRows * kI2RowBoxBytes;                       // [raw, act][slab][grid row]
  uint8_t* smem_stage = smem;                                                               // 1024-aligned boxes first
  uint8_t* smem_w = smem_stage + stage_bytes;
  uint8_t* smem_e = smem_w + w_bytes;                                                       // [2 positions][2 planes][cells]
  uint4* s_lut = reinterpret_cast<uint4*>(smem_e + 2 * kI2EBytes);                          // [256] byte -> 8 x 16-bit {0, 1}
  uint16_t* s_mask = reinterpret_cast<uint16_t*>(s_lut + 256);                              // [stages][552]
  float* s_gs = reinterpret_cast<float*>(s_mask + kI2MaskStages * kI2MaskElems);            // [stages][n_w]
  float* s_sc = s_gs + kI2MaskStages * 128;
  float* s_sh = s_sc + 128;
  uint64_t* e_full = reinterpret_cast<uint64_t*>(s_sh + 128);  // [2]
  uint64_t* e_empty = e_full + 2;                              // [2]
  uint64_t* acc_full = e_empty + 2;                            // [4]
  uint64_t* acc_empty = acc_full + kI2AccStages;               // [4]
  uint64_t* stage_ready = acc_empty + kI2AccStages;            // [kI2MaskStages]
  uint64_t* stage_free = stage_ready + kI2MaskStages;          // [kI2MaskStages]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(stage_free + kI2MaskStages);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n_slices = C / n_w;
  const int slice = blockIdx.x % n_slices;
  const int n0 = slice * n_w;
  const int cta_in_slice = blockIdx.x / n_slices, ctas_per_slice = gridDim.x / n_slices;
  const int n_pos = cta_in_slice < n ? (n - cta_in_slice + ctas_per_slice - 1) / ctas_per_slice : 0;  // positions of this CTA

  {  // resident weights of this slice (already in core-matrix order), LUT, BN constants, zeroed cell arrays
    const uint4* src = reinterpret_cast<const uint4*>(w_packed) + static_cast<size_t>(slice) * (w_bytes / 16);
    uint4* dst = reinterpret_cast<uint4*>(smem_w);
    for (int i = tid; i < w_bytes / 16; i += kI2Threads) dst[i] = src[i];
    const uint32_t one = f16 ? 0x3C00u : 0x3F80u;  // 1.0 in the operand format
    for (int i = tid; i < 256; i += kI2Threads) {
      uint32_t w[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) w[j] = (((i >> (2 * j)) & 1) ? one : 0u) | (((i >> (2 * j + 1)) & 1) ? (one << 16) : 0u);
      s_lut[i] = make_uint4(w[0], w[1], w[2], w[3]);
    }
    uint4* ez = reinterpret_cast<uint4*>(smem_e);
    for (int i = tid; i < 2 * kI2EBytes / 16; i += kI2Threads) ez[i] = make_uint4(0, 0, 0, 0);  // pads stay zero for good
    constexpr float kLog2e = 1.4426950408889634f;
    for (int c = tid; c < n_w; c += kI2Threads) {
      s_sc[c] = scale[n0 + c] * kLog2e;
      s_sh[c] = shift[n0 + c] * kLog2e;
    }
  }
  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&map_raw);
    ptx::prefetch_tensormap(&map_act);
    for (int s = 0; s < 2; ++s) {
      ptx::mbar_init(&e_full[s], 1);
      ptx::mbar_init(&e_empty[s], 1);
    }
    for (int s = 0; s < kI2AccStages; ++s) {
      ptx::mbar_init(&acc_full[s], 1);
      ptx::mbar_init(&acc_empty[s], kI2EpiWarps);
    }
    for (int s = 0; s < kI2MaskStages; ++s) {
      ptx::mbar_init(&stage_ready[s], 1);
      ptx::mbar_init(&stage_free[s], 1 + kI2EpiWarps);  // the builders (masks) and every epilogue warp (bias, after the position's last tile)
    }
    ptx::fence_mbar_init();
  } else if (warp == 1) {
    ptx::tmem_alloc(tmem_ptr, static_cast<uint32_t>(kI2AccStages * n_w));
    ptx::tmem_relinquish();
  }
  ptx::fence_proxy_async();  // weights and zeroed cell arrays were written through the generic proxy
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ===== stager =====
    for (int i = 0; i < n_pos; ++i) {
      const int sb = i % kI2MaskStages;
      ptx::mbar_wait(&stage_free[sb], ((static_cast<uint32_t>(i / kI2MaskStages)) & 1u) ^ 1u);
      const int b = cta_in_slice + i * ctas_per_slice;
      if (lane == 0) {
        ptx::mbar_arrive_expect_tx(&stage_ready[sb], static_cast<uint32_t>(kI2MaskElems * 2 + n_w * 4));
        ptx::bulk_load_1d(s_mask + sb * kI2MaskElems, masks_padded + static_cast<size_t>(b) * kI2MaskElems, kI2MaskElems * 2, &stage_ready[sb]);
        ptx::bulk_load_1d(s_gs + sb * 128, gs + static_cast<size_t>(b) * C + n0, static_cast<uint32_t>(n_w * 4), &stage_ready[sb]);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    const uint32_t idesc = ptx::make_idesc_op(128, n_w, f16);
    const uint32_t w_lo = desc_lo(ptx::smem_u32(smem_w), 128);
    uint32_t tile = 0;
    for (int i = 0; i < n_pos; ++i) {
      const int eb = i % kI2EBufs;
      ptx::mbar_wait(&e_full[eb], static_cast<uint32_t>(i / kI2EBufs) & 1u);
      ptx::tc_fence_after_sync();
      const uint32_t e_lo = desc_lo(ptx::smem_u32(smem_e + eb * kI2EBytes), kI2PlaneBytes);
      for (int t = 0; t < 4; ++t, ++tile) {
        const uint32_t as = tile % kI2AccStages;
        ptx::mbar_wait(&acc_empty[as], ((tile / kI2AccStages) & 1u) ^ 1u);
        ptx::tc_fence_after_sync();
        if (ptx::elect_one()) {
          const uint32_t g0 = static_cast<uint32_t>(kI2Front + (2 + kI2TileRows * t) * kMaskPadW);  // first cell of the tile
#pragma unroll 5
          for (int tap = 0; tap < ((debug & 1) ? 0 : kI2Taps); ++tap) {
            const int shift = (tap / 5 - 2) * kMaskPadW + (tap % 5 - 2);
            ptx::umma_f16_lohi(tmem_base + as * static_cast<uint32_t>(n_w), e_lo + g0 + static_cast<uint32_t>(shift), desc_hi(128),
                               w_lo + 16u * tap, desc_hi(kI2WSbo), idesc, tap > 0 ? 1u : 0u);
          }
          ptx::umma_commit(&acc_full[as]);
          if (t == 3) ptx::umma_commit(&e_empty[eb]);  // the position's cell records have been read
        }
        __syncwarp();
      }
    }
  } else if (warp < 2 + kI2BuildWarps) {
    // ===== builders: mask grid -> cell records =====
    const int bt = tid - 64;
    const uint32_t lut = ptx::smem_u32(s_lut);
    for (int i = 0; i < n_pos; ++i) {
      const int sb = i % kI2MaskStages, eb = i % kI2EBufs;
      ptx::mbar_wait(&stage_ready[sb], static_cast<uint32_t>(i / kI2MaskStages) & 1u);
      ptx::mbar_wait(&e_empty[eb], (static_cast<uint32_t>(i / kI2EBufs) & 1u) ^ 1u);
      const uint16_t* mg = s_mask + sb * kI2MaskElems;
      const uint32_t e0 = ptx::smem_u32(smem_e + eb * kI2EBytes) + kI2Front * 16u;
      for (int g = bt; g < kI2MaskElems; g += kI2Builders) {
        const uint32_t m = mg[g];
        const float4 lo = ptx::lds_f4_const(lut + (m & 0xffu) * 16u), hi = ptx::lds_f4_const(lut + (m >> 8) * 16u);
        ptx::sts_f4(e0 + static_cast<uint32_t>(g) * 16u, lo);
        ptx::sts_f4(e0 + kI2PlaneBytes + static_cast<uint32_t>(g) * 16u, hi);
      }
      ptx::fence_proxy_async();
      ptx::named_bar_sync(1, kI2Builders);
      if (bt == 0) {
        ptx::mbar_arrive(&e_full[eb]);
        ptx::mbar_arrive(&stage_free[sb]);
      }
    }
  } else {
    // ===== epilogue: 4 warps per TMEM lane quarter; thread = one cell x a quarter of the slice's columns =====
    const int ew = warp - 2 - kI2BuildWarps;  // 0..15
    const int q = warp & 3;                   // TMEM lane quarter
    const int cg = ew >> 2;                   // which quarter of the slice's columns
    const int cols_per = n_w / 4;             // 32 or 16
    const uint32_t sc = ptx::smem_u32(s_sc), sh = ptx::smem_u32(s_sh);
    const int cell = q * 32 + lane;           // lane of the accumulator = cell of the tile
    const int gr = cell / kMaskPadW, gc = cell % kMaskPadW;  // grid row within the tile, cell within the grid row
    const bool stored = gr < kI2TileRows;     // lanes 120-127 belong to the next grid row: never stored
    const uint32_t sw = static_cast<uint32_t>(gc & 7);
    const uint32_t stage_u32 = ptx::smem_u32(smem_stage);
    uint32_t tile = 0;
    for (int i = 0; i < n_pos; ++i) {
      const int b = cta_in_slice + i * ctas_per_slice;
      const int sb = i % kI2MaskStages;
      const uint32_t gp = ptx::smem_u32(s_gs + sb * 128 + cg * cols_per);
      for (int t = 0; t < 4; ++t, ++tile) {
        const uint32_t as = tile % kI2AccStages;
        ptx::mbar_wait(&acc_full[as], (tile / kI2AccStages) & 1u);
        ptx::tc_fence_after_sync();
        // all the arithmetic first, results packed in registers; the staging boxes are only touched afterwards, so the TMA
        // engine reads the previous tile's boxes under this tile's math
        uint4 rr[2][2], pp[2][2];  // [chunk][half]: fp16 stream / activated operand
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          if (h < cols_per / 16) {
            const int col = cg * cols_per + h * 16;  // within the slice
            uint32_t v[16];
            if (!(debug & 4)) {
              ptx::tmem_ld_32x16(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * static_cast<uint32_t>(n_w) + static_cast<uint32_t>(col), v);
              ptx::tmem_ld_wait();
            } else {
#pragma unroll
              for (int k = 0; k < 16; ++k) v[k] = 0;
            }
            if (h == cols_per / 16 - 1) {  // this warp's part of the accumulator is in registers
              ptx::tc_fence_before_sync();
              __syncwarp();
              if (lane == 0) ptx::mbar_arrive(&acc_empty[as]);
            }
            float x[16], a[16];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const float4 g4 = ptx::lds_f4(gp + static_cast<uint32_t>(4 * h + k) * 16u);
              x[4 * k] = __uint_as_float(v[4 * k]) + g4.x;
              x[4 * k + 1] = __uint_as_float(v[4 * k + 1]) + g4.y;
              x[4 * k + 2] = __uint_as_float(v[4 * k + 2]) + g4.z;
              x[4 * k + 3] = __uint_as_float(v[4 * k + 3]) + g4.w;
            }
            // the position's bias has been read for the last time.  (Arriving before these reads - next to acc_empty above -
            // let the stager's next bulk copy land under them: wrong biases for the first positions of a CTA, while the bias
            // rows are still L2 hits)
            if (t == 3 && h == cols_per / 16 - 1) {
              __syncwarp();
              if (lane == 0) ptx::mbar_arrive(&stage_free[sb]);
            }
            bn_mish8(x, a, sc, sh, col);
            bn_mish8(x + 8, a + 8, sc, sh, col + 8);
            rr[h][0] = make_uint4(tc_pack_f16(x[0], x[1]), tc_pack_f16(x[2], x[3]), tc_pack_f16(x[4], x[5]), tc_pack_f16(x[6], x[7]));
            rr[h][1] = make_uint4(tc_pack_f16(x[8], x[9]), tc_pack_f16(x[10], x[11]), tc_pack_f16(x[12], x[13]), tc_pack_f16(x[14], x[15]));
            pp[h][0] = make_uint4(tc_pack_act(a[0], a[1], f16), tc_pack_act(a[2], a[3], f16), tc_pack_act(a[4], a[5], f16), tc_pack_act(a[6], a[7], f16));
            pp[h][1] = make_uint4(tc_pack_act(a[8], a[9], f16), tc_pack_act(a[10], a[11], f16), tc_pack_act(a[12], a[13], f16), tc_pack_act(a[14], a[15], f16));
          }
        }
        // the previous tile's stores (issued by lane 0 of every epilogue warp, see below) have read the staging boxes
        if (lane == 0) ptx::bulk_wait_read<0>();
        ptx::named_bar_sync(2, kI2EpiWarps * 32);
        if (stored) {
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            if (h < cols_per / 16) {
              const int col = cg * cols_per + h * 16;
              const int slab = col >> 6, chunk = (col & 63) >> 3;  // 64-channel slab of the slice, 16-byte chunk within the 128-byte row
              const uint32_t box = stage_u32 + static_cast<uint32_t>((slab * kI2TileRows + gr) * kI2RowBoxBytes + gc * 128);
              const uint32_t c0 = ((static_cast<uint32_t>(chunk)) ^ sw) << 4, c1 = ((static_cast<uint32_t>(chunk) + 1u) ^ sw) << 4;
              ptx::sts_u4(box + c0, rr[h][0]);
              ptx::sts_u4(box + c1, rr[h][1]);
              const uint32_t abox = box + static_cast<uint32_t>(n_slabs * kI2TileRows * kI2RowBoxBytes);
              ptx::sts_u4(abox + c0, pp[h][0]);
              ptx::sts_u4(abox + c1, pp[h][1]);
            }
          }
        }
        __syncwarp();
        ptx::fence_proxy_async();
        ptx::named_bar_sync(2, kI2EpiWarps * 32);
        if (lane == 0) {
          // 2 outputs x n_slabs x 5 grid rows stores per tile, dealt round-robin to the epilogue warps (one thread issuing all of
          // them, and everyone waiting for it at the next tile's barrier, was a third of the tile time)
          const int n_stores = 2 * n_slabs * kI2TileRows;
          for (int sidx = ew; sidx < n_stores; sidx += kI2EpiWarps) {
            const int arr = sidx / (n_slabs * kI2TileRows), rem = sidx % (n_slabs * kI2TileRows);
            const int slab = rem / kI2TileRows, r = rem % kI2TileRows;
            const int br = kI2TileRows * t + r;  // board row
            if (br >= P3_BOARD_LEN || (debug & 2)) continue;
            // the 19 board cells of the grid row start 2 cells (256 B) into its staging box; the 128B swizzle is a function of
            // the shared-memory address, so the shifted source needs no re-arrangement
            const uint32_t src = stage_u32 + static_cast<uint32_t>(((arr * n_slabs + slab) * kI2TileRows + r) * kI2RowBoxBytes) + 2 * 128;
            tma_store_4d(arr == 0 ? &map_raw : &map_act, src, n0 + slab * 64, 0, br, b);
          }
          ptx::bulk_commit();
        }
      }
    }
    if (lane == 0) ptx::bulk_wait_all();
  }

  ptx::tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after_sync();
    ptx::tmem_dealloc(tmem_base, static_cast<uint32_t>(kI2AccStages * n_w));
  }
}

size_t init_tc2_smem_bytes(int n_w) {
  return static_cast<size_t>(2 * (n_w / 64) * kI2TileRows * kI2RowBoxBytes) + static_cast<size_t>(n_w / 8) * kI2WSbo + 2 * kI2EBytes + 256 * 16 +
         kI2MaskStages * (kI2MaskElems * 2 + 128 * 4) + 2 * 128 * 4 + 256 + 1024;
}

}  // namespace

int init_tc2_slice_width(int C) { return C % 128 == 0 ? 128 : 64; }

bool init_tc2_supported(int nplanes, int nscalars, int C) {
  (void)nscalars;
  return nplanes <= 16 && C % 64 == 0 && C >= 64 && init_tc2_smem_bytes(init_tc2_slice_width(C)) <= 227 * 1024;
}

// [25][nplanes][C] fp32 tap-major table -> per N slice of n_w channels, K-major core-matrix order [n / 8][k / 8][n % 8][k % 8],
// k = tap * 16 + plane (the layout of init_tc.cu with a wider slice)
int init_tc2_pack_weights(const float* wt, int nplanes, int C, std::vector<__nv_bfloat16>& out, bool op_f16) {
  const int n_w = init_tc2_slice_width(C);
  out.assign(static_cast<size_t>(C) * kI2K, __float2bfloat16(0.0f));
  auto to_op = [&](float v) {
    if (!op_f16) return __float2bfloat16(v);
    const __half h = __float2half_rn(v);
    return *reinterpret_cast<const __nv_bfloat16*>(&h);
  };
  for (int c = 0; c < C; ++c) {
    const int slice = c / n_w, nl = c % n_w;
    for (int t = 0; t < kI2Taps; ++t)
      for (int p = 0; p < nplanes; ++p) {
        const int k = t * 16 + p;
        const size_t idx = static_cast<size_t>(slice) * n_w * kI2K + (static_cast<size_t>(nl / 8) * kI2Kc + k / 8) * 64 + (nl % 8) * 8 + (k % 8);
        out[idx] = to_op(wt[(static_cast<size_t>(t) * nplanes + p) * C + c]);
      }
  }
  return P3_OK;
}

struct InitTc2Plan {
  CUtensorMap map_raw, map_act;
  int n = 0, C = 0, grid = 0, f16 = 0, n_w = 0, debug = 0;
  size_t smem = 0;
  const uint16_t* masks_padded = nullptr;
  const float *gs = nullptr, *scale = nullptr, *shift = nullptr;
  const __nv_bfloat16* w_packed = nullptr;
};

namespace {
// the [n x 400, C] 16-bit output viewed as (channel, board column c, board row r, position): point (r, c) of position b is row
// b * 400 + 20 + r * 20 + c.  Extents 19 x 19: the zero column c = 19 and everything outside the board are out of bounds = clipped.
int make_out_map(CUtensorMap* map, void* base, CUtensorMapDataType dt, int C, int n) {
  EncodeTiledFn fn = tc_encode_fn();
  if (!fn) return fail(P3_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
  cuuint64_t gdim[4] = {static_cast<cuuint64_t>(C), P3_BOARD_LEN, P3_BOARD_LEN, static_cast<cuuint64_t>(n)};
  cuuint64_t gstride[3] = {static_cast<cuuint64_t>(C) * 2, static_cast<cuuint64_t>(kRowPitch) * C * 2, static_cast<cuuint64_t>(kRowsPerPos) * C * 2};
  cuuint32_t box[4] = {64, P3_BOARD_LEN, 1, 1};
  cuuint32_t es[4] = {1, 1, 1, 1};
  void* origin = static_cast<uint8_t*>(base) + static_cast<size_t>(kRowBase) * C * 2;
  CUresult r = fn(map, dt, 4, origin, gdim, gstride, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(P3_ERR_CUDA, "cuTensorMapEncodeTiled (init_tc2 output map) failed: " + std::to_string(r));
  return P3_OK;
}
}  // namespace

int init_tc2_plan_create(const uint16_t* masks_padded, const float* gs, int n, int C, const __nv_bfloat16* w_packed, __half* raw_out,
                         __nv_bfloat16* act_out, const float* scale, const float* shift, InitTc2Plan** out, bool op_f16) {
  InitTc2Plan* p = new InitTc2Plan();
  p->masks_padded = masks_padded; p->gs = gs; p->n = n; p->C = C; p->w_packed = w_packed; p->scale = scale; p->shift = shift;
  p->f16 = op_f16 ? 1 : 0;
  p->n_w = init_tc2_slice_width(C);
  p->smem = init_tc2_smem_bytes(p->n_w);
  int rc = make_out_map(&p->map_raw, raw_out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, C, n);
  if (rc == P3_OK) rc = make_out_map(&p->map_act, act_out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, C, n);
  if (rc == P3_OK) {
    cudaError_t e = cudaFuncSetAttribute(init_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) rc = fail(P3_ERR_CUDA, std::string("init_tc2 smem attribute: ") + cudaGetErrorString(e));
  }
  if (rc != P3_OK) {
    delete p;
    return rc;
  }
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int n_slices = C / p->n_w;
  p->grid = std::max(n_slices, std::min(sms, n * n_slices) / n_slices * n_slices);
  if (const char* d = std::getenv("P3_INIT_TC2_DEBUG")) p->debug = std::atoi(d);  // perf / bring-up ablations (results are wrong when set)
  *out = p;
  return P3_OK;
}

void init_tc2_plan_destroy(InitTc2Plan* p) { delete p; }

int init_tc2_launch(const InitTc2Plan* p, cudaStream_t stream) {
  init_tc2_kernel<<<p->grid, kI2Threads, p->smem, stream>>>(p->map_raw, p->map_act, p->masks_padded, p->gs, p->n, p->C, p->n_w, p->w_packed,
                                                            p->scale, p->shift, p->f16, p->debug);
  P3_CUDA(cudaGetLastError());
  return P3_OK;
}

}  // namespace p3
