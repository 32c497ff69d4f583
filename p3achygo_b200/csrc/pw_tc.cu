// 1x1 ("pointwise") trunk layers of the bf16 engine that are not part of a fused block boundary (chain_tc.cu): the first
// block's reduce conv, the convs around the broadcast mix, the last block's expand conv.  One GEMM per launch,
//     x' = (x +) W * t ;  raw stream <- x' (fp16, optional) ;  act <- act(x') (bf16: the next layer's BN + mish folded in)
// on a CTA pair (tcgen05 cta_group::2, M = 256), with the structure that made the fused kernel HBM-efficient:
//   warp 0      TMA producer: this CTA's half of the weights once (resident), then its 128-row A tiles through a slab ring
//   warp 1      MMA issuer (leader CTA): two accumulator stages in TMEM, so tile it + 1 is multiplied during epilogue it
//   warps 2-17  epilogue, 4 warps per TMEM lane quarter, thread = one row x 16 of a slab's 64 columns
//   warps 18-21 one I/O warp per quarter: every TMA load (residual) and store of the quarter's staging boxes; the residual is
//               loaded into the very box its x' is stored from (in place), boxes are prepared several steps ahead.
// Per 64-column slab a quarter fills one box (raw or act) or two (raw and act).
#include <cuda.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>

#include "common.cuh"
#include "math.cuh"
#include "ptx.cuh"
#include "tc_util.cuh"

namespace p3 {

constexpr int kPwEpiWarps = 16;
constexpr int kPwThreads = (2 + kPwEpiWarps + 4) * 32;  // 704
constexpr int kPwSlabBytes = 128 * 128;                 // 128 rows x 64 bf16
constexpr int kPwBoxBytes = 32 * 128;                   // 32 rows x 64 two-byte elements (128B swizzle)
constexpr int kPwMaxBoxes = 8;
constexpr int kPwMaxStages = 6;
constexpr int kPwMaxN = 256;
constexpr int kPwSmemBudget = 227 * 1024;
constexpr int kPwBarRegion = 768;
constexpr int kPwMisc = 1024 /*align*/ + kPwBarRegion + 2 * kPwMaxN * 4;

struct TcPwPlan {
  CUtensorMap map_a1, map_w1, map_res, map_raw, map_act;
  int rows = 0, k1 = 0, n1 = 0, n_tile = 0, a1_stages = 0, n_boxes = 0, grid = 0;
  int has_res = 0, has_raw = 0, has_act = 0, act_mode = kActNone, f16 = 0, reverse = 0;
  size_t smem_bytes = 0;
  const float *scale = nullptr, *shift = nullptr;
};

namespace {

int pw_pick_n_tile(int n1) {  // largest divisor of n1 that is <= 256 and a multiple of 64
  for (int n = kPwMaxN; n >= 64; n -= 64)
    if (n1 % n == 0) return n;
  return 0;
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kPwThreads, 1)
tc_pw_pair_kernel(const __grid_constant__ CUtensorMap map_a1, const __grid_constant__ CUtensorMap map_w1,
                  const __grid_constant__ CUtensorMap map_res, const __grid_constant__ CUtensorMap map_raw,
                  const __grid_constant__ CUtensorMap map_act, int rows, int k1, int n1, int n_tile, int a1_stages, int n_boxes,
                  int has_res, int has_raw, int has_act, int act_mode, const float* __restrict__ scale,
                  const float* __restrict__ shift, int f16, int reverse) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int k1_slabs = k1 / 64, n_slabs = n_tile / 64;
  const int w_slab_bytes = (n_tile / 2) * 128, w_bytes = k1_slabs * w_slab_bytes;
  uint8_t* smem_w = smem;
  uint8_t* smem_a1 = smem_w + w_bytes;
  uint8_t* smem_box = smem_a1 + a1_stages * kPwSlabBytes;  // [4 quarters][n_boxes]
  uint64_t* a1_full = reinterpret_cast<uint64_t*>(smem_box + 4 * n_boxes * kPwBoxBytes);
  uint64_t* a1_empty = a1_full + kPwMaxStages;
  uint64_t* acc_full = a1_empty + kPwMaxStages;   // [2]
  uint64_t* acc_empty = acc_full + 2;             // [2] leader's: the epilogue warps of both CTAs
  uint64_t* w_bar = acc_empty + 2;
  uint64_t* box_ready = w_bar + 1;                        // [4][kPwMaxBoxes]
  uint64_t* box_written = box_ready + 4 * kPwMaxBoxes;    // [4][kPwMaxBoxes]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(box_written + 4 * kPwMaxBoxes);
  float* s_scale = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(a1_full) + kPwBarRegion);  // x log2(e), see bn_mish8
  float* s_shift = s_scale + kPwMaxN;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = ptx::cluster_ctarank();
  const int n_slices = n1 / n_tile;
  const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
  const int slice = pair % n_slices;
  const int n0 = slice * n_tile;
  const int pair_in_slice = pair / n_slices, pairs_per_slice = n_pairs / n_slices;
  const int m_tiles = pair_tile_count(rows);  // position-aligned pair tiles (common.cuh)
  const int rev_last = reverse ? m_tiles - 1 : -1;
  const int n_it = pair_in_slice < m_tiles ? (m_tiles - pair_in_slice + pairs_per_slice - 1) / pairs_per_slice : 0;
  const int boxes_per_slab = (has_raw && has_act) ? 2 : 1;

  constexpr float kLog2e = 1.4426950408889634f;
  for (int c = threadIdx.x; c < n_tile; c += blockDim.x) {
    s_scale[c] = (act_mode == kActMishBN ? scale[n0 + c] : 1.0f) * kLog2e;
    s_shift[c] = (act_mode == kActMishBN ? shift[n0 + c] : 0.0f) * kLog2e;
  }
  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&map_a1);
    ptx::prefetch_tensormap(&map_w1);
    if (has_res) ptx::prefetch_tensormap(&map_res);
    if (has_raw) ptx::prefetch_tensormap(&map_raw);
    if (has_act) ptx::prefetch_tensormap(&map_act);
    for (int s = 0; s < kPwMaxStages; ++s) {
      ptx::mbar_init(&a1_full[s], 1);
      ptx::mbar_init(&a1_empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      ptx::mbar_init(&acc_full[s], 1);
      ptx::mbar_init(&acc_empty[s], 2 * kPwEpiWarps);
    }
    ptx::mbar_init(w_bar, 1);
    for (int s = 0; s < 4 * kPwMaxBoxes; ++s) {
      ptx::mbar_init(&box_ready[s], 1);
      ptx::mbar_init(&box_written[s], 4);
    }
    ptx::fence_mbar_init();
  } else if (warp == 1) {
    ptx::tmem_alloc_pair(tmem_ptr, 512);
    ptx::tmem_relinquish_pair();
  }
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::cluster_sync_all();
  ptx::tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_ptr;
  if (threadIdx.x == 0) ptx::griddep_launch_dependents();

  if (warp == 0) {
    // ===== TMA producer (both CTAs) =====
    const uint32_t w_bar_leader = ptx::mapa_shared(ptx::smem_u32(w_bar), 0);
    if (ptx::elect_one()) {
      if (rank == 0) ptx::mbar_arrive_expect_tx(w_bar, static_cast<uint32_t>(2 * w_bytes));
      for (int ks = 0; ks < k1_slabs; ++ks)
        ptx::tma_load_2d_pair(smem_w + ks * w_slab_bytes, &map_w1, w_bar_leader, ks * 64, n0 + static_cast<int>(rank) * (n_tile / 2));
    }
    __syncwarp();
    ptx::griddep_wait();  // the weights above do not depend on the previous kernel; the activations do
    int stage = 0;
    uint32_t phase = 0;
    for (int it = 0; it < n_it; ++it) {
      const int m0 = pair_tile_row0(pair_in_slice + it * pairs_per_slice, static_cast<int>(rank), rev_last);
      for (int ks = 0; ks < k1_slabs; ++ks) {
        ptx::mbar_wait(&a1_empty[stage], phase ^ 1);
        if (ptx::elect_one()) {
          if (rank == 0) ptx::mbar_arrive_expect_tx(&a1_full[stage], 2 * kPwSlabBytes);
          ptx::tma_load_2d_pair_h<P3_HINT_ACT_LOAD>(smem_a1 + stage * kPwSlabBytes, &map_a1, ptx::mapa_shared(ptx::smem_u32(&a1_full[stage]), 0),
                                ks * 64, m0);
        }
        __syncwarp();
        if (++stage == a1_stages) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    if (rank == 0) {
      // ===== MMA issuer (leader CTA only) =====
      const uint32_t idesc = ptx::make_idesc_op(256, n_tile, f16);
      const uint32_t w_lo = ptx::desc_lo_sw128(ptx::smem_u32(smem_w)), a1_lo = ptx::desc_lo_sw128(ptx::smem_u32(smem_a1));
      ptx::mbar_wait_cluster(w_bar, 0);
      ptx::tc_fence_after_sync();
      int stage = 0;
      uint32_t phase = 0;
      for (int it = 0; it < n_it; ++it) {
        const int as = it & 1;
        ptx::mbar_wait_cluster(&acc_empty[as], ((static_cast<uint32_t>(it) >> 1) & 1u) ^ 1u);
        ptx::tc_fence_after_sync();
        const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(as * n_tile);
        for (int ks = 0; ks < k1_slabs; ++ks) {
          ptx::mbar_wait_cluster(&a1_full[stage], phase);
          ptx::tc_fence_after_sync();
          const uint32_t a_lo = a1_lo + static_cast<uint32_t>(stage) * (kPwSlabBytes / 16);
          const uint32_t b_lo = w_lo + static_cast<uint32_t>(ks) * static_cast<uint32_t>(w_slab_bytes / 16);
          if (ptx::elect_one()) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
              ptx::umma_f16_pair_lohi(tmem_d, a_lo + 2 * k, ptx::desc_hi_sw128(), b_lo + 2 * k, ptx::desc_hi_sw128(), idesc,
                                      (ks > 0 || k > 0) ? 1u : 0u);
            ptx::umma_commit_pair(&a1_empty[stage]);
          }
          __syncwarp();
          if (++stage == a1_stages) {
            stage = 0;
            phase ^= 1;
          }
        }
        if (ptx::elect_one()) ptx::umma_commit_pair(&acc_full[as]);
        __syncwarp();
      }
    }
  } else if (warp >= 2 + kPwEpiWarps) {
    // ===== I/O warp of quarter q.  Box-step sequence (the epilogue warps walk the same one): for each tile, for each
    // 64-column slab: [the residual / raw box (or the act box when no raw stream is written)], then [the act box] if
    // both streams are written.  `ahead` runs n_boxes - 1 box-steps in front of `cur`. =====
    const int q = warp - (2 + kPwEpiWarps);
    uint64_t* my_ready = box_ready + kPwMaxBoxes * q;
    uint64_t* my_written = box_written + kPwMaxBoxes * q;
    uint8_t* my_box = smem_box + q * n_boxes * kPwBoxBytes;
    const int q_row = q * 32;  // within this CTA's 128 rows of a tile
    const int steps_per_tile = n_slabs * boxes_per_slab;
    const uint32_t total_steps = static_cast<uint32_t>(n_it) * steps_per_tile;
    auto row_of = [&](uint32_t s) {
      return pair_tile_row0(pair_in_slice + static_cast<int>(s / steps_per_tile) * pairs_per_slice, static_cast<int>(rank), rev_last) + q_row;
    };
    auto col_of = [&](uint32_t s) { return n0 + static_cast<int>((s % steps_per_tile) / boxes_per_slab) * 64; };
    auto prepare = [&](uint32_t s) {
      if (s >= total_steps) return;
      const uint32_t b = s % n_boxes;
      if (has_res && (s % boxes_per_slab) == 0) {
        ptx::mbar_arrive_expect_tx(&my_ready[b], kPwBoxBytes);
        ptx::tma_load_2d_h<P3_HINT_RES_LOAD>(my_box + b * kPwBoxBytes, &map_res, &my_ready[b], col_of(s), row_of(s));
      } else {
        ptx::mbar_arrive(&my_ready[b]);
      }
    };
    ptx::griddep_wait();  // residual loads / output stores touch buffers of the previous kernel
    if (lane == 0)
      for (uint32_t s = 0; s < static_cast<uint32_t>(n_boxes); ++s) prepare(s);  // all boxes start out free
    for (uint32_t s = 0; s < total_steps; ++s) {
      const uint32_t b = s % n_boxes;
      ptx::mbar_wait(&my_written[b], (s / n_boxes) & 1u);
      if (lane == 0) {
        const bool to_raw = has_raw && (s % boxes_per_slab) == 0;
        ptx::tma_store_2d_h<P3_HINT_ACT_STORE>(to_raw ? &map_raw : &map_act, ptx::smem_u32(my_box) + b * kPwBoxBytes, col_of(s), row_of(s));
        ptx::bulk_commit();
        ptx::bulk_wait_read<1>();  // the previous box-step's store has read its box: it serves box-step s - 1 + n_boxes
        if (s > 0) prepare(s - 1 + n_boxes);
      }
      __syncwarp();
    }
    if (lane == 0) ptx::bulk_wait_all();
  } else {
    // ===== epilogue (both CTAs) =====
    const int ew = warp - 2;
    const int q = warp & 3;
    const int cg = ew >> 2;
    const uint32_t sw = static_cast<uint32_t>(lane & 7);
    const uint32_t ch0 = ((2u * cg) ^ sw) << 4, ch1 = ((2u * cg + 1u) ^ sw) << 4;
    const uint32_t box_base = ptx::smem_u32(smem_box) + static_cast<uint32_t>(q * n_boxes) * kPwBoxBytes + static_cast<uint32_t>(lane) * 128u;
    uint64_t* my_ready = box_ready + kPwMaxBoxes * q;
    uint64_t* my_written = box_written + kPwMaxBoxes * q;
    const uint32_t acc_empty_l = ptx::mapa_shared(ptx::smem_u32(acc_empty), 0);
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    const uint32_t sc = ptx::smem_u32(s_scale), sh = ptx::smem_u32(s_shift);
    uint32_t ob = 0, ob_phase = 0;  // staging box of the current box-step (ordinal mod n_boxes) and its use parity
    auto next_box = [&]() {
      if (++ob == static_cast<uint32_t>(n_boxes)) {
        ob = 0;
        ob_phase ^= 1u;
      }
    };
    for (int it = 0; it < n_it; ++it) {
      const int as = it & 1;
      const int m = pair_tile_row0(pair_in_slice + it * pairs_per_slice, static_cast<int>(rank), rev_last) + q * 32 + lane;
      const bool live = m < rows && row_is_live(m % kRowsPerPos);
      ptx::mbar_wait(&acc_full[as], (static_cast<uint32_t>(it) >> 1) & 1u);
      ptx::tc_fence_after_sync();
      for (int j = 0; j < n_slabs; ++j) {
        const int col = j * 64 + cg * 16;
        uint32_t v[16];
        ptx::tmem_ld_32x16(lane_addr + static_cast<uint32_t>(as * n_tile + col), v);
        uint32_t obuf = box_base + ob * kPwBoxBytes;
        ptx::mbar_wait(&my_ready[ob], ob_phase);
        float x[16];
        if (has_res) {
          const float4 t0 = ptx::lds_f4(obuf + ch0), t1 = ptx::lds_f4(obuf + ch1);
          const uint32_t u[8] = {__float_as_uint(t0.x), __float_as_uint(t0.y), __float_as_uint(t0.z), __float_as_uint(t0.w),
                                 __float_as_uint(t1.x), __float_as_uint(t1.y), __float_as_uint(t1.z), __float_as_uint(t1.w)};
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float2 f2 = __half22float2(*reinterpret_cast<const __half2*>(&u[i]));
            x[2 * i] = f2.x;
            x[2 * i + 1] = f2.y;
          }
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i) x[i] = 0.0f;
        }
        ptx::tmem_ld_wait();
        if (j == n_slabs - 1) {  // the accumulator is in registers: hand the stage back to the MMA warp
          ptx::tc_fence_before_sync();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive_remote(acc_empty_l + 8u * static_cast<uint32_t>(as));
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) x[i] += __uint_as_float(v[i]);
        if (has_raw) {
          // padding rows need no masking: their A rows and residual rows are zeros (layout invariant), so x' = 0 exactly
          ptx::sts_u4(obuf + ch0, make_uint4(tc_pack_f16(x[0], x[1]), tc_pack_f16(x[2], x[3]), tc_pack_f16(x[4], x[5]), tc_pack_f16(x[6], x[7])));
          ptx::sts_u4(obuf + ch1, make_uint4(tc_pack_f16(x[8], x[9]), tc_pack_f16(x[10], x[11]), tc_pack_f16(x[12], x[13]), tc_pack_f16(x[14], x[15])));
          ptx::fence_proxy_async();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(&my_written[ob]);
          next_box();
        }
        if (has_act) {
          float a[16];
          if (act_mode == kActIdentity) {
#pragma unroll
            for (int i = 0; i < 16; ++i) a[i] = x[i];
          } else {
            bn_mish8(x, a, sc, sh, col);
            bn_mish8(x + 8, a + 8, sc, sh, col + 8);
          }
          uint4 p0 = make_uint4(tc_pack_act(a[0], a[1], f16), tc_pack_act(a[2], a[3], f16), tc_pack_act(a[4], a[5], f16), tc_pack_act(a[6], a[7], f16));
          uint4 p1 = make_uint4(tc_pack_act(a[8], a[9], f16), tc_pack_act(a[10], a[11], f16), tc_pack_act(a[12], a[13], f16), tc_pack_act(a[14], a[15], f16));
          if (!live) p0 = p1 = make_uint4(0, 0, 0, 0);  // padding rows of the layout stay zero
          if (has_raw) {  // second box of the slab
            obuf = box_base + ob * kPwBoxBytes;
            ptx::mbar_wait(&my_ready[ob], ob_phase);
          }
          ptx::sts_u4(obuf + ch0, p0);
          ptx::sts_u4(obuf + ch1, p1);
          ptx::fence_proxy_async();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(&my_written[ob]);
          next_box();
        }
      }
    }
  }

  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::cluster_sync_all();
  if (warp == 1) {
    ptx::tc_fence_after_sync();
    ptx::tmem_dealloc_pair(tmem_base, 512);
  }
}

}  // namespace

bool tc_pw_supported(int k1, int n1) {
  if (k1 <= 0 || k1 % 64 != 0 || n1 <= 0) return false;
  const int n_tile = pw_pick_n_tile(n1);
  if (n_tile == 0) return false;
  const size_t w_bytes = static_cast<size_t>(k1 / 64) * (n_tile / 2) * 128;
  return w_bytes + 2 * kPwSlabBytes + 4 * 3 * kPwBoxBytes + kPwMisc <= static_cast<size_t>(kPwSmemBudget);
}

int tc_pw_plan_create(const __nv_bfloat16* in, const __nv_bfloat16* w, int rows, int k1, int n1, const ConvEpilogue& ep,
                      TcPwPlan** out) {
  if (!tc_pw_supported(k1, n1)) return fail(P3_ERR_UNSUPPORTED, "tc_pw: shape not supported");
  if ((ep.residual || ep.raw_out) && !ep.raw_f16) return fail(P3_ERR_UNSUPPORTED, "tc_pw: the residual stream must be fp16");
  if (!ep.raw_out && !ep.act_out) return fail(P3_ERR_INVALID_ARG, "tc_pw: nothing to write");
  TcPwPlan* p = new TcPwPlan();
  p->f16 = ep.op_f16 ? 1 : 0;
  p->rows = rows;
  p->k1 = k1;
  p->n1 = n1;
  p->n_tile = pw_pick_n_tile(n1);
  p->has_res = ep.residual != nullptr;
  p->has_raw = ep.raw_out != nullptr;
  p->has_act = ep.act_out != nullptr;
  p->act_mode = ep.act_mode;
  p->scale = ep.scale;
  p->shift = ep.shift;
  const size_t w_bytes = static_cast<size_t>(k1 / 64) * (p->n_tile / 2) * 128;
  // split what is left of shared memory between the A ring (at most two tiles) and the staging boxes
  const int k1_slabs = k1 / 64;
  size_t left = static_cast<size_t>(kPwSmemBudget) - kPwMisc - w_bytes;
  int boxes = 3, stages = 2;
  const int want_boxes = (p->has_raw && p->has_act) ? 6 : 5;
  while (true) {
    bool grew = false;
    if (boxes < want_boxes && static_cast<size_t>(stages) * kPwSlabBytes + static_cast<size_t>(4 * (boxes + 1)) * kPwBoxBytes <= left) {
      ++boxes;
      grew = true;
    }
    if (stages < std::min(kPwMaxStages, 2 * k1_slabs) &&
        static_cast<size_t>(stages + 1) * kPwSlabBytes + static_cast<size_t>(4 * boxes) * kPwBoxBytes <= left) {
      ++stages;
      grew = true;
    }
    if (!grew) break;
  }
  p->n_boxes = boxes;
  p->a1_stages = stages;
  p->smem_bytes = w_bytes + static_cast<size_t>(stages) * kPwSlabBytes + static_cast<size_t>(4 * boxes) * kPwBoxBytes + kPwMisc;
  const CUtensorMapDataType bf = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, hf = CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
  const CUtensorMapSwizzle sw = CU_TENSOR_MAP_SWIZZLE_128B;
  int rc = tc_make_map_2d(&p->map_a1, in, bf, 2, k1, rows, 64, 128, sw);
  if (rc == P3_OK) rc = tc_make_map_2d(&p->map_w1, w, bf, 2, k1, n1, 64, p->n_tile / 2, sw);
  p->map_res = p->map_a1;  // placeholders, never dereferenced when the stream is absent
  p->map_raw = p->map_a1;
  p->map_act = p->map_a1;
  if (rc == P3_OK && ep.residual) rc = tc_make_map_2d(&p->map_res, ep.residual, hf, 2, n1, rows, 64, 32, sw);
  if (rc == P3_OK && ep.raw_out) rc = tc_make_map_2d(&p->map_raw, ep.raw_out, hf, 2, n1, rows, 64, 32, sw);
  if (rc == P3_OK && ep.act_out) rc = tc_make_map_2d(&p->map_act, ep.act_out, bf, 2, n1, rows, 64, 32, sw);
  if (rc == P3_OK) {
    cudaError_t e = cudaFuncSetAttribute(tc_pw_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kPwSmemBudget);
    if (e != cudaSuccess) rc = fail(P3_ERR_CUDA, std::string("cudaFuncSetAttribute(smem): ") + cudaGetErrorString(e));
  }
  if (rc != P3_OK) {
    delete p;
    return rc;
  }
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int max_pairs = sms / 2;
  {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(static_cast<unsigned>(sms / 2 * 2));
    cfg.blockDim = dim3(kPwThreads);
    cfg.dynamicSmemBytes = p->smem_bytes;
    int n_clusters = 0;
    if (cudaOccupancyMaxActiveClusters(&n_clusters, tc_pw_pair_kernel, &cfg) == cudaSuccess && n_clusters > 0)
      max_pairs = std::min(max_pairs, n_clusters);
    else
      cudaGetLastError();
  }
  const int n_slices = n1 / p->n_tile;
  const int m_tiles = pair_tile_count(rows);
  const int pairs = std::max(n_slices, std::min(max_pairs, m_tiles * n_slices) / n_slices * n_slices);
  p->grid = 2 * pairs;
  *out = p;
  return P3_OK;
}

void tc_pw_plan_destroy(TcPwPlan* p) { delete p; }
void tc_pw_plan_set_reverse(TcPwPlan* p, bool reverse) {
  if (p) p->reverse = reverse ? 1 : 0;
}

int tc_pw_launch(const TcPwPlan* p, cudaStream_t stream) {
  P3_CUDA(tc_launch_pdl(tc_pw_pair_kernel, p->grid, kPwThreads, p->smem_bytes, stream, p->map_a1, p->map_w1, p->map_res, p->map_raw,
                        p->map_act, p->rows, p->k1, p->n1, p->n_tile, p->a1_stages, p->n_boxes, p->has_res, p->has_raw, p->has_act,
                        p->act_mode, p->scale, p->shift, p->f16, p->reverse));
  return P3_OK;
}

}  // namespace p3
