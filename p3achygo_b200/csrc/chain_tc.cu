// Fused block boundary of the bottleneck tower on a CTA pair (tcgen05 cta_group::2):
//
//     x'  = x + W_expand * t                       (1x1, Cb -> C, + residual stream; python/model.py:404-412)
//     u   = mish(BN_a(x'))                         (pre-activation of the NEXT block's first conv, model.py:276-281)
//     out = act2( W_reduce * u )                   (1x1, C -> Cb of the next block; act2 = that block's next BN + mish)
//
// Unfused these are two HBM-bound launches that round-trip the C-wide activated copy `u` through HBM (write 2*C
// bytes + read 2*C bytes per board row); here `u` only ever exists in shared memory as the A operand of the second
// GEMM.  Per 128-row tile and CTA: read t (Cb bf16) + x (C fp16), write x' (C fp16) + out (N2 bf16).
//
// One persistent CTA pair per TPC, 576 threads per CTA:
//   warp 0      TMA producer: this CTA's halves of W1 / W2 once (resident), then its own 128-row A1 tiles
//   warp 1      MMA issuer (leader CTA only): GEMM1 (M=256, N=N1) into acc1, then GEMM2 (M=256, N=N2) into acc2, one
//               64-wide K slab at a time as the epilogue warps publish the slabs of `u`
//   warps 2-17  epilogue, 4 warps per TMEM lane quarter; a quarter (32 rows) works through 64-column slabs:
//               tcgen05.ld -> + residual (TMA-prefetched box) -> x' box (TMA store) and the slab of `u` (swizzled K-major,
//               straight into the A2 ring) ; then the same for acc2 -> out boxes.  Quarters are independent of each
//               other (own named barrier, own staging, own bulk-store groups), so their phases interleave on the SM.
#include <cuda.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>

#include "common.cuh"
#include "math.cuh"
#include "ptx.cuh"
#include "tc_util.cuh"

namespace p3 {

constexpr int kChEpiWarps = 16;
constexpr int kChThreads = (2 + kChEpiWarps + 4) * 32;  // 704: producer, MMA, 16 epilogue, 4 I/O (one per TMEM lane quarter)
constexpr int kChSlabBytes = 128 * 128;   // 128 rows x 64 bf16 (one K slab of an A operand)
constexpr int kChBoxBytes = 32 * 128;     // a quarter's 32 rows x 64 two-byte elements
// A quarter cycles through a pool of n_boxes (3..5) staging boxes: TMA load of the residual -> the quarter's 4 warps replace
// it IN PLACE by x' (or write an epi2 output box) -> TMA store -> free.  Loads are issued n_boxes - 1 steps ahead of their use.
constexpr int kChMaxBoxes = 5;
constexpr int kChMaxN = 256;
constexpr int kChSmemBudget = 227 * 1024;

struct TcChainPlan {
  CUtensorMap map_a1, map_w1, map_w2, map_res, map_raw, map_out2;
  int rows = 0, k1 = 0, n1 = 0, n2 = 0, grid = 0, tmem_cols = 512, acc2_stages = 1, a1_stages = 2, n_boxes = 3;
  size_t smem_bytes = 0;
  const float *scale1 = nullptr, *shift1 = nullptr, *scale2 = nullptr, *shift2 = nullptr;
  int act2_mode = kActMishBN;
  int f16 = 0;                          // operands are IEEE fp16 instead of bf16 (P3_PRECISION_FP16)
  // tail form (the tower's last expand + the heads' 1x1 conv): u = x' (the heads take the raw trunk output), x' is not stored,
  // and `out` is written as fp32, channel-major [n2_valid, out_ld] (what heads.cu reads), straight from the registers
  int tail = 0, out_ld = 0, n2_valid = 0;
  int reverse = 0;                      // tile order (common.cuh, pair_tile_row0)
  int l2pf = 0;                         // residual boxes are prefetched into L2 this many steps beyond the box pool's look-ahead
  float* out_t = nullptr;
  unsigned long long* trace = nullptr;  // P3_TC_TRACE
};

namespace {

// resident weight halves + the A2 ring; the A1 ring (a1_stages slabs) and the box pool (n_boxes per quarter) share the rest
__host__ __device__ constexpr int chain_fixed_smem(int k1, int n1, int n2) {
  return (k1 / 64) * (n1 / 2) * 128 + (n1 / 64) * (n2 / 2) * 128 + 2 * kChSlabBytes;
}
constexpr int kChBarRegion = 512;
constexpr int kChBarBytes = kChBarRegion + 4 * kChMaxN * 4;

__device__ __forceinline__ void bn_mish16(const float* x, float* a, uint32_t sc, uint32_t sh, int col0) {
  bn_mish8(x, a, sc, sh, col0);
  bn_mish8(x + 8, a + 8, sc, sh, col0 + 8);
}

template <bool kTrace>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kChThreads, 1)
tc_chain_pair_kernel(const __grid_constant__ CUtensorMap map_a1, const __grid_constant__ CUtensorMap map_w1,
                     const __grid_constant__ CUtensorMap map_w2, const __grid_constant__ CUtensorMap map_res,
                     const __grid_constant__ CUtensorMap map_raw, const __grid_constant__ CUtensorMap map_out2, int rows, int k1,
                     int n1, int n2, int acc2_stages, int a1_stages, int n_boxes, int tmem_cols, const float* __restrict__ scale1,
                     const float* __restrict__ shift1, const float* __restrict__ scale2, const float* __restrict__ shift2,
                     int act2_mode, unsigned long long* trace, int f16, int tail, float* __restrict__ out_t, int out_ld,
                     int n2_valid, int reverse, int l2pf) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int k1_slabs = k1 / 64, n1_slabs = n1 / 64, n2_slabs = n2 / 64;
  const int w1_slab_bytes = (n1 / 2) * 128, w2_slab_bytes = (n2 / 2) * 128;
  const int w1_bytes = k1_slabs * w1_slab_bytes, w2_bytes = n1_slabs * w2_slab_bytes;
  uint8_t* smem_w1 = smem;
  uint8_t* smem_w2 = smem_w1 + w1_bytes;
  uint8_t* smem_a1 = smem_w2 + w2_bytes;
  uint8_t* smem_a2 = smem_a1 + a1_stages * kChSlabBytes;
  uint8_t* smem_box = smem_a2 + 2 * kChSlabBytes;  // [4 quarters][n_boxes] staging boxes
  uint64_t* a1_full = reinterpret_cast<uint64_t*>(smem_box + 4 * n_boxes * kChBoxBytes);
  uint64_t* a1_empty = a1_full + 4;
  uint64_t* acc1_full = a1_empty + 4;
  uint64_t* acc1_empty = acc1_full + 1;
  uint64_t* a2_full = acc1_empty + 1;
  uint64_t* a2_empty = a2_full + 2;
  uint64_t* acc2_full = a2_empty + 2;
  uint64_t* acc2_empty = acc2_full + 2;
  uint64_t* w_bar = acc2_empty + 2;
  uint64_t* box_ready = w_bar + 1;                     // [4 quarters][kChMaxBoxes] residual landed / box free for an epi2 output
  uint64_t* box_written = box_ready + 4 * kChMaxBoxes;    //                        the quarter's 4 warps have written the box
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(box_written + 4 * kChMaxBoxes);
  float* s_scale1 = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(a1_full) + kChBarRegion);
  float* s_shift1 = s_scale1 + kChMaxN;
  float* s_scale2 = s_shift1 + kChMaxN;
  float* s_shift2 = s_scale2 + kChMaxN;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = ptx::cluster_ctarank();
  const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
  const int m_tiles = pair_tile_count(rows);  // position-aligned pair tiles (common.cuh)
  const int rev_last = reverse ? m_tiles - 1 : -1;
  const int n_it = pair < m_tiles ? (m_tiles - pair + n_pairs - 1) / n_pairs : 0;  // tiles of this pair
  const int defer = acc2_stages >= 2 ? 1 : 0;  // epi2 runs one tile behind epi1 (needs the second acc2 stage)

  constexpr float kLog2e = 1.4426950408889634f;  // folded BN constants, pre-multiplied by log2(e) (see bn_mish8)
  for (int c = threadIdx.x; c < n1; c += blockDim.x) {
    s_scale1[c] = scale1[c] * kLog2e;
    s_shift1[c] = shift1[c] * kLog2e;
  }
  for (int c = threadIdx.x; c < n2; c += blockDim.x) {
    s_scale2[c] = (act2_mode == kActMishBN ? scale2[c] : 1.0f) * kLog2e;
    s_shift2[c] = (act2_mode == kActMishBN ? shift2[c] : 0.0f) * kLog2e;
  }
  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&map_a1);
    ptx::prefetch_tensormap(&map_w1);
    ptx::prefetch_tensormap(&map_w2);
    ptx::prefetch_tensormap(&map_res);
    ptx::prefetch_tensormap(&map_raw);
    ptx::prefetch_tensormap(&map_out2);
    for (int s = 0; s < 4; ++s) {
      ptx::mbar_init(&a1_full[s], 1);   // leader's is live: one arrive.expect_tx for both CTAs' bytes
      ptx::mbar_init(&a1_empty[s], 1);  // multicast commit
    }
    ptx::mbar_init(acc1_full, 1);
    ptx::mbar_init(acc1_empty, 2 * kChEpiWarps);  // leader's: the epilogue warps of both CTAs
    for (int s = 0; s < 2; ++s) {
      ptx::mbar_init(&a2_full[s], 8);  // leader's: the 4 I/O warps of both CTAs
      ptx::mbar_init(&a2_empty[s], 1);
      ptx::mbar_init(&acc2_full[s], 1);
      ptx::mbar_init(&acc2_empty[s], 2 * kChEpiWarps);
    }
    ptx::mbar_init(w_bar, 1);
    for (int s = 0; s < 4 * kChMaxBoxes; ++s) {
      ptx::mbar_init(&box_ready[s], 1);
      ptx::mbar_init(&box_written[s], 4);
    }
    ptx::fence_mbar_init();
  } else if (warp == 1) {
    ptx::tmem_alloc_pair(tmem_ptr, static_cast<uint32_t>(tmem_cols));
    ptx::tmem_relinquish_pair();
  }
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::cluster_sync_all();
  ptx::tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_ptr;
  if (threadIdx.x == 0) ptx::griddep_launch_dependents();

  if (warp == 0) {
    // ===== TMA producer (both CTAs): resident weight halves, then this CTA's A1 tiles =====
    const uint32_t w_bar_leader = ptx::mapa_shared(ptx::smem_u32(w_bar), 0);
    if (ptx::elect_one()) {
      if (rank == 0) ptx::mbar_arrive_expect_tx(w_bar, static_cast<uint32_t>(2 * (w1_bytes + w2_bytes)));
      for (int ks = 0; ks < k1_slabs; ++ks)
        ptx::tma_load_2d_pair(smem_w1 + ks * w1_slab_bytes, &map_w1, w_bar_leader, ks * 64, static_cast<int>(rank) * (n1 / 2));
      for (int j = 0; j < n1_slabs; ++j)
        ptx::tma_load_2d_pair(smem_w2 + j * w2_slab_bytes, &map_w2, w_bar_leader, j * 64, static_cast<int>(rank) * (n2 / 2));
    }
    __syncwarp();
    ptx::griddep_wait();  // the weights above do not depend on the previous kernel; the activations do
    int stage = 0;
    uint32_t phase = 0;
    for (int mt = pair; mt < m_tiles; mt += n_pairs) {
      const int m0 = pair_tile_row0(mt, static_cast<int>(rank), rev_last);
      for (int ks = 0; ks < k1_slabs; ++ks) {
        ptx::mbar_wait(&a1_empty[stage], phase ^ 1);
        if (ptx::elect_one()) {
          if (rank == 0) ptx::mbar_arrive_expect_tx(&a1_full[stage], 2 * kChSlabBytes);
          ptx::tma_load_2d_pair_h<P3_HINT_ACT_LOAD>(smem_a1 + stage * kChSlabBytes, &map_a1, ptx::mapa_shared(ptx::smem_u32(&a1_full[stage]), 0),
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
      const uint32_t idesc1 = ptx::make_idesc_op(256, n1, f16), idesc2 = ptx::make_idesc_op(256, n2, f16);
      const uint32_t w1_lo = ptx::desc_lo_sw128(ptx::smem_u32(smem_w1)), w2_lo = ptx::desc_lo_sw128(ptx::smem_u32(smem_w2));
      const uint32_t a1_lo = ptx::desc_lo_sw128(ptx::smem_u32(smem_a1)), a2_lo = ptx::desc_lo_sw128(ptx::smem_u32(smem_a2));
      ptx::mbar_wait_cluster(w_bar, 0);
      ptx::tc_fence_after_sync();
      int stage = 0;
      uint32_t phase = 0, g = 0;
      // GEMM1 of tile `t`: acc1 = A1 * W1^T (acc1 must have been drained by epi1 of tile t - 1)
      auto gemm1 = [&](int t) {
        ptx::mbar_wait_cluster(acc1_empty, (static_cast<uint32_t>(t) & 1u) ^ 1u);
        ptx::tc_fence_after_sync();
        for (int ks = 0; ks < k1_slabs; ++ks) {
          ptx::mbar_wait_cluster(&a1_full[stage], phase);
          ptx::tc_fence_after_sync();
          const uint32_t a_lo = a1_lo + static_cast<uint32_t>(stage) * (kChSlabBytes / 16);
          const uint32_t b_lo = w1_lo + static_cast<uint32_t>(ks) * static_cast<uint32_t>(w1_slab_bytes / 16);
          if (ptx::elect_one()) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
              ptx::umma_f16_pair_lohi(tmem_base, a_lo + 2 * k, ptx::desc_hi_sw128(), b_lo + 2 * k, ptx::desc_hi_sw128(), idesc1,
                                      (ks > 0 || k > 0) ? 1u : 0u);
            ptx::umma_commit_pair(&a1_empty[stage]);
          }
          __syncwarp();
          if (++stage == a1_stages) {
            stage = 0;
            phase ^= 1;
          }
        }
        if (ptx::elect_one()) ptx::umma_commit_pair(acc1_full);
        __syncwarp();
      };
      if (n_it > 0) gemm1(0);
      for (int it = 0; it < n_it; ++it) {
        // ---- GEMM2: acc2 = u * W2^T, one K slab of u at a time as the epilogue publishes it.  GEMM1 of the NEXT tile goes in
        // front of the last slab: acc1 is free as soon as epi1 has loaded its last columns (start of its last step), so the next
        // accumulator is ready before the epilogue gets to it.
        const int as = it % acc2_stages;
        ptx::mbar_wait_cluster(&acc2_empty[as], ((static_cast<uint32_t>(it / acc2_stages)) & 1u) ^ 1u);
        ptx::tc_fence_after_sync();
        const uint32_t tmem_d2 = tmem_base + static_cast<uint32_t>(n1 + as * n2);
        for (int j = 0; j < n1_slabs; ++j, ++g) {
          if (j == n1_slabs - 1 && it + 1 < n_it) gemm1(it + 1);
          const uint32_t slot = g & 1u;
          ptx::mbar_wait_cluster(&a2_full[slot], (g >> 1) & 1u);
          ptx::tc_fence_after_sync();
          const uint32_t a_lo = a2_lo + slot * (kChSlabBytes / 16);
          const uint32_t b_lo = w2_lo + static_cast<uint32_t>(j) * static_cast<uint32_t>(w2_slab_bytes / 16);
          if (ptx::elect_one()) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
              ptx::umma_f16_pair_lohi(tmem_d2, a_lo + 2 * k, ptx::desc_hi_sw128(), b_lo + 2 * k, ptx::desc_hi_sw128(), idesc2,
                                      (j > 0 || k > 0) ? 1u : 0u);
            ptx::umma_commit_pair(&a2_empty[slot]);
          }
          __syncwarp();
        }
        if (ptx::elect_one()) ptx::umma_commit_pair(&acc2_full[as]);
        __syncwarp();
      }
    }
  } else if (warp >= 2 + kChEpiWarps) {
    // ===== I/O warp of quarter q (both CTAs): every TMA load / store of the quarter's staging boxes, so that the epilogue
    // warps never issue TMA, never wait for a bulk group and never meet at a CTA barrier.  Per 64-column step of epi1:
    //   residual box read by the 4 warps -> refill it 2 steps ahead ; x' box written -> TMA store, publish the A2 slab to
    //   the leader's MMA warp ; the store issued one step earlier has read its box -> hand the box back.
    const int q = warp - (2 + kChEpiWarps);
    uint64_t* my_ready = box_ready + kChMaxBoxes * q;
    uint64_t* my_written = box_written + kChMaxBoxes * q;
    const uint32_t a2_full_l = ptx::mapa_shared(ptx::smem_u32(a2_full), 0);
    uint8_t* my_box = smem_box + q * n_boxes * kChBoxBytes;
    const int q_row = q * 32;  // within this CTA's 128 rows of a tile
    // The quarter's step sequence (the epilogue warps walk the same one): per iteration `it`, the n1_slabs epi1 steps of
    // tile it, then the n2_slabs epi2 steps of tile it - defer.  `ahead` runs n_boxes - 1 steps in front of `cur`
    // (the box of step cur - 1, just read by its store, serves step cur + n_boxes - 1).
    struct Cursor {
      int it, idx;   // iteration, index within the iteration's steps
      uint32_t s;    // step ordinal
      uint32_t b;    // box of the step: s mod n_boxes
      uint32_t ph;   // use parity of that box: (s / n_boxes) & 1
    };
    auto steps_in = [&](int it) { return (it < n_it ? n1_slabs : 0) + ((it - defer >= 0 && it - defer < n_it) ? n2_slabs : 0); };
    auto advance = [&](Cursor& c) {
      ++c.s;
      if (++c.b == static_cast<uint32_t>(n_boxes)) {
        c.b = 0;
        c.ph ^= 1u;
      }
      if (++c.idx >= steps_in(c.it)) {
        c.idx = 0;
        ++c.it;
      }
    };
    const int n_iter = n_it + defer;
    // make box s % n_boxes ready for step c: TMA-load the residual (epi1 step) or just hand the free box over (epi2 step)
    auto prepare = [&](const Cursor& c) {
      if (c.it >= n_iter) return;
      const uint32_t b = c.b;
      const bool is_epi1 = c.it < n_it && c.idx < n1_slabs;
      if (is_epi1) {
        ptx::mbar_arrive_expect_tx(&my_ready[b], kChBoxBytes);
        ptx::tma_load_2d_h<P3_HINT_RES_LOAD>(my_box + b * kChBoxBytes, &map_res, &my_ready[b], c.idx * 64, pair_tile_row0(pair + c.it * n_pairs, static_cast<int>(rank), rev_last) + q_row);
      } else {
        ptx::mbar_arrive(&my_ready[b]);
      }
    };
    Cursor cur{0, 0, 0, 0, 0}, ahead{0, 0, 0, 0, 0}, pf{0, 0, 0, 0, 0};
    // shapes whose resident weights leave only 3 boxes per quarter (2 loads in flight) cannot cover the HBM latency from shared
    // memory alone: their residual boxes are also prefetched into L2, l2pf steps further ahead
    auto prefetch = [&](const Cursor& c) {
      if (c.it < n_it && c.idx < n1_slabs)
        ptx::tma_prefetch_2d(&map_res, c.idx * 64, pair_tile_row0(pair + c.it * n_pairs, static_cast<int>(rank), rev_last) + q_row);
    };
    ptx::griddep_wait();  // residual loads / output stores touch buffers of the previous kernel
    if (n_it > 0) {
      for (int i = 0; i < n_boxes; ++i) {  // all boxes start out free
        if (lane == 0) prepare(ahead);
        advance(ahead);
      }
      pf = ahead;
      for (int i = 0; i < l2pf; ++i) {
        if (lane == 0) prefetch(pf);
        advance(pf);
      }
    }
    uint32_t g = 0;
    while (cur.it < n_iter) {
      const uint32_t b = cur.b;
      const bool is_epi1 = cur.it < n_it && cur.idx < n1_slabs;
      ptx::mbar_wait(&my_written[b], cur.ph);
      if (lane == 0) {
        if (is_epi1) {
          ptx::mbar_arrive_remote(a2_full_l + 8u * (g & 1u));  // this quarter's rows of the A2 slab are in place
          if (!tail)
            ptx::tma_store_2d_h<P3_HINT_RES_STORE>(&map_raw, ptx::smem_u32(my_box) + b * kChBoxBytes, cur.idx * 64,
                              pair_tile_row0(pair + cur.it * n_pairs, static_cast<int>(rank), rev_last) + q_row);
        } else if (!tail) {  // (tail form: the epilogue warps have stored their fp32 columns themselves)
          ptx::tma_store_2d_h<P3_HINT_ACT_STORE>(&map_out2, ptx::smem_u32(my_box) + b * kChBoxBytes, (cur.idx - (cur.it < n_it ? n1_slabs : 0)) * 64,
                            pair_tile_row0(pair + (cur.it - defer) * n_pairs, static_cast<int>(rank), rev_last) + q_row);
        }
        ptx::bulk_commit();
        ptx::bulk_wait_read<1>();  // the previous step's store has read its box: that box serves the step n_boxes - 1 ahead
        if (cur.s > 0) prepare(ahead);
      }
      if (is_epi1) ++g;
      if (cur.s > 0) {
        advance(ahead);
        if (l2pf > 0) {
          if (lane == 0) prefetch(pf);
          advance(pf);
        }
      }
      advance(cur);
      __syncwarp();
    }
    if (lane == 0) ptx::bulk_wait_all();
  } else {
    // ===== epilogue (both CTAs): 4 warps per TMEM lane quarter, thread = one row x 16 of a slab's 64 columns =====
    const int ew = warp - 2;
    const int q = warp & 3;   // TMEM lane quarter of this warp
    const int cg = ew >> 2;   // which 16 of a slab's 64 columns
    // a thread's two 16-byte chunks inside a [rows x 128 B] 128B-swizzled box / K-major slab: row*128 + ((c ^ (row & 7)) << 4)
    const uint32_t sw = static_cast<uint32_t>(lane & 7);
    const uint32_t ch0 = ((2u * cg) ^ sw) << 4, ch1 = ((2u * cg + 1u) ^ sw) << 4;
    const uint32_t box_row = static_cast<uint32_t>(lane) * 128u;
    const uint32_t slab_row = static_cast<uint32_t>(q * 32 + lane) * 128u;
    const uint32_t box_base = ptx::smem_u32(smem_box) + static_cast<uint32_t>(q * n_boxes) * kChBoxBytes + box_row;
    const uint32_t a2_base = ptx::smem_u32(smem_a2);
    uint64_t* my_ready = box_ready + kChMaxBoxes * q;
    uint64_t* my_written = box_written + kChMaxBoxes * q;
    const uint32_t acc1_empty_l = ptx::mapa_shared(ptx::smem_u32(acc1_empty), 0);
    const uint32_t acc2_empty_l = ptx::mapa_shared(ptx::smem_u32(acc2_empty), 0);
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    const uint32_t sc1 = ptx::smem_u32(s_scale1), sh1 = ptx::smem_u32(s_shift1);
    const uint32_t sc2 = ptx::smem_u32(s_scale2), sh2 = ptx::smem_u32(s_shift2);

    uint32_t g = 0;                 // A2-slab ordinal of this quarter
    uint32_t ob = 0, ob_phase = 0;  // staging box of the current step (step ordinal mod n_boxes) and its use parity
    int it = 0;

    // P3_TC_TRACE: per-phase clock64 sums of one epilogue thread (perf experiments)
    const bool tr = kTrace && trace != nullptr && blockIdx.x == 3 && ew == 0 && lane == 0;
    uint32_t ts[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    const long long t_loop = tr ? clock64() : 0;
    for (it = 0; it < n_it + defer; ++it) {
      if (it < n_it) {
        // ---- epi1: x' = acc1 + x ; u = mish(BN(x')).  Padding rows need no masking here: their A1 rows and residual rows
        // are zeros (layout invariant), so x' = 0 exactly; their u rows only reach out rows that epi2 zeroes.
        long long tc0 = tr ? clock64() : 0;
        ptx::mbar_wait(acc1_full, static_cast<uint32_t>(it) & 1u);
        ptx::tc_fence_after_sync();
        if (tr) ts[8] += static_cast<uint32_t>(clock64() - tc0);
        for (int j = 0; j < n1_slabs; ++j, ++g) {
          const int col = j * 64 + cg * 16;
          uint32_t v[16];
          long long t[8];
          if (tr) t[0] = clock64();
          ptx::tmem_ld_32x16(lane_addr + static_cast<uint32_t>(col), v);
          ptx::mbar_wait(&my_ready[ob], ob_phase);
          if (tr) t[1] = clock64();
          const uint32_t obuf = box_base + ob * kChBoxBytes;
          const float4 t0 = ptx::lds_f4(obuf + ch0), t1 = ptx::lds_f4(obuf + ch1);
          float x[16];
          {
            const uint32_t u[8] = {__float_as_uint(t0.x), __float_as_uint(t0.y), __float_as_uint(t0.z), __float_as_uint(t0.w),
                                   __float_as_uint(t1.x), __float_as_uint(t1.y), __float_as_uint(t1.z), __float_as_uint(t1.w)};
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float2 f2 = __half22float2(*reinterpret_cast<const __half2*>(&u[i]));
              x[2 * i] = f2.x;
              x[2 * i + 1] = f2.y;
            }
          }
          ptx::tmem_ld_wait();
          if (tr) t[2] = clock64();
          if (j == n1_slabs - 1) {  // acc1 is in registers: hand it back to the MMA warp
            ptx::tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive_remote(acc1_empty_l);
          }
#pragma unroll
          for (int i = 0; i < 16; ++i) x[i] += __uint_as_float(v[i]);
          const uint4 r0 = make_uint4(tc_pack_f16(x[0], x[1]), tc_pack_f16(x[2], x[3]), tc_pack_f16(x[4], x[5]), tc_pack_f16(x[6], x[7]));
          const uint4 r1 = make_uint4(tc_pack_f16(x[8], x[9]), tc_pack_f16(x[10], x[11]), tc_pack_f16(x[12], x[13]), tc_pack_f16(x[14], x[15]));
          float a[16];
          if (tail) {
#pragma unroll
            for (int i = 0; i < 16; ++i) a[i] = x[i];
          } else {
            bn_mish16(x, a, sc1, sh1, col);
          }
          const uint4 p0 = make_uint4(tc_pack_act(a[0], a[1], f16), tc_pack_act(a[2], a[3], f16), tc_pack_act(a[4], a[5], f16), tc_pack_act(a[6], a[7], f16));
          const uint4 p1 = make_uint4(tc_pack_act(a[8], a[9], f16), tc_pack_act(a[10], a[11], f16), tc_pack_act(a[12], a[13], f16), tc_pack_act(a[14], a[15], f16));
          // A2 slot free: the MMAs that read its previous slab have completed.  x' replaces the residual in place.
          const uint32_t slot = g & 1u;
          if (tr) t[3] = clock64() + (p0.x & 0);
          ptx::mbar_wait(&a2_empty[slot], ((g >> 1) & 1u) ^ 1u);
          if (tr) t[4] = t[5] = clock64();
          if (!tail) {
            ptx::sts_u4(obuf + ch0, r0);
            ptx::sts_u4(obuf + ch1, r1);
          }
          const uint32_t ap = a2_base + slot * kChSlabBytes + slab_row;
          ptx::sts_u4(ap + ch0, p0);
          ptx::sts_u4(ap + ch1, p1);
          ptx::fence_proxy_async();  // generic-proxy smem writes -> visible to the tensor core / TMA engine
          if (tr) t[6] = clock64();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(&my_written[ob]);
          if (++ob == static_cast<uint32_t>(n_boxes)) {
            ob = 0;
            ob_phase ^= 1u;
          }
          if (tr) {
            t[7] = clock64();
            // 0 box ready wait  1 tmem ld wait  2 math  3 a2_empty wait  4 -  5 sts + fence  6 arrive
#pragma unroll
            for (int i = 0; i < 7; ++i) ts[i] += static_cast<uint32_t>(t[i + 1] - t[i]);
            ++ts[10];
          }
        }
      }

      // ---- epi2: out = act2(acc2) of tile it2 (one tile behind epi1 when acc2 is double-buffered, so GEMM2's tail and the
      // commit latency are never waited for)
      const int it2 = it - defer;
      if (it2 >= 0 && it2 < n_it) {
        const int m = pair_tile_row0(pair + it2 * n_pairs, static_cast<int>(rank), rev_last) + q * 32 + lane;
        const bool live = m < rows && row_is_live(m % kRowsPerPos);
        const int as = it2 % acc2_stages;
        const long long tc0 = tr ? clock64() : 0;
        ptx::mbar_wait(&acc2_full[as], static_cast<uint32_t>(it2 / acc2_stages) & 1u);
        ptx::tc_fence_after_sync();
        const long long tc1 = tr ? clock64() : 0;
        if (tr) ts[9] += static_cast<uint32_t>(tc1 - tc0);
        for (int b = 0; b < n2_slabs; ++b) {
          const int col = b * 64 + cg * 16;
          uint32_t v[16];
          ptx::tmem_ld_32x16(lane_addr + static_cast<uint32_t>(n1 + as * n2 + col), v);
          ptx::tmem_ld_wait();
          if (b == n2_slabs - 1) {
            ptx::tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive_remote(acc2_empty_l + 8u * static_cast<uint32_t>(as));
          }
          float x[16], a[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) x[i] = __uint_as_float(v[i]);
          if (act2_mode == kActIdentity) {
#pragma unroll
            for (int i = 0; i < 16; ++i) a[i] = x[i];
          } else {  // kActMishBN (scale, shift) / kActMish (1, 0)
            bn_mish16(x, a, sc2, sh2, col);
          }
          ptx::mbar_wait(&my_ready[ob], ob_phase);
          if (tail) {
            // fp32, channel-major: a warp's 32 rows of one column are 128 contiguous bytes
            if (m < rows) {
              float* op = out_t + static_cast<size_t>(col) * out_ld + m;
#pragma unroll
              for (int i = 0; i < 16; ++i)
                if (col + i < n2_valid) op[static_cast<size_t>(i) * out_ld] = live ? a[i] : 0.0f;
            }
          } else {
            uint4 p0 = make_uint4(tc_pack_act(a[0], a[1], f16), tc_pack_act(a[2], a[3], f16), tc_pack_act(a[4], a[5], f16), tc_pack_act(a[6], a[7], f16));
            uint4 p1 = make_uint4(tc_pack_act(a[8], a[9], f16), tc_pack_act(a[10], a[11], f16), tc_pack_act(a[12], a[13], f16), tc_pack_act(a[14], a[15], f16));
            if (!live) p0 = p1 = make_uint4(0, 0, 0, 0);  // padding rows of the layout stay zero
            const uint32_t obuf = box_base + ob * kChBoxBytes;
            ptx::sts_u4(obuf + ch0, p0);
            ptx::sts_u4(obuf + ch1, p1);
          }
          ptx::fence_proxy_async();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(&my_written[ob]);
          if (++ob == static_cast<uint32_t>(n_boxes)) {
            ob = 0;
            ob_phase ^= 1u;
          }
        }
        if (tr) {
          ts[11] += static_cast<uint32_t>(clock64() - tc1);
          ++ts[12];
        }
      }
    }
    if (tr) {
      for (int i = 0; i < 13; ++i) atomicAdd(&trace[i], static_cast<unsigned long long>(ts[i]));
      atomicAdd(&trace[13], static_cast<unsigned long long>(clock64() - t_loop));
      atomicAdd(&trace[14], 1ull);
    }
  }

  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::cluster_sync_all();  // the peer's MMAs / TMEM traffic are complete before either CTA frees its columns
  if (warp == 1) {
    ptx::tc_fence_after_sync();
    ptx::tmem_dealloc_pair(tmem_base, static_cast<uint32_t>(tmem_cols));
  }
}

}  // namespace

bool tc_chain_supported(int k1, int n1, int n2) {
  if (k1 <= 0 || n1 <= 0 || n2 <= 0) return false;
  if (k1 % 64 || n1 % 64 || n2 % 64) return false;
  if (k1 > 256 || n1 > kChMaxN || n2 > kChMaxN || n1 + n2 > 512) return false;
  // at least a 2-slab A1 ring (one slab if K1 = 64) and 3 boxes per quarter
  return static_cast<size_t>(chain_fixed_smem(k1, n1, n2)) + std::min(2, k1 / 64) * kChSlabBytes + 4 * 3 * kChBoxBytes + 1024 + kChBarBytes <=
         static_cast<size_t>(kChSmemBudget);
}

int tc_chain_plan_create(const __nv_bfloat16* in, const __nv_bfloat16* w1, const __nv_bfloat16* w2, int rows, int k1, int n1,
                         int n2, const void* residual_f16, void* raw_f16, const float* scale1, const float* shift1, void* out2,
                         const float* scale2, const float* shift2, int act2_mode, TcChainPlan** out, bool op_f16, float* tail_out_t,
                         int tail_out_ld, int tail_n2_valid) {
  if (!tc_chain_supported(k1, n1, n2)) return fail(P3_ERR_UNSUPPORTED, "tc_chain: shape not supported");
  if (tail_out_t && (act2_mode != kActIdentity || tail_n2_valid <= 0 || tail_n2_valid > n2 || tail_out_ld < rows))
    return fail(P3_ERR_INVALID_ARG, "tc_chain: tail form needs an identity output and n2_valid <= n2");
  if (!in || !w1 || !w2 || !residual_f16 || !raw_f16 || !scale1 || !shift1 || !out2)
    return fail(P3_ERR_INVALID_ARG, "tc_chain: null argument");
  TcChainPlan* p = new TcChainPlan();
  p->rows = rows;
  p->k1 = k1;
  p->n1 = n1;
  p->n2 = n2;
  p->scale1 = scale1;
  p->shift1 = shift1;
  p->scale2 = scale2;
  p->shift2 = shift2;
  p->act2_mode = act2_mode;
  p->f16 = op_f16 ? 1 : 0;
  p->tail = tail_out_t ? 1 : 0;
  p->out_t = tail_out_t;
  p->out_ld = tail_out_ld;
  p->n2_valid = tail_n2_valid;
  p->acc2_stages = (512 - n1) / n2 >= 2 ? 2 : 1;
  p->tmem_cols = 512;
  {  // split what is left of shared memory between the A1 ring (up to one tile) and the box pool (up to 5 per quarter)
    const size_t left = static_cast<size_t>(kChSmemBudget) - 1024 - kChBarBytes - chain_fixed_smem(k1, n1, n2);
    int stages = std::min(2, k1 / 64), boxes = 3;
    while (true) {
      bool grew = false;
      if (boxes < kChMaxBoxes && static_cast<size_t>(stages) * kChSlabBytes + static_cast<size_t>(4 * (boxes + 1)) * kChBoxBytes <= left) {
        ++boxes;
        grew = true;
      }
      if (stages < std::min(4, k1 / 64) && static_cast<size_t>(stages + 1) * kChSlabBytes + static_cast<size_t>(4 * boxes) * kChBoxBytes <= left) {
        ++stages;
        grew = true;
      }
      if (!grew) break;
    }
    p->a1_stages = stages;
    p->n_boxes = boxes;
    p->l2pf = boxes <= 3 ? 3 : 0;
    if (const char* pf = std::getenv("P3_CHAIN_L2PF")) p->l2pf = boxes <= 3 ? std::max(0, std::atoi(pf)) : 0;
    if (const char* pf = std::getenv("P3_CHAIN_L2PF_ALL")) p->l2pf = boxes <= 3 ? p->l2pf : std::max(0, std::atoi(pf));
    p->smem_bytes = static_cast<size_t>(chain_fixed_smem(k1, n1, n2)) + static_cast<size_t>(stages) * kChSlabBytes +
                    static_cast<size_t>(4 * boxes) * kChBoxBytes + 1024 + kChBarBytes;
  }
  const CUtensorMapDataType bf = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, hf = CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
  const CUtensorMapSwizzle sw = CU_TENSOR_MAP_SWIZZLE_128B;
  int rc = tc_make_map_2d(&p->map_a1, in, bf, 2, k1, rows, 64, 128, sw);
  if (rc == P3_OK) rc = tc_make_map_2d(&p->map_w1, w1, bf, 2, k1, n1, 64, n1 / 2, sw);
  if (rc == P3_OK) rc = tc_make_map_2d(&p->map_w2, w2, bf, 2, n1, n2, 64, n2 / 2, sw);
  if (rc == P3_OK) rc = tc_make_map_2d(&p->map_res, residual_f16, hf, 2, n1, rows, 64, 32, sw);
  if (rc == P3_OK) rc = tc_make_map_2d(&p->map_raw, raw_f16, hf, 2, n1, rows, 64, 32, sw);
  if (rc == P3_OK) rc = tc_make_map_2d(&p->map_out2, out2, bf, 2, n2, rows, 64, 32, sw);
  if (rc == P3_OK) {
    cudaError_t e = cudaFuncSetAttribute(tc_chain_pair_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kChSmemBudget);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(tc_chain_pair_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kChSmemBudget);
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
    cfg.blockDim = dim3(kChThreads);
    cfg.dynamicSmemBytes = p->smem_bytes;
    int n_clusters = 0;
    if (cudaOccupancyMaxActiveClusters(&n_clusters, tc_chain_pair_kernel<false>, &cfg) == cudaSuccess && n_clusters > 0)
      max_pairs = std::min(max_pairs, n_clusters);
    else
      cudaGetLastError();
  }
  const int m_tiles = pair_tile_count(rows);
  p->grid = 2 * std::max(1, std::min(max_pairs, m_tiles));
  if (std::getenv("P3_TC_TRACE")) {
    cudaMalloc(&p->trace, 16 * sizeof(unsigned long long));
    cudaMemset(p->trace, 0, 16 * sizeof(unsigned long long));
  }
  *out = p;
  return P3_OK;
}

void tc_chain_plan_set_reverse(TcChainPlan* p, bool reverse) {
  if (p) p->reverse = reverse ? 1 : 0;
}

void tc_chain_plan_destroy(TcChainPlan* p) {
  if (p && p->trace) {
    unsigned long long h[16];
    cudaMemcpy(h, p->trace, sizeof h, cudaMemcpyDeviceToHost);
    if (h[10] && h[12] && h[14])
      std::fprintf(stderr, "[p3 trace] chain %d->%d->%d epi1 cycles/step: box_ready %llu  tmem_ld %llu  math %llu  a2_empty %llu  (-) %llu  "
                           "sts+fence %llu  arrive %llu  (-) %llu | per tile: acc1 wait %llu  acc2 wait %llu  epi2 %llu | "
                           "per launch: loop %llu cycles, %llu tiles\n",
                   p->k1, p->n1, p->n2, h[0] / h[10], h[1] / h[10], h[2] / h[10], h[3] / h[10], h[4] / h[10], h[5] / h[10],
                   h[6] / h[10], h[7] / h[10], h[8] / h[12], h[9] / h[12], h[11] / h[12], h[13] / h[14], h[12] / h[14]);
    cudaFree(p->trace);
  }
  delete p;
}

int tc_chain_launch(const TcChainPlan* p, cudaStream_t stream) {
  auto kern = p->trace ? tc_chain_pair_kernel<true> : tc_chain_pair_kernel<false>;
  P3_CUDA(tc_launch_pdl(kern, p->grid, kChThreads, p->smem_bytes, stream, p->map_a1, p->map_w1, p->map_w2, p->map_res, p->map_raw,
                        p->map_out2, p->rows, p->k1, p->n1, p->n2, p->acc2_stages, p->a1_stages, p->n_boxes, p->tmem_cols, p->scale1, p->shift1, p->scale2,
                        p->shift2, p->act2_mode, p->trace, p->f16, p->tail, p->out_t, p->out_ld, p->n2_valid, p->reverse, p->l2pf));
  return P3_OK;
}

}  // namespace p3
