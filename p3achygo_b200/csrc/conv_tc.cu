// tcgen05 / TMEM / TMA implicit-GEMM convolution for the residual tower (P3_PRECISION_BF16).
//
// conv(mish(BN(x))) of python/model.py:276-281 over the padded board-row layout (common.cuh): the
// activated bf16 input is a 2-D matrix [rows, cin]; a k x k "same" conv is `taps` row-shifted GEMMs
//     acc[m, n] = sum_t sum_k  A[m + tap_off[t], k] * W[t][n][k]
// whose zero padding is the layout's zero rows plus TMA's out-of-bounds zero fill (negative and
// past-the-end row coordinates).  No im2col matrix is materialised.
//
// Two warp-specialised persistent kernels (one CTA per SM, 320 threads):
//   warp 0      TMA producer (one elected thread)
//   warp 1      MMA issuer (one elected thread; also owns the TMEM allocation): tcgen05.mma M=128, K=16 into
//               one of two fp32 accumulators in TMEM; tcgen05.commit releases smem slots / publishes accumulators
//   warps 2-9   epilogue, two warps per TMEM lane quarter: tcgen05.ld -> (+ residual) -> raw fp32 stream
//               and/or bf16 activated copy with the NEXT layer's BN + mish folded in (ConvEpilogue)
//
//   tc_conv_kernel          generic (1x1 layers, head conv, 3x3 layers whose weights do not fit on chip):
//                           A and W slabs streamed through a smem ring; the epilogue's global I/O is ALL TMA:
//                           residual tiles are bulk-loaded into smem ahead of use and outputs leave through
//                           swizzled staging tiles + cp.async.bulk.tensor stores, because these layers are
//                           HBM-bound and a few warps of ld/st.global cannot keep enough bytes in flight
//                           (measured: 1.2 TB/s with per-thread accesses).
//   tc_conv3x3_res_kernel   3x3 layers with resident weights and tap reuse (see below).
#include <cuda.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <type_traits>
#include <vector>

#include "common.cuh"
#include "math.cuh"
#include "ptx.cuh"
#include "tc_util.cuh"

namespace p3 {

constexpr int kTileM = 128;
constexpr int kSlabK = 64;  // bf16 elements per 128-byte swizzled row
constexpr int kUmmaK = 16;
constexpr int kMaxTapsTc = 9;
constexpr int kABytes = kTileM * kSlabK * 2;  // 16 KB
// BN + mish in the epilogue is latency-bound unless every scheduler interleaves several warps (ablation on B200):
// 16 epilogue warps = 4 per scheduler
constexpr int kNumEpiWarps = 16;
constexpr int kNumEpiThreads = kNumEpiWarps * 32;
constexpr int kNumThreads = 64 + kNumEpiThreads;  // 576
constexpr int kMaxCout = 512;
constexpr int kSmemBudget = 224 * 1024;
constexpr int kBarBytes = 512 + 2 * kMaxCout * 4;  // barriers + tmem ptr, then folded-BN scale / shift of this layer
constexpr int kMaxNTile = 128;
// generic kernel staging tiles: 128 rows x 32 columns.  Residual tiles are prefetched kNumResBuf chunks ahead (TMA
// load latency ~1.5 us vs ~1 us of HBM time per chunk); output tiles are double-buffered so the epilogue never waits
// for a bulk store to drain (measured: single-buffered staging cost ~1.5 us per chunk).
constexpr int kStageF32Bytes = kTileM * 32 * 4;   // 16 KB
constexpr int kStageBf16Bytes = kTileM * 32 * 2;  // 8 KB
constexpr int kNumResBuf = 4;
constexpr int kNumOutBuf = 2;
// residual / raw tiles are fp32 (128 B rows, 128B swizzle) or, when the residual stream is fp16, 64 B rows (64B swizzle)
__host__ __device__ constexpr int staging_bytes(bool res, bool raw, bool act, bool raw_f16) {
  return (res ? kNumResBuf * (raw_f16 ? kStageBf16Bytes : kStageF32Bytes) : 0) +
         (raw ? kNumOutBuf * (raw_f16 ? kStageBf16Bytes : kStageF32Bytes) : 0) + (act ? kNumOutBuf * kStageBf16Bytes : 0);
}

struct TcTaps {
  int off[kMaxTapsTc];
};

struct TcConvPlan {
  CUtensorMap map_a, map_w, map_res, map_raw, map_act;
  int rows, cin, cout, taps, n_tile, stages, tmem_cols, grid;
  size_t smem_bytes;
  TcTaps tap;
  ConvEpilogue ep;
  bool resident = false;  // tc_conv3x3_res_kernel
  bool pair = false;      // tc_conv3x3_pair_kernel (CTA pairs, cta_group::2)
  int n_half = 0;         // pair kernel: output channels held per CTA
  bool staged = false;    // pair kernel: output leaves through per-warp smem staging + TMA stores (map_raw = the box map)
  int reverse = 0;        // pair kernel: tile order (common.cuh, pair_tile_row0)
  int debug = 0;          // P3_TC_DEBUG ablation bits (perf experiments only; results are wrong when set)
  unsigned long long* trace = nullptr;  // P3_TC_TRACE: per-phase clock64 sums of one epilogue leader (perf experiments)
};

namespace {

__device__ __forceinline__ uint32_t pack_bf16_only(float a, float b) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}
__device__ __forceinline__ uint32_t pack_f16(float a, float b);
// activated operand pair in the engine's operand format (bf16, or IEEE fp16 for P3_PRECISION_FP16)
__device__ __forceinline__ uint32_t pack_act(float a, float b, int f16) { return f16 ? pack_f16(a, b) : pack_bf16_only(a, b); }

// two floats -> packed IEEE fp16 pair, saturating to +-65504 (the residual stream never overflows to inf)
__device__ __forceinline__ uint32_t pack_f16(float a, float b) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
  return r;
}

template <int kMode>
__device__ __forceinline__ float activate(float x, float sc, float sh) {
  if (kMode == kActMishBN) return mish_f32<false>(fmaf(x, sc, sh));
  if (kMode == kActMish) return mish_f32<false>(x);
  return x;
}

// ===================================================================================================
// generic kernel
// ===================================================================================================
__global__ void __launch_bounds__(kNumThreads, 1)
tc_conv_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w,
               const __grid_constant__ CUtensorMap map_res, const __grid_constant__ CUtensorMap map_raw,
               const __grid_constant__ CUtensorMap map_act, int rows, int cin, int cout, int taps, TcTaps tap, int n_tile,
               int stages, int tmem_cols, int has_res, int has_raw, int has_act, const float* __restrict__ scale,
               const float* __restrict__ shift, int act_mode, int raw_f16, int raw_t, int debug, unsigned long long* trace, int f16) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // carve: [staging: residual x4 | raw x2 | act x2, each only if used] | [ring: stages x (A 16 KB | B n_tile*128 B)] | barriers
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int raw_tile_bytes = raw_f16 ? kStageBf16Bytes : kStageF32Bytes;
  uint8_t* st_res = smem;
  uint8_t* st_raw = st_res + (has_res ? kNumResBuf * raw_tile_bytes : 0);
  uint8_t* st_act = st_raw + (has_raw ? kNumOutBuf * raw_tile_bytes : 0);
  uint8_t* ring = smem + staging_bytes(has_res, has_raw, has_act, raw_f16);
  const int stage_bytes = kABytes + n_tile * kSlabK * 2;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(ring + static_cast<size_t>(stages) * stage_bytes);
  uint64_t* empty_bar = full_bar + stages;
  uint64_t* tmem_full = empty_bar + stages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint64_t* res_full = tmem_empty + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(res_full + kNumResBuf);
  float* s_scale = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(full_bar) + 512);
  float* s_shift = s_scale + kMaxCout;
  for (int c = threadIdx.x; c < cout; c += blockDim.x) {
    s_scale[c] = act_mode == kActMishBN ? scale[c] : 1.0f;
    s_shift[c] = act_mode == kActMishBN ? shift[c] : 0.0f;
  }

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m_tiles = (rows + kTileM - 1) / kTileM;
  const int n_tiles = cout / n_tile;
  const int total_tiles = m_tiles * n_tiles;
  const int k_slabs = cin / kSlabK;
  const int k_steps = taps * k_slabs;
  const int n_chunks = n_tile / 32;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&map_a);
    ptx::prefetch_tensormap(&map_w);
    for (int s = 0; s < stages; ++s) {
      ptx::mbar_init(&full_bar[s], 1);
      ptx::mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      ptx::mbar_init(&tmem_full[s], 1);
      ptx::mbar_init(&tmem_empty[s], kNumEpiWarps);
    }
    for (int s = 0; s < kNumResBuf; ++s) ptx::mbar_init(&res_full[s], 1);
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
    // ===== TMA producer: the whole warp runs the loop (uniform control flow), one elected lane issues =====
    {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int m0 = (tile / n_tiles) * kTileM, n0 = (tile % n_tiles) * n_tile;
        for (int t = 0; t < taps; ++t) {
          for (int ks = 0; ks < k_slabs; ++ks) {
            ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
            if (ptx::elect_one()) {
              uint8_t* sa = ring + static_cast<size_t>(stage) * stage_bytes;
              if (debug & 2) {
                ptx::mbar_arrive(&full_bar[stage]);  // ablation: no A / W traffic
              } else {
                ptx::mbar_arrive_expect_tx(&full_bar[stage], static_cast<uint32_t>(stage_bytes));
                ptx::tma_load_2d(sa, &map_a, &full_bar[stage], ks * kSlabK, m0 + tap.off[t]);
                ptx::tma_load_2d(sa + kABytes, &map_w, &full_bar[stage], ks * kSlabK, t * cout + n0);
              }
            }
            __syncwarp();
            if (++stage == stages) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: warp-converged loop, tcgen05.mma / commit issued by one elected lane =====
    {
      const uint32_t idesc = ptx::make_idesc_op(kTileM, n_tile, f16);
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
          const uint32_t a_lo = ptx::desc_lo_sw128(sa), b_lo = ptx::desc_lo_sw128(sa + kABytes);
          if (ptx::elect_one()) {
#pragma unroll
            for (int k = 0; k < kSlabK / kUmmaK; ++k)  // +32 bytes (2 x 16 B) along K inside the 128-byte swizzled row
              if (!(debug & 4))
                ptx::umma_f16_lohi(tmem_d, a_lo + 2 * k, ptx::desc_hi_sw128(), b_lo + 2 * k, ptx::desc_hi_sw128(), idesc,
                                   (step > 0 || k > 0) ? 1u : 0u);
            ptx::umma_commit(&empty_bar[stage]);  // ring slot reusable once these MMAs have read it
          }
          __syncwarp();
          if (++stage == stages) {
            stage = 0;
            phase ^= 1;
          }
        }
        if (ptx::elect_one()) ptx::umma_commit(&tmem_full[acc]);  // accumulator complete
        __syncwarp();
      }
    }
  } else {
    // ===== epilogue: 16 warps in TWO independent groups of 8.  The chunks (32 output columns) of this CTA's tiles are
    // numbered G = 0, 1, 2, ...; group `grp` handles the chunks with G % 2 == grp, so two chunks are always in flight
    // (one chunk is a serial chain: tcgen05.ld -> residual -> BN+mish -> staging -> fence -> TMA store; measured
    // ~1.5 us end to end, which bounded the 1x1 layers when all 16 warps marched through it in lock step).
    // Within a group: 2 warps per TMEM lane quarter, 16 columns each; thread = one output row.  All global traffic
    // goes through TMA; each group owns its staging tiles, named barriers, residual ring and bulk-store groups. =====
    const int ew = warp - 2;
    const int grp = ew >> 3;              // epilogue group 0 / 1
    const int quarter = warp & 3;         // TMEM lane quarter this warp may access
    const int half = (ew >> 2) & 1;       // which 16 columns of the 32-column chunk
    const int r = quarter * 32 + lane;    // row within the tile
    const bool leader = (ew & 7) == 0 && lane == 0;
    const int bar_a = 1 + 2 * grp, bar_b = 2 + 2 * grp;
    constexpr int kGrpThreads = 256;
    // this thread's row inside a 128 B-row (fp32, 128B swizzle) / 64 B-row (bf16, 64B swizzle) staging tile
    const uint32_t f32_row = static_cast<uint32_t>(r) * 128u;
    const uint32_t bf_row = static_cast<uint32_t>(r) * 64u;
    uint8_t* my_res = st_res + grp * (2 * raw_tile_bytes);   // 2-deep residual ring per group
    uint8_t* my_raw = st_raw + grp * raw_tile_bytes;
    uint8_t* my_act = st_act + grp * kStageBf16Bytes;
    uint64_t* my_res_full = res_full + 2 * grp;

    // residual tile of global chunk G -> ring slot (G/2) & 1 of this group
    auto issue_res = [&](uint32_t G) {
      const int tile_i = blockIdx.x + static_cast<int>(G / n_chunks) * gridDim.x;
      if (tile_i >= total_tiles) return;
      const int ci = static_cast<int>(G % n_chunks);
      const int m0i = (tile_i / n_tiles) * kTileM, n0i = (tile_i % n_tiles) * n_tile;
      const uint32_t b = (G >> 1) & 1u;
      ptx::mbar_arrive_expect_tx(&my_res_full[b], static_cast<uint32_t>(raw_tile_bytes));
      ptx::tma_load_2d(my_res + b * raw_tile_bytes, &map_res, &my_res_full[b], n0i + ci * 32, m0i);
    };
    if (leader && has_res) {  // prefetch this group's first two chunks
      issue_res(grp);
      issue_res(grp + 2);
    }

    int iter = 0;
    const long long t_loop = (trace != nullptr && blockIdx.x == 3 && leader && grp == 0) ? clock64() : 0;
    uint32_t tsum[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};  // P3_TC_TRACE: per-phase cycle sums of this thread (registers)
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++iter) {
      const int m0 = (tile / n_tiles) * kTileM, n0 = (tile % n_tiles) * n_tile;
      const int acc = iter & 1;
      const uint32_t acc_phase = (iter >> 1) & 1;
      const int m = m0 + r;
      const bool live = m < rows && row_is_live(m % kRowsPerPos);
      bool waited = false;

      for (int c = 0; c < ((debug & 1) ? 0 : n_chunks); ++c) {
        const uint32_t G = static_cast<uint32_t>(iter) * n_chunks + c;
        if ((G & 1u) != static_cast<uint32_t>(grp)) continue;
        const bool tr = trace != nullptr && blockIdx.x == 3 && leader && grp == 0;
        if (!waited) {
          const long long tw = tr ? clock64() : 0;
          ptx::mbar_wait(&tmem_full[acc], acc_phase);
          ptx::tc_fence_after_sync();
          waited = true;
          if (tr) tsum[8] += static_cast<uint32_t>(clock64() - tw);
        }
        const uint32_t k = G >> 1;  // this group's chunk ordinal
        const uint32_t buf = k & 1u;
        long long tc[8];
        if (tr) tc[0] = clock64();
        uint32_t v[16];
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) +
                               static_cast<uint32_t>(acc * n_tile + c * 32 + half * 16);
        ptx::tmem_ld_32x16(taddr, v);
        float x[16];
        if (has_res) {
          ptx::mbar_wait(&my_res_full[buf], (k >> 1) & 1u);
          if (raw_f16) {  // 16 halves = 2 chunks of the 64 B row (64B swizzle)
            const uint32_t rp = ptx::smem_u32(my_res) + buf * raw_tile_bytes + bf_row;
#pragma unroll
            for (int j = 0; j < 2; ++j) {
              const float4 t4 = ptx::lds_f4(rp + (((half * 2 + j) ^ ((r >> 1) & 3)) << 4));
              const uint32_t u[4] = {__float_as_uint(t4.x), __float_as_uint(t4.y), __float_as_uint(t4.z), __float_as_uint(t4.w)};
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const float2 f2 = __half22float2(*reinterpret_cast<const __half2*>(&u[q]));
                x[8 * j + 2 * q] = f2.x;
                x[8 * j + 2 * q + 1] = f2.y;
              }
            }
          } else {
            const uint32_t rp = ptx::smem_u32(my_res) + buf * raw_tile_bytes + f32_row;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float4 t4 = ptx::lds_f4(rp + (((half * 4 + j) ^ (r & 7)) << 4));
              x[4 * j] = t4.x;
              x[4 * j + 1] = t4.y;
              x[4 * j + 2] = t4.z;
              x[4 * j + 3] = t4.w;
            }
          }
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) x[j] = 0.0f;
        }
        if (tr) tc[1] = clock64();
        ptx::tmem_ld_wait();
        if (tr) tc[2] = clock64();
#pragma unroll
        for (int j = 0; j < 16; ++j) x[j] = live ? (__uint_as_float(v[j]) + x[j]) : 0.0f;
        uint4 pk0 = make_uint4(0, 0, 0, 0), pk1 = pk0;
        if (has_act) {
          const int nb = n0 + c * 32 + half * 16;
          float a[16];
          if (act_mode == kActIdentity || (debug & 16)) {
#pragma unroll
            for (int j = 0; j < 16; ++j) a[j] = x[j];
          } else {  // kActMishBN (scale, shift) / kActMish (1, 0)
            const uint32_t sca = ptx::smem_u32(s_scale + nb), sha = ptx::smem_u32(s_shift + nb);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float4 s4 = ptx::lds_f4_const(sca + 16 * j), h4 = ptx::lds_f4_const(sha + 16 * j);
              a[4 * j] = mish_f32<false>(fmaf(x[4 * j], s4.x, h4.x));
              a[4 * j + 1] = mish_f32<false>(fmaf(x[4 * j + 1], s4.y, h4.y));
              a[4 * j + 2] = mish_f32<false>(fmaf(x[4 * j + 2], s4.z, h4.z));
              a[4 * j + 3] = mish_f32<false>(fmaf(x[4 * j + 3], s4.w, h4.w));
            }
#pragma unroll
            for (int j = 0; j < 16; ++j) a[j] = live ? a[j] : 0.0f;
          }
          pk0 = make_uint4(pack_act(a[0], a[1], f16), pack_act(a[2], a[3], f16), pack_act(a[4], a[5], f16), pack_act(a[6], a[7], f16));
          pk1 = make_uint4(pack_act(a[8], a[9], f16), pack_act(a[10], a[11], f16), pack_act(a[12], a[13], f16), pack_act(a[14], a[15], f16));
        }

        // group staging free? (the group's previous bulk stores have read it) and everyone is done with my_res[buf]
        if (tr) tc[3] = clock64() + (pk0.x & 0);
        if (leader) ptx::bulk_wait_read<0>();
        if (tr) tc[4] = clock64();
        ptx::named_bar_sync(bar_a, kGrpThreads);
        if (tr) tc[5] = clock64();
        if (leader && has_res) issue_res(G + 4);  // refill the ring slot just consumed with this group's chunk k + 2
        if (has_raw && raw_f16) {
          const uint32_t wp = ptx::smem_u32(my_raw) + bf_row;
#pragma unroll
          for (int j = 0; j < 2; ++j)
            ptx::sts_u4(wp + (((half * 2 + j) ^ ((r >> 1) & 3)) << 4),
                        make_uint4(pack_f16(x[8 * j], x[8 * j + 1]), pack_f16(x[8 * j + 2], x[8 * j + 3]),
                                   pack_f16(x[8 * j + 4], x[8 * j + 5]), pack_f16(x[8 * j + 6], x[8 * j + 7])));
        } else if (has_raw && raw_t) {  // channel-major output: staging tile [32 columns][128 rows] fp32, plain layout
          const uint32_t wp = ptx::smem_u32(my_raw) + static_cast<uint32_t>(half * 16 * kTileM + r) * 4u;
#pragma unroll
          for (int j = 0; j < 16; ++j)
            asm volatile("st.shared.f32 [%0], %1;" ::"r"(wp + static_cast<uint32_t>(j * kTileM) * 4u), "f"(x[j]) : "memory");
        } else if (has_raw) {
          const uint32_t wp = ptx::smem_u32(my_raw) + f32_row;
#pragma unroll
          for (int j = 0; j < 4; ++j)
            ptx::sts_f4(wp + (((half * 4 + j) ^ (r & 7)) << 4), make_float4(x[4 * j], x[4 * j + 1], x[4 * j + 2], x[4 * j + 3]));
        }
        if (has_act) {
          const uint32_t wp = ptx::smem_u32(my_act) + bf_row;
          ptx::sts_u4(wp + (((half * 2 + 0) ^ ((r >> 1) & 3)) << 4), pk0);
          ptx::sts_u4(wp + (((half * 2 + 1) ^ ((r >> 1) & 3)) << 4), pk1);
        }
        ptx::fence_proxy_async();  // make the generic-proxy smem writes visible to the TMA engine
        if (tr) tc[6] = clock64();
        ptx::named_bar_sync(bar_b, kGrpThreads);
        if (leader && !(debug & 8)) {
          if (has_raw && raw_t) ptx::tma_store_2d(&map_raw, my_raw, m0, n0 + c * 32);
          else if (has_raw) ptx::tma_store_2d(&map_raw, my_raw, n0 + c * 32, m0);
          if (has_act) ptx::tma_store_2d(&map_act, my_act, n0 + c * 32, m0);
          ptx::bulk_commit();
        }
        if (tr) {
          tc[7] = clock64();
          // phases: 0 res-wait  1 tmem ld wait  2 math  3 bulk_wait_read  4 barrier A  5 staging+fence  6 barrier B + TMA issue
#pragma unroll
          for (int i = 0; i < 7; ++i) tsum[i] += static_cast<uint32_t>(tc[i + 1] - tc[i]);
          ++tsum[7];
        }
      }
      if (!waited) {  // (only with the ablation that skips all chunks) still consume the accumulator
        ptx::mbar_wait(&tmem_full[acc], acc_phase);
        ptx::tc_fence_after_sync();
      }
      // this warp is done with the accumulator -> hand it back to the MMA warp (16 arrivals)
      ptx::tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&tmem_empty[acc]);
    }
    if (leader) ptx::bulk_wait_all();
    if (trace != nullptr && blockIdx.x == 3 && leader && grp == 0) {
      atomicAdd(&trace[11], static_cast<unsigned long long>(clock64() - t_loop));
      atomicAdd(&trace[12], 1ull);
      for (int i = 0; i < 8; ++i) atomicAdd(&trace[i], static_cast<unsigned long long>(tsum[i]));
      atomicAdd(&trace[10], static_cast<unsigned long long>(tsum[8]));
    }
  }

  ptx::tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after_sync();
    ptx::tmem_dealloc(tmem_base, static_cast<uint32_t>(tmem_cols));
  }
}

// ===================================================================================================
// 3x3 layers with RESIDENT weights and tap reuse (the dominant kernel of the tower).
//
// ncu on the streaming kernel showed the 3x3 layers bound by L2->SM traffic: every 128-row tile re-read its
// A rows 9 times (once per tap) and streamed all 9*cin*cout weights again.  Here
//   * each CTA owns a 64-wide slice of the output channels and keeps that slice of ALL taps' weights in
//     shared memory for the whole launch (9 * cin * 64 bf16 <= 144 KB), loaded once by TMA;
//   * per 64-channel slab ONE haloed A box of 176 rows (128 + 2*21 halo rows, padded to a multiple of 8)
//     is loaded, and the 9 taps are 9 shifted views of it: the smem matrix descriptor's start address moves
//     by (21 + dy*20 + dx) rows of 128 B inside the 128B-swizzled tile (the swizzle is a function of the
//     absolute smem address, so a row-shifted start needs no base-offset correction — verified on B200).
// L2->SM traffic per 128 output rows drops from 2 * 9 * cin * 256 B to (176/128) * cin * 256 B.
// ===================================================================================================
constexpr int kResN = 64;
constexpr int kResHalo = 21;
constexpr int kResRows = 176;
constexpr int kResABytes = kResRows * 128;   // 22 528 B, a multiple of 1024
constexpr int kResWSlabBytes = kResN * 128;  // 8 KB per (tap, 64-channel slab)
constexpr int kResStages = 3;
// the epilogue is BN + mish on 8192 elements per tile: latency-bound unless every scheduler has several warps to
// interleave (ablation on B200: math alone took 1.8x the MMA time with 8 epilogue warps) -> 16 warps, 4 per scheduler
constexpr int kResEpiWarps = 16;
constexpr int kResThreads = 64 + kResEpiWarps * 32;  // 576

__global__ void __launch_bounds__(kResThreads, 1)
tc_conv3x3_res_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w, int rows,
                      int cin, int cout, TcTaps tap, __nv_bfloat16* __restrict__ act_out,
                      const float* __restrict__ scale, const float* __restrict__ shift, int act_mode, int debug, int f16) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int k_slabs = cin / kSlabK;
  const int w_bytes = 9 * k_slabs * kResWSlabBytes;
  uint8_t* smem_w = smem;
  uint8_t* smem_a = smem + w_bytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem_a + kResStages * kResABytes);
  uint64_t* empty_bar = full_bar + kResStages;
  uint64_t* tmem_full = empty_bar + kResStages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint64_t* w_bar = tmem_empty + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(w_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_slices = cout / kResN;
  const int slice = blockIdx.x % n_slices;
  const int n0 = slice * kResN;
  const int cta_in_slice = blockIdx.x / n_slices, ctas_per_slice = gridDim.x / n_slices;
  const int m_tiles = (rows + kTileM - 1) / kTileM;
  constexpr int kTmemCols = 2 * kResN;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&map_a);
    ptx::prefetch_tensormap(&map_w);
    for (int s = 0; s < kResStages; ++s) {
      ptx::mbar_init(&full_bar[s], 1);
      ptx::mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      ptx::mbar_init(&tmem_full[s], 1);
      ptx::mbar_init(&tmem_empty[s], kResEpiWarps);
    }
    ptx::mbar_init(w_bar, 1);
    ptx::fence_mbar_init();
  } else if (warp == 1) {
    ptx::tmem_alloc(tmem_ptr, kTmemCols);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    {
      // resident weights: all taps x slabs of this CTA's 64 output channels, once
      if (ptx::elect_one()) {
        ptx::mbar_arrive_expect_tx(w_bar, static_cast<uint32_t>(w_bytes));
        for (int t = 0; t < 9; ++t)
          for (int ks = 0; ks < k_slabs; ++ks)
            ptx::tma_load_2d(smem_w + (t * k_slabs + ks) * kResWSlabBytes, &map_w, w_bar, ks * kSlabK, t * cout + n0);
      }
      __syncwarp();
      int stage = 0;
      uint32_t phase = 0;
      for (int mt = cta_in_slice; mt < m_tiles; mt += ctas_per_slice) {
        const int m0 = mt * kTileM;
        for (int ks = 0; ks < k_slabs; ++ks) {
          ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
          if (ptx::elect_one()) {
            if (debug & 2) {
              ptx::mbar_arrive(&full_bar[stage]);  // ablation: no A traffic
            } else {
              ptx::mbar_arrive_expect_tx(&full_bar[stage], kResABytes);
              ptx::tma_load_2d(smem_a + stage * kResABytes, &map_a, &full_bar[stage], ks * kSlabK, m0 - kResHalo);
            }
          }
          __syncwarp();
          if (++stage == kResStages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    {
      const uint32_t idesc = ptx::make_idesc_op(kTileM, kResN, f16);
      const uint32_t w_lo_base = ptx::desc_lo_sw128(ptx::smem_u32(smem_w));
      const uint32_t w_tap_stride = static_cast<uint32_t>(k_slabs) * (kResWSlabBytes / 16);
      uint32_t tap_lo[9];
#pragma unroll
      for (int t = 0; t < 9; ++t) tap_lo[t] = static_cast<uint32_t>(kResHalo + tap.off[t]) * 8u;
      ptx::mbar_wait(w_bar, 0);
      ptx::tc_fence_after_sync();
      int stage = 0;
      uint32_t phase = 0;
      int iter = 0;
      for (int mt = cta_in_slice; mt < m_tiles; mt += ctas_per_slice, ++iter) {
        const int acc = iter & 1;
        const uint32_t acc_phase = (iter >> 1) & 1;
        ptx::mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        ptx::tc_fence_after_sync();
        const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(acc * kResN);
        for (int ks = 0; ks < k_slabs; ++ks) {
          ptx::mbar_wait(&full_bar[stage], phase);
          ptx::tc_fence_after_sync();
          const uint32_t a_lo = ptx::desc_lo_sw128(ptx::smem_u32(smem_a + stage * kResABytes));
          const uint32_t w_lo = w_lo_base + static_cast<uint32_t>(ks) * (kResWSlabBytes / 16);
          if (ptx::elect_one()) {
#pragma unroll
          for (int t = 0; t < 9; ++t) {
            if (debug & 4) continue;  // ablation: no MMAs
            // tap (dy,dx) = shifted view: skip (21 + dy*20 + dx) rows of 128 B (8 x 16 B) inside the swizzled tile;
            // weights of (tap t, slab ks) sit (t*k_slabs + ks) * 8 KB into the resident block
            const uint32_t a_t = a_lo + tap_lo[t];
            const uint32_t w_t = w_lo + static_cast<uint32_t>(t) * w_tap_stride;
#pragma unroll
            for (int k = 0; k < kSlabK / kUmmaK; ++k)
              ptx::umma_f16_lohi(tmem_d, a_t + 2 * k, ptx::desc_hi_sw128(), w_t + 2 * k, ptx::desc_hi_sw128(), idesc,
                                 (ks > 0 || t > 0 || k > 0) ? 1u : 0u);
          }
          ptx::umma_commit(&empty_bar[stage]);
          }
          __syncwarp();
          if (++stage == kResStages) {
            stage = 0;
            phase ^= 1;
          }
        }
        if (ptx::elect_one()) ptx::umma_commit(&tmem_full[acc]);
        __syncwarp();
      }
    }
  } else {
    // epilogue: 16 warps (4 per TMEM lane quarter), thread = one output row x 16 of the slice's 64 columns; writes only
    // the bf16 activated copy (16 KB per tile).  The folded BN of the 16 columns lives in registers for the whole launch.
    const int ew = warp - 2;
    const int quarter = warp & 3;
    const int cg = ew >> 2;  // column group 0..3
    const int nb = n0 + cg * 16;
    float sc[16], sh[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      sc[i] = act_mode == kActMishBN ? __ldg(scale + nb + i) : 1.0f;
      sh[i] = act_mode == kActMishBN ? __ldg(shift + nb + i) : 0.0f;
    }
    int iter = 0;
    for (int mt = cta_in_slice; mt < m_tiles; mt += ctas_per_slice, ++iter) {
      const int acc = iter & 1;
      const uint32_t acc_phase = (iter >> 1) & 1;
      ptx::mbar_wait(&tmem_full[acc], acc_phase);
      ptx::tc_fence_after_sync();
      const int m = mt * kTileM + quarter * 32 + lane;
      const bool in_range = m < rows;
      const bool live = in_range && row_is_live(m % kRowsPerPos);
      uint32_t v[16];
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) +
                             static_cast<uint32_t>(acc * kResN + cg * 16);
      ptx::tmem_ld_32x16(taddr, v);
      ptx::tmem_ld_wait();
      // accumulator values are in registers: release the TMEM stage before the math
      ptx::tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&tmem_empty[acc]);
      if (in_range && !(debug & 1)) {  // ablation bit 0: no epilogue math / stores
        uint32_t packed[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float a0 = __uint_as_float(v[2 * i]), a1 = __uint_as_float(v[2 * i + 1]);
          if (debug & 16) {
            // ablation: no activation math
          } else if (act_mode == kActIdentity) {
          } else {  // kActMishBN (scale/shift) or kActMish (1, 0)
            a0 = mish_f32<false>(fmaf(a0, sc[2 * i], sh[2 * i]));
            a1 = mish_f32<false>(fmaf(a1, sc[2 * i + 1], sh[2 * i + 1]));
          }
          packed[i] = pack_act(live ? a0 : 0.0f, live ? a1 : 0.0f, f16);
        }
        uint4* ap = reinterpret_cast<uint4*>(act_out + static_cast<size_t>(m) * cout + nb);
        if (!(debug & 8) || packed[0] == 0x12345678u) {  // ablation bit 3: no stores
          ap[0] = make_uint4(packed[0], packed[1], packed[2], packed[3]);
          ap[1] = make_uint4(packed[4], packed[5], packed[6], packed[7]);
        }
      }
    }
  }

  ptx::tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after_sync();
    ptx::tmem_dealloc(tmem_base, kTmemCols);
  }
}

// ===================================================================================================
// 3x3 layers on a CTA PAIR (cta_group::2): the same resident-weight / shifted-view scheme, but one tcgen05.mma covers
// M = 256 rows (128 per CTA) x N = 2 * n_half output channels, with each CTA holding only ITS half of the weights' N
// rows.  Per MMA each SM now reads A 4 KB + B n_half*32 B from its shared memory for twice the math of the single-CTA
// N = 64 kernel above, which was capped by shared-memory operand reads (ablation: MMA pipeline alone 1.06 PFLOP/s).
//   rank 0 (leader): warp 1 issues every MMA and the multicast commits; its full / tmem_empty / weight barriers collect
//                    the TMA bytes and the epilogue releases of BOTH CTAs
//   both ranks     : warp 0 TMA-loads its own 176-row haloed A box and its own weight half (cp.async.bulk.tensor
//                    .cta_group::2, completion counted on the leader's barrier); 16 epilogue warps drain the CTA's own
//                    128 x N accumulator (4 warps per TMEM lane quarter, N/4 columns each)
// ===================================================================================================
constexpr int kPairThreads = 64 + 16 * 32;  // 576
__device__ unsigned long long g_last_kernel_end = 0;  // P3_TC_TRACE only
__device__ __forceinline__ unsigned long long global_timer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
constexpr int kPairMaxN = 128;

// kRes: the layer closes a residual block at this width (classic blocks, the inner blocks of nested-bottleneck nets):
// `res` (fp16 [rows, cout], may be null) is added to the accumulator, the sum goes to `raw` (fp16, may be null, may alias
// res) and its activation to act_out.  Those two streams use 32-byte per-thread loads / stores (shared memory is full).
template <int kCpw, bool kRes>  // epilogue columns per warp = N / 4
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kPairThreads, 1)
tc_conv3x3_pair_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w,
                       const __grid_constant__ CUtensorMap map_o64, const __grid_constant__ CUtensorMap map_o32, int rows,
                       int cin, int cout, int n_half, int stages, int staged, int tmem_cols, TcTaps tap,
                       __nv_bfloat16* __restrict__ act_out, const float* __restrict__ scale,
                       const float* __restrict__ shift, int act_mode, int debug, unsigned long long* trace,
                       const __half* res, __half* raw, int f16, int reverse) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  // P3_TC_TRACE: globaltimer stamps of CTA 0 (ns since its first instruction), summed over launches
  const bool tr = trace != nullptr && blockIdx.x == 0;
  const unsigned long long t_start = tr ? global_timer() : 0ull;
  if (tr && threadIdx.x == 0) {
    atomicAdd(&trace[0], 1ull);
    if (g_last_kernel_end != 0) atomicAdd(&trace[1], t_start - g_last_kernel_end);
  }
  const int k_slabs = cin / kSlabK;
  const int w_slab_bytes = n_half * 128;
  const int w_bytes = 9 * k_slabs * w_slab_bytes;
  uint8_t* smem_w = smem;
  uint8_t* smem_a = smem + ((w_bytes + 1023) & ~1023);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem_a + stages * kResABytes + (staged ? kResEpiWarps * 2048 : 0));
  uint64_t* empty_bar = full_bar + 4;
  uint64_t* tmem_full = empty_bar + 4;
  uint64_t* tmem_empty = tmem_full + 4;
  uint64_t* w_bar = tmem_empty + 4;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(w_bar + 1);
  float* s_scale = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(full_bar) + 256);
  // accumulator stages in TMEM: as many N-column accumulators as the allocation holds (4 at N <= 128), so that
  // epilogue / MMA jitter is absorbed instead of stalling the tensor pipe
  const int acc_stages = (tmem_cols / (2 * n_half)) >= 4 ? 4 : 2;
  float* s_shift = s_scale + kPairMaxN;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = ptx::cluster_ctarank();
  const int N = 2 * n_half;
  const int n_slices = cout / N;
  const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
  const int slice = pair % n_slices;
  const int n0 = slice * N;
  const int pair_in_slice = pair / n_slices, pairs_per_slice = n_pairs / n_slices;
  const int m_tiles = pair_tile_count(rows);  // position-aligned pair tiles (common.cuh)
  const int rev_last = reverse ? m_tiles - 1 : -1;

  for (int c = threadIdx.x; c < N; c += blockDim.x) {
    s_scale[c] = act_mode == kActMishBN ? scale[n0 + c] : 1.0f;
    s_shift[c] = act_mode == kActMishBN ? shift[n0 + c] : 0.0f;
  }
  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&map_a);
    ptx::prefetch_tensormap(&map_w);
    for (int s = 0; s < stages; ++s) {
      ptx::mbar_init(&full_bar[s], 1);   // leader's is the live one: one arrive.expect_tx for both CTAs' bytes
      ptx::mbar_init(&empty_bar[s], 1);  // multicast commit arrives on both CTAs' copies
    }
    for (int s = 0; s < acc_stages; ++s) {
      ptx::mbar_init(&tmem_full[s], 1);
      ptx::mbar_init(&tmem_empty[s], 2 * kResEpiWarps);  // leader's: epilogue warps of both CTAs
    }
    ptx::mbar_init(w_bar, 1);
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
  if (tr && threadIdx.x == 0) atomicAdd(&trace[2], global_timer() - t_start);
  if (threadIdx.x == 0) ptx::griddep_launch_dependents();

  if (warp == 0) {
    // ===== TMA producer (both CTAs) =====
    const uint32_t w_bar_leader = ptx::mapa_shared(ptx::smem_u32(w_bar), 0);
    if (ptx::elect_one()) {
      if (rank == 0) ptx::mbar_arrive_expect_tx(w_bar, static_cast<uint32_t>(2 * w_bytes));
      for (int t = 0; t < 9; ++t)
        for (int ks = 0; ks < k_slabs; ++ks)
          ptx::tma_load_2d_pair(smem_w + (t * k_slabs + ks) * w_slab_bytes, &map_w, w_bar_leader, ks * kSlabK,
                                t * cout + n0 + static_cast<int>(rank) * n_half);
    }
    __syncwarp();
    ptx::griddep_wait();  // (PDL) the resident weights above do not depend on the previous kernel; the activations do
    int stage = 0;
    uint32_t phase = 0;
    for (int mt = pair_in_slice; mt < m_tiles; mt += pairs_per_slice) {
      const int m0 = pair_tile_row0(mt, static_cast<int>(rank), rev_last);
      for (int ks = 0; ks < k_slabs; ++ks) {
        ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
        if (ptx::elect_one()) {
          if (debug & 2) {
            if (rank == 0) ptx::mbar_arrive(&full_bar[stage]);  // ablation: no A traffic
          } else {
            if (rank == 0) ptx::mbar_arrive_expect_tx(&full_bar[stage], 2 * kResABytes);
            ptx::tma_load_2d_pair_h<P3_HINT_ACT_LOAD>(smem_a + stage * kResABytes, &map_a, ptx::mapa_shared(ptx::smem_u32(&full_bar[stage]), 0),
                                  ks * kSlabK, m0 - kResHalo);
          }
        }
        __syncwarp();
        if (++stage == stages) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    if (rank == 0) {
      // ===== MMA issuer (leader CTA only) =====
      const uint32_t idesc = ptx::make_idesc_op(2 * kTileM, N, f16);
      const uint32_t w_lo_base = ptx::desc_lo_sw128(ptx::smem_u32(smem_w));
      const uint32_t w_slab16 = static_cast<uint32_t>(w_slab_bytes / 16);
      const uint32_t w_tap_stride = static_cast<uint32_t>(k_slabs) * w_slab16;
      uint32_t tap_lo[9];
#pragma unroll
      for (int t = 0; t < 9; ++t) tap_lo[t] = static_cast<uint32_t>(kResHalo + tap.off[t]) * 8u;
      ptx::mbar_wait_cluster(w_bar, 0);
      ptx::tc_fence_after_sync();
      if (tr && lane == 0) atomicAdd(&trace[3], global_timer() - t_start);
      int stage = 0;
      uint32_t phase = 0;
      int iter = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int mt = pair_in_slice; mt < m_tiles; mt += pairs_per_slice, ++iter) {
        ptx::mbar_wait_cluster(&tmem_empty[acc], acc_phase ^ 1);
        ptx::tc_fence_after_sync();
        const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(acc * N);
        for (int ks = 0; ks < k_slabs; ++ks) {
          ptx::mbar_wait_cluster(&full_bar[stage], phase);
          ptx::tc_fence_after_sync();
          const uint32_t a_lo = ptx::desc_lo_sw128(ptx::smem_u32(smem_a + stage * kResABytes));
          const uint32_t w_lo = w_lo_base + static_cast<uint32_t>(ks) * w_slab16;
          if (ptx::elect_one()) {
#pragma unroll
            for (int t = 0; t < 9; ++t) {
              if (debug & 4) continue;  // ablation: no MMAs
              const uint32_t a_t = a_lo + tap_lo[t];
              const uint32_t w_t = w_lo + static_cast<uint32_t>(t) * w_tap_stride;
#pragma unroll
              for (int k = 0; k < kSlabK / kUmmaK; ++k)
                ptx::umma_f16_pair_lohi(tmem_d, a_t + 2 * k, ptx::desc_hi_sw128(), w_t + 2 * k, ptx::desc_hi_sw128(), idesc,
                                        (ks > 0 || t > 0 || k > 0) ? 1u : 0u);
            }
            ptx::umma_commit_pair(&empty_bar[stage]);
          }
          __syncwarp();
          if (++stage == stages) {
            stage = 0;
            phase ^= 1;
          }
        }
        if (ptx::elect_one()) ptx::umma_commit_pair(&tmem_full[acc]);
        __syncwarp();
        if (++acc == acc_stages) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
    }
  } else {
    // ===== epilogue (both CTAs): thread = one of the CTA's 128 rows x N/4 columns =====
    const int ew = warp - 2;
    const int quarter = warp & 3;
    const int c0 = (ew >> 2) * kCpw;  // first column (within the slice) of this warp
    const uint32_t sca = ptx::smem_u32(s_scale + c0), sha = ptx::smem_u32(s_shift + c0);
    const uint32_t empty_leader = ptx::mapa_shared(ptx::smem_u32(&tmem_empty[0]), 0);
    const uint32_t my_stage = ptx::smem_u32(smem_a + stages * kResABytes) + static_cast<uint32_t>(ew) * 2048u;  // 2 KB per warp
    int iter = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int mt = pair_in_slice; mt < m_tiles; mt += pairs_per_slice, ++iter) {
      ptx::mbar_wait(&tmem_full[acc], acc_phase);
      ptx::tc_fence_after_sync();
      if (tr && iter == 0 && ew == 0 && lane == 0) atomicAdd(&trace[4], global_timer() - t_start);
      const int m = pair_tile_row0(mt, static_cast<int>(rank), rev_last) + quarter * 32 + lane;
      const bool in_range = m < rows;
      const bool live = in_range && row_is_live(m % kRowsPerPos);
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + static_cast<uint32_t>(acc * N + c0);
      __nv_bfloat16* ap = act_out + static_cast<size_t>(m) * cout + n0 + c0;
      // passes of 16 (then 8) columns keep the live registers small; the TMEM stage is released to the leader's barrier
      // as soon as the LAST pass's values are in registers (the other accumulator stages cover the wait)
      bool first_pass = true;
      auto pass = [&](const int col, auto width_tag, const bool last) {
        constexpr int kW = decltype(width_tag)::value;
        uint32_t v[kW];
        if (kW == 16) ptx::tmem_ld_32x16(taddr + col, v);
        else ptx::tmem_ld_32x8(taddr + col, v);
        uint4 rq[kW / 8];
        if (kRes) {  // residual values of this row, in flight while the TMEM load completes
#pragma unroll
          for (int g = 0; g < kW / 8; ++g) rq[g] = make_uint4(0, 0, 0, 0);
          if (res != nullptr && in_range) {
            const uint4* rp = reinterpret_cast<const uint4*>(res + static_cast<size_t>(m) * cout + n0 + c0 + col);
#pragma unroll
            for (int g = 0; g < kW / 8; ++g) rq[g] = rp[g];
          }
        }
        ptx::tmem_ld_wait();
        if (kRes) {
#pragma unroll
          for (int g = 0; g < kW / 8; ++g) {
            const uint32_t u[4] = {rq[g].x, rq[g].y, rq[g].z, rq[g].w};
            uint32_t o[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float2 f2 = __half22float2(*reinterpret_cast<const __half2*>(&u[i]));
              const float x0 = __uint_as_float(v[g * 8 + 2 * i]) + f2.x, x1 = __uint_as_float(v[g * 8 + 2 * i + 1]) + f2.y;
              v[g * 8 + 2 * i] = __float_as_uint(x0);
              v[g * 8 + 2 * i + 1] = __float_as_uint(x1);
              o[i] = pack_f16(x0, x1);
            }
            // padding rows / columns of the residual stream stay zero too (a 3x3 output there is not zero; the first layer
            // does not rewrite them every step: init_tc2.cu)
            if (raw != nullptr && in_range)
              *reinterpret_cast<uint4*>(raw + static_cast<size_t>(m) * cout + n0 + c0 + col + g * 8) =
                  live ? make_uint4(o[0], o[1], o[2], o[3]) : make_uint4(0, 0, 0, 0);
          }
        }
        if (last) {
          ptx::tc_fence_before_sync();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive_remote(empty_leader + 8u * static_cast<uint32_t>(acc));
        }
        if (debug & 1) return;  // ablation: no epilogue math / stores
        uint4 pk[kW / 8];
#pragma unroll
        for (int g = 0; g < kW / 8; ++g) {
          const uint32_t o = static_cast<uint32_t>(col + g * 8) * 4u;
          const float4 s0 = ptx::lds_f4_const(sca + o), s1 = ptx::lds_f4_const(sca + o + 16);
          const float4 h0 = ptx::lds_f4_const(sha + o), h1 = ptx::lds_f4_const(sha + o + 16);
          const float sc[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
          const float sh[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
          float a[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float x = __uint_as_float(v[g * 8 + i]);
            a[i] = (act_mode == kActIdentity || (debug & 16)) ? x : mish_f32<false>(fmaf(x, sc[i], sh[i]));
          }
          const uint4 q = make_uint4(pack_act(a[0], a[1], f16), pack_act(a[2], a[3], f16), pack_act(a[4], a[5], f16), pack_act(a[6], a[7], f16));
          pk[g] = live ? q : make_uint4(0, 0, 0, 0);  // padding rows / columns of the layout stay zero
        }
        if (staged) {
          if (first_pass) {  // this warp's previous bulk store has read the staging tile
            if (lane == 0) ptx::bulk_wait_read<0>();
            __syncwarp();
          }
#pragma unroll
          for (int g = 0; g < kW / 8; ++g) {
            const uint32_t k = static_cast<uint32_t>(col / 8 + g);  // 16-byte chunk of this warp's row
            const uint32_t ks = kCpw == 32 ? (k ^ ((lane >> 1) & 3u)) : kCpw == 16 ? (k ^ ((lane >> 2) & 1u)) : k;
            ptx::sts_u4(my_stage + static_cast<uint32_t>(lane) * (kCpw * 2) + (ks << 4), pk[g]);
          }
        } else if (in_range && (!(debug & 8) || pk[0].x == 0x12345678u)) {
          // direct stores: one full, aligned 32-byte sector per thread where the column offset allows it
          if (kW == 16 && kCpw % 16 == 0) ptx::stg_u8(ap + col, pk[0], pk[kW / 8 - 1]);
          else if (kW == 16) { *reinterpret_cast<uint4*>(ap + col) = pk[0]; *reinterpret_cast<uint4*>(ap + col + 8) = pk[kW / 8 - 1]; }
          else *reinterpret_cast<uint4*>(ap + col) = pk[0];
        }
        first_pass = false;
      };
#pragma unroll 1
      for (int h = 0; h < kCpw / 16; ++h) pass(h * 16, std::integral_constant<int, 16>{}, (kCpw & 8) == 0 && h == kCpw / 16 - 1);
      if (kCpw & 8) pass(kCpw & 16, std::integral_constant<int, 8>{}, true);
      if (staged && !(debug & 1)) {
        // this warp's 32 rows x kCpw columns leave as ONE TMA bulk store (no per-thread global stores, no CTA barrier)
        ptx::fence_proxy_async();
        __syncwarp();
        if (lane == 0 && !(debug & 8)) {
          ptx::tma_store_2d_h<P3_HINT_ACT_STORE>(&map_o64, my_stage, n0 + c0, pair_tile_row0(mt, static_cast<int>(rank), rev_last) + quarter * 32);
          ptx::bulk_commit();
        }
      }
      if (++acc == acc_stages) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
    if (staged && lane == 0) ptx::bulk_wait_all();
  }

  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::cluster_sync_all();  // the peer's MMAs / TMEM traffic are complete before either CTA frees its columns
  if (warp == 1) {
    ptx::tc_fence_after_sync();
    ptx::tmem_dealloc_pair(tmem_base, static_cast<uint32_t>(tmem_cols));
  }
  if (trace != nullptr && threadIdx.x == 0) {
    const unsigned long long t_end = global_timer();
    if (tr) atomicAdd(&trace[5], t_end - t_start);
    atomicMax(&g_last_kernel_end, t_end);
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

// 2-D row-major [dim1, dim0] tensor of `elem_bytes` elements, box [box1, box0], zero OOB fill.
int make_map_2d(CUtensorMap* map, const void* base, CUtensorMapDataType dt, int elem_bytes, uint64_t dim0, uint64_t dim1,
                uint32_t box0, uint32_t box1, CUtensorMapSwizzle swz) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return fail(P3_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
  cuuint64_t gdim[2] = {dim0, dim1};
  cuuint64_t gstride[1] = {dim0 * elem_bytes};
  cuuint32_t box[2] = {box0, box1};
  cuuint32_t estride[2] = {1, 1};
  CUresult r = fn(map, dt, 2, const_cast<void*>(base), gdim, gstride, box, estride, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(P3_ERR_CUDA, "cuTensorMapEncodeTiled failed: " + std::to_string(r));
  return P3_OK;
}
int make_map_bf16_k64(CUtensorMap* map, const void* base, uint64_t dim0, uint64_t dim1, uint32_t box1) {
  return make_map_2d(map, base, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dim0, dim1, kSlabK, box1, CU_TENSOR_MAP_SWIZZLE_128B);
}

typedef void (*PairKernelFn)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const CUtensorMap, int, int, int, int,
                             int, int, int, TcTaps, __nv_bfloat16*,
                             const float*, const float*, int, int, unsigned long long*, const __half*, __half*, int, int);
PairKernelFn pair_kernel_for(int N, bool res = false) {
  switch (N) {
    case 128: return res ? tc_conv3x3_pair_kernel<32, true> : tc_conv3x3_pair_kernel<32, false>;
    case 96: return res ? tc_conv3x3_pair_kernel<24, true> : tc_conv3x3_pair_kernel<24, false>;
    case 64: return res ? tc_conv3x3_pair_kernel<16, true> : tc_conv3x3_pair_kernel<16, false>;
    default: return res ? tc_conv3x3_pair_kernel<8, true> : tc_conv3x3_pair_kernel<8, false>;
  }
}

int pick_n_tile(int cout) {  // largest divisor of cout that is <= 128 and a multiple of 32
  for (int n = kMaxNTile; n >= 32; n -= 32)
    if (cout % n == 0) return n;
  return 0;
}

}  // namespace

bool tc_conv_supported(int cin, int cout) {
  if (cin <= 0 || cin % kSlabK != 0) return false;
  return pick_n_tile(cout) >= 32;
}

int tc_conv_plan_create(const __nv_bfloat16* in, const __nv_bfloat16* w, int rows, int cin, int cout, int taps,
                        const int* tap_off_host, const ConvEpilogue& ep, TcConvPlan** out) {
  if (!tc_conv_supported(cin, cout)) return fail(P3_ERR_UNSUPPORTED, "tc_conv: cin % 64 != 0 or cout not tileable");
  if (taps > kMaxTapsTc) return fail(P3_ERR_INVALID_ARG, "tc_conv: too many taps");
  TcConvPlan* p = new TcConvPlan();
  p->rows = rows;
  p->cin = cin;
  p->cout = cout;
  p->taps = taps;
  p->ep = ep;
  for (int t = 0; t < taps; ++t) p->tap.off[t] = tap_off_host[t];
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const char* env_res = std::getenv("P3_TC_RESIDENT");
  if (const char* dbg = std::getenv("P3_TC_DEBUG")) p->debug = std::atoi(dbg);
  if (std::getenv("P3_TC_TRACE")) {
    cudaMalloc(&p->trace, 16 * sizeof(unsigned long long));
    cudaMemset(p->trace, 0, 16 * sizeof(unsigned long long));
  }
  const size_t res_smem =
      static_cast<size_t>(9) * (cin / kSlabK) * kResWSlabBytes + kResStages * kResABytes + 1024 + kBarBytes;
  bool shifts_ok = taps == 9;
  for (int t = 0; t < taps && shifts_ok; ++t) shifts_ok = std::abs(tap_off_host[t]) <= kResHalo;
  // the resident kernel writes only the activated copy (all 3x3 layers except a classic block's second conv)
  p->resident = shifts_ok && cout % kResN == 0 && res_smem <= static_cast<size_t>(kSmemBudget) && ep.residual == nullptr &&
                ep.raw_out == nullptr && ep.act_out != nullptr && !(env_res && std::atoi(env_res) == 0);
  // CTA-pair kernel: largest N = 2 * n_half in {128, 96, 64, 32} dividing cout whose weight half fits next to >= 2 A stages
  const char* env_pair = std::getenv("P3_TC_PAIR");
  const bool with_res = ep.residual != nullptr || ep.raw_out != nullptr;
  const char* env_pr = std::getenv("P3_TC_PAIR_RES");
  if (shifts_ok && (!with_res || (ep.raw_f16 && !(env_pr && std::atoi(env_pr) == 0))) && ep.act_out != nullptr &&
      !(env_pair && std::atoi(env_pair) == 0) && !(env_res && std::atoi(env_res) == 0)) {
    for (int N = kPairMaxN; N >= 32 && !p->pair; N -= 32) {
      if (cout % N != 0) continue;
      const size_t wb = (static_cast<size_t>(9) * (cin / kSlabK) * (N / 2) * 128 + 1023) & ~size_t(1023);
      const char* env_st = std::getenv("P3_TC_PAIR_STAGES");  // perf experiments
      const char* env_sg = std::getenv("P3_TC_PAIR_STAGED");
      // preferred: 2 A stages (measured as fast as 3: the A boxes mostly hit L2) + an N x 128 staging tile for TMA stores;
      // else direct per-thread stores with up to 3 A stages
      const size_t fixed = wb + 1024 + 256 + 2 * kPairMaxN * 4;
      const size_t need_staged = fixed + 2 * static_cast<size_t>(kResABytes) + static_cast<size_t>(kResEpiWarps) * 2048;
      if (need_staged <= static_cast<size_t>(kSmemBudget) && !(env_sg && std::atoi(env_sg) == 0)) {
        p->pair = true;
        p->staged = true;
        p->stages = 2;
        p->smem_bytes = need_staged;
      } else {
        for (int st = env_st ? std::atoi(env_st) : 3; st >= 2 && !p->pair; --st) {
          const size_t need = fixed + static_cast<size_t>(st) * kResABytes;
          if (need <= static_cast<size_t>(kSmemBudget)) {
            p->pair = true;
            p->stages = st;
            p->smem_bytes = need;
          }
        }
      }
      if (p->pair) {
        p->n_half = N / 2;
        p->n_tile = N;
      }
    }
  }
  int rc;
  if (p->pair) {
    p->resident = false;
    p->tmem_cols = 512;  // one CTA per SM: take all columns -> 4 accumulator stages at N <= 128
    const int n_slices = cout / p->n_tile;
    const int m_tiles = pair_tile_count(rows);
    rc = make_map_bf16_k64(&p->map_a, in, cin, rows, kResRows);
    if (rc == P3_OK) rc = make_map_bf16_k64(&p->map_w, w, cin, static_cast<uint64_t>(taps) * cout, p->n_half);
    p->map_raw = p->map_a;  // placeholders unless staged
    p->map_act = p->map_a;
    if (rc == P3_OK && p->staged) {  // per-warp output boxes: 32 rows x N/4 columns
      const int cpw = p->n_tile / 4;
      rc = make_map_2d(&p->map_raw, ep.act_out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, cout, rows, cpw, 32,
                       cpw == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : cpw == 16 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_NONE);
    }
    if (rc == P3_OK) {
      cudaError_t e = cudaFuncSetAttribute(pair_kernel_for(p->n_tile, with_res), cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget);
      if (e != cudaSuccess) rc = fail(P3_ERR_CUDA, std::string("cudaFuncSetAttribute(smem): ") + cudaGetErrorString(e));
    }
    int max_pairs = sms / 2;
    if (rc == P3_OK) {  // how many CTA pairs can be co-resident (one CTA per SM, both SMs of a TPC)
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(static_cast<unsigned>(sms / 2 * 2));
      cfg.blockDim = dim3(kPairThreads);
      cfg.dynamicSmemBytes = p->smem_bytes;
      int n_clusters = 0;
      if (cudaOccupancyMaxActiveClusters(&n_clusters, pair_kernel_for(p->n_tile, with_res), &cfg) == cudaSuccess && n_clusters > 0)
        max_pairs = std::min(max_pairs, n_clusters);
      else
        cudaGetLastError();
    }
    const int pairs = std::max(n_slices, std::min(max_pairs, m_tiles * n_slices) / n_slices * n_slices);
    p->grid = 2 * pairs;
  } else if (p->resident) {
    p->n_tile = kResN;
    p->stages = kResStages;
    p->smem_bytes = res_smem;
    p->tmem_cols = 2 * kResN;
    const int n_slices = cout / kResN;
    const int m_tiles = (rows + kTileM - 1) / kTileM;
    p->grid = std::max(n_slices, std::min(sms, m_tiles * n_slices) / n_slices * n_slices);
    rc = make_map_bf16_k64(&p->map_a, in, cin, rows, kResRows);
    if (rc == P3_OK) rc = make_map_bf16_k64(&p->map_w, w, cin, static_cast<uint64_t>(taps) * cout, kResN);
    if (rc == P3_OK) {
      cudaError_t e = cudaFuncSetAttribute(tc_conv3x3_res_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget);
      if (e != cudaSuccess) rc = fail(P3_ERR_CUDA, std::string("cudaFuncSetAttribute(smem): ") + cudaGetErrorString(e));
    }
  } else {
    p->n_tile = pick_n_tile(cout);
    const int stage_bytes = kABytes + p->n_tile * kSlabK * 2;
    const int staging = staging_bytes(ep.residual != nullptr, ep.raw_out != nullptr, ep.act_out != nullptr, ep.raw_f16);
    p->stages = std::min(8, (kSmemBudget - 1024 - kBarBytes - staging) / stage_bytes);
    p->smem_bytes = static_cast<size_t>(p->stages) * stage_bytes + 1024 /*align slack*/ + kBarBytes + staging;
    int cols = 32;
    while (cols < 2 * p->n_tile) cols *= 2;
    p->tmem_cols = cols;
    const int total_tiles = ((rows + kTileM - 1) / kTileM) * (cout / p->n_tile);
    p->grid = std::min(total_tiles, sms);
    rc = make_map_bf16_k64(&p->map_a, in, cin, rows, kTileM);
    if (rc == P3_OK) rc = make_map_bf16_k64(&p->map_w, w, cin, static_cast<uint64_t>(taps) * cout, p->n_tile);
    // epilogue maps: fp32 [rows, cout] boxes of 128 rows x 32 cols (128 B rows, 128B swizzle);
    //                bf16 [rows, cout] boxes of 128 rows x 32 cols (64 B rows, 64B swizzle)
    const CUtensorMapDataType raw_dt = ep.raw_f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
    const int raw_es = ep.raw_f16 ? 2 : 4;
    const CUtensorMapSwizzle raw_sw = ep.raw_f16 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B;
    if (rc == P3_OK && ep.residual) rc = make_map_2d(&p->map_res, ep.residual, raw_dt, raw_es, cout, rows, 32, kTileM, raw_sw);
    if (rc == P3_OK && ep.raw_out && ep.raw_transposed)  // [cout, rows] fp32: boxes of 32 channels x 128 rows, rows contiguous
      rc = make_map_2d(&p->map_raw, ep.raw_out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, rows, cout, kTileM, 32, CU_TENSOR_MAP_SWIZZLE_NONE);
    else if (rc == P3_OK && ep.raw_out)
      rc = make_map_2d(&p->map_raw, ep.raw_out, raw_dt, raw_es, cout, rows, 32, kTileM, raw_sw);
    if (rc == P3_OK && ep.act_out)
      rc = make_map_2d(&p->map_act, ep.act_out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, cout, rows, 32, kTileM,
                       CU_TENSOR_MAP_SWIZZLE_64B);
    if (!ep.residual) p->map_res = p->map_a;  // never dereferenced (has_res = 0); keeps the kernel parameter well-formed
    if (!ep.raw_out) p->map_raw = p->map_a;
    if (!ep.act_out) p->map_act = p->map_a;
    if (rc == P3_OK) {
      cudaError_t e = cudaFuncSetAttribute(tc_conv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget);
      if (e != cudaSuccess) rc = fail(P3_ERR_CUDA, std::string("cudaFuncSetAttribute(smem): ") + cudaGetErrorString(e));
    }
  }
  if (rc != P3_OK) {
    delete p;
    return rc;
  }
  *out = p;
  return P3_OK;
}

void tc_conv_plan_destroy(TcConvPlan* plan) {
  if (plan && plan->trace) {
    unsigned long long h[16];
    cudaMemcpy(h, plan->trace, sizeof h, cudaMemcpyDeviceToHost);
    if (plan->pair && h[0])
      std::fprintf(stderr, "[p3 trace] pair 3x3 cin=%d cout=%d launches=%llu ns/launch (CTA 0): gap-since-previous-kernel-end %llu  "
                           "prologue %llu  weights-ready %llu  first-accumulator %llu  kernel %llu\n",
                   plan->cin, plan->cout, h[0], h[1] / h[0], h[2] / h[0], h[3] / h[0], h[4] / h[0], h[5] / h[0]);
    else if (h[7])
      std::fprintf(stderr, "[p3 trace] cin=%d cout=%d taps=%d res=%d raw=%d act=%d chunks=%llu cycles/chunk: res_wait %llu  tmem_ld %llu  math %llu  "
                           "bulk_wait %llu  barA %llu  stage+fence %llu  barB+tma %llu  | per launch: epilogue loop %llu cycles, of which waiting for accumulators %llu\n",
                   plan->cin, plan->cout, plan->taps, plan->ep.residual != nullptr, plan->ep.raw_out != nullptr, plan->ep.act_out != nullptr,
                   h[7], h[0] / h[7], h[1] / h[7], h[2] / h[7], h[3] / h[7], h[4] / h[7], h[5] / h[7], h[6] / h[7],
                   h[12] ? h[11] / h[12] : 0ull, h[12] ? h[10] / h[12] : 0ull);
    cudaFree(plan->trace);
  }
  delete plan;
}

bool tc_conv_plan_set_reverse(TcConvPlan* p, bool reverse) {  // false: not a pair-kernel plan (its tile order is fixed)
  if (!p || !p->pair) return false;
  p->reverse = reverse ? 1 : 0;
  return true;
}

int tc_conv_launch(const TcConvPlan* p, cudaStream_t stream) {
  const ConvEpilogue& ep = p->ep;
  if (p->pair) {
    const bool with_res = ep.residual != nullptr || ep.raw_out != nullptr;
    P3_CUDA(tc_launch_pdl(pair_kernel_for(p->n_tile, with_res), p->grid, kPairThreads, p->smem_bytes, stream, p->map_a, p->map_w,
                          p->map_raw, p->map_act, p->rows, p->cin, p->cout, p->n_half, p->stages, p->staged ? 1 : 0, p->tmem_cols,
                          p->tap, reinterpret_cast<__nv_bfloat16*>(ep.act_out), ep.scale, ep.shift, ep.act_mode, p->debug & 0xff,
                          p->trace, reinterpret_cast<const __half*>(ep.residual), reinterpret_cast<__half*>(ep.raw_out),
                          ep.op_f16 ? 1 : 0, p->reverse));
  } else if (p->resident) {
    tc_conv3x3_res_kernel<<<p->grid, kResThreads, p->smem_bytes, stream>>>(
        p->map_a, p->map_w, p->rows, p->cin, p->cout, p->tap, reinterpret_cast<__nv_bfloat16*>(ep.act_out), ep.scale,
        ep.shift, ep.act_mode, p->debug & 0xff, ep.op_f16 ? 1 : 0);
  } else {
    tc_conv_kernel<<<p->grid, kNumThreads, p->smem_bytes, stream>>>(
        p->map_a, p->map_w, p->map_res, p->map_raw, p->map_act, p->rows, p->cin, p->cout, p->taps, p->tap, p->n_tile,
        p->stages, p->tmem_cols, ep.residual != nullptr && !(p->debug & 32), ep.raw_out != nullptr, ep.act_out != nullptr,
        ep.scale, ep.shift, ep.act_mode, ep.raw_f16 ? 1 : 0, (ep.raw_transposed && !ep.raw_f16) ? 1 : 0, p->debug >> 8, p->trace,
        ep.op_f16 ? 1 : 0);
  }
  P3_CUDA(cudaGetLastError());
  return P3_OK;
}

}  // namespace p3
