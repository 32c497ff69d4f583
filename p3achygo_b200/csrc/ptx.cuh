// Hand-written sm_100a PTX wrappers: mbarrier, TMA (cp.async.bulk.tensor), tcgen05 (MMA, TMEM
// alloc / ld / commit / fences).  Encodings follow the PTX ISA for sm_100a; descriptor bit layouts
// are documented next to the builders below.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace p3 {
namespace ptx {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// One lane of a converged warp (elect.sync): the compiler knows the predicate is "exactly one lane", so address /
// descriptor arithmetic around the guarded instruction stays on the uniform datapath (no R2UR waterfall loops).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t"
      "}"
      : "=r"(pred));
  return pred != 0;
}

// ---- explicit shared-state-space vector accesses (pointers carved out of dynamic smem at run time lose their
// address space and compile to generic LD.E / ST.E otherwise) ----
__device__ __forceinline__ float4 lds_f4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ float4 lds_f4_const(uint32_t addr) {  // data written before the last barrier, never after
  float4 v;
  asm("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts_f4(uint32_t addr, float4 v) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void sts_u4(uint32_t addr, uint4 v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// 256-bit global store (sm_100+, PTX 8.8): one full 32-byte sector per thread
__device__ __forceinline__ void stg_u8(void* gptr, uint4 a, uint4 b) {
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(gptr), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w),
               "r"(b.x), "r"(b.y), "r"(b.z), "r"(b.w)
               : "memory");
}

// ---- programmatic dependent launch (PDL) -------------------------------------------------------------
// A kernel launched with cudaLaunchAttributeProgrammaticStreamSerialization may start while its predecessor in the stream
// is still draining (its CTAs become resident as the predecessor's CTAs exit).  Everything that does not depend on the
// predecessor (barrier init, TMEM allocation, weight loads) runs ahead; a thread calls griddep_wait() before it first
// touches global memory the predecessor reads or writes.
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
// Lets the dependent grid be scheduled once every CTA of this grid has executed it (or exited); the dependent's CTAs then become
// resident as this grid's CTAs leave their SMs and run their own prologue up to griddep_wait().
__device__ __forceinline__ void griddep_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// ---- mbarrier ------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t"
      "}"
      : "=r"(done)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return done != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}

// ---- TMA -----------------------------------------------------------------------------------------------
__device__ __forceinline__ void prefetch_tensormap(const void* desc) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(desc)) : "memory");
}
// 2-D tiled load global -> shared, completion counted in bytes on `bar`. Out-of-bounds elements
// (including negative coordinates) are written as zeros.
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const void* desc, uint64_t* bar, int32_t c0, int32_t c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(desc)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}

// L2 prefetch of a 2-D box (no shared-memory destination, no completion)
__device__ __forceinline__ void tma_prefetch_2d(const void* desc, int32_t c0, int32_t c1) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global [%0, {%1, %2}];" ::"l"(reinterpret_cast<uint64_t>(desc)), "r"(c0), "r"(c1) : "memory");
}

// 1-D bulk copy global -> shared (16-byte aligned, size a multiple of 16), completion counted in bytes on `bar`
__device__ __forceinline__ void bulk_load_1d(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(gsrc)), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// 2-D tiled store shared -> global (bulk async group); out-of-bounds parts of the box are clipped.
__device__ __forceinline__ void tma_store_2d(const void* desc, const void* smem_src, int32_t c0, int32_t c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(desc)), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_store_2d(const void* desc, const void*, int, int, uint32_t smem_src_addr, int32_t c0, int32_t c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(desc)), "r"(smem_src_addr), "r"(c0), "r"(c1)
               : "memory");
}
// ---- L2 cache hints on TMA traffic ---------------------------------------------------------------------------------------
// Every activation tensor of the tower is written once and read once (by the next launch); the only re-read is the residual
// stream, four launches later.  P3_HINT_ACT_LOAD / P3_HINT_ACT_STORE (0 = none, 1 = evict_first, 2 = evict_last, 3 = evict_normal)
// choose the createpolicy-equivalent 64-bit descriptors below for those loads / stores (the values CUTLASS hard-codes for
// createpolicy.fractional.L2::evict_*.b64 with fraction 1.0).
#ifndef P3_HINT_ACT_LOAD
#define P3_HINT_ACT_LOAD 0
#endif
#ifndef P3_HINT_ACT_STORE
#define P3_HINT_ACT_STORE 0
#endif
#ifndef P3_HINT_RES_LOAD
#define P3_HINT_RES_LOAD P3_HINT_ACT_LOAD
#endif
#ifndef P3_HINT_RES_STORE
#define P3_HINT_RES_STORE P3_HINT_ACT_STORE
#endif
__host__ __device__ constexpr uint64_t l2_policy(int kind) {
  return kind == 1 ? 0x12F0000000000000ull : kind == 2 ? 0x14F0000000000000ull : 0x1000000000000000ull;
}
template <int kHint>
__device__ __forceinline__ void tma_load_2d_h(void* smem_dst, const void* desc, uint64_t* bar, int32_t c0, int32_t c1) {
  if constexpr (kHint == 0) {
    tma_load_2d(smem_dst, desc, bar, c0, c1);
  } else {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], [%2], %5;"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(desc)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "l"(l2_policy(kHint))
        : "memory");
  }
}
template <int kHint>
__device__ __forceinline__ void tma_store_2d_h(const void* desc, uint32_t smem_src_addr, int32_t c0, int32_t c1) {
  if constexpr (kHint == 0) {
    tma_store_2d(desc, nullptr, 0, 0, smem_src_addr, c0, c1);
  } else {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group.L2::cache_hint [%0, {%2, %3}], [%1], %4;"
                 ::"l"(reinterpret_cast<uint64_t>(desc)), "r"(smem_src_addr), "r"(c0), "r"(c1), "l"(l2_policy(kHint))
                 : "memory");
  }
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// wait until the shared-memory SOURCE of all committed bulk stores has been read (staging reusable)
__device__ __forceinline__ void bulk_wait_read_all() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
// wait until all committed bulk stores are complete (globally visible)
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
// named barrier among `count` threads (ids 1..15; 0 is __syncthreads)
__device__ __forceinline__ void named_bar_sync(int id, int count) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}

// ---- tcgen05: TMEM allocation --------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {  // whole warp, .sync.aligned
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before_sync() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after_sync() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}

// ---- tcgen05: MMA ------------------------------------------------------------------------------------------
// D[tmem] (+)= A[smem] * B[smem]; one thread issues for the CTA.
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Same, with the 64-bit descriptors given as (lo, hi) words: the MMA-issuing thread only ADDS small constants to
// the lo words (start address field, 16-byte units) between MMAs instead of rebuilding descriptors.
__device__ __forceinline__ void umma_f16_lohi(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                              uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      ".reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %2};\n\t"
      "mov.b64 db, {%3, %4};\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t"
      "}"
      ::"r"(tmem_d), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
// lo / hi words of the K-major SWIZZLE_128B descriptor (see make_desc_sw128)
__device__ __forceinline__ uint32_t desc_lo_sw128(uint32_t smem_addr) { return ((smem_addr >> 4) & 0x3FFFu) | (1u << 16); }
__host__ __device__ constexpr uint32_t desc_hi_sw128() { return (1024u >> 4) | (1u << 14) | (2u << 29); }
// wait until at most N committed bulk-store groups still have to read their shared-memory source
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}

// Arrive on `bar` when all previously issued MMAs of this thread have completed (implies
// tcgen05.fence::before_thread_sync).
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// ---- thread-block clusters / CTA pairs (cta_group::2) -----------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {  // every thread of every CTA in the cluster
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `local` (a shared::cta address of THIS CTA's layout) inside CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_shared(uint32_t local, uint32_t rank) {
  uint32_t a;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(a) : "r"(local), "r"(rank));
  return a;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {  // arrive on a (possibly remote) barrier
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// Same, default semantics (release at CTA scope) on a shared::cluster address: what CUTLASS' ClusterBarrier::arrive(cta_id)
// emits.  The .release.cluster form above compiles to MEMBAR.ALL.GPU + CGAERRBAR in front of the arrive; use this one when
// the data being published was already made visible by other means (tcgen05 fences, fence.proxy.async + a local barrier).
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {  // waits for remote arrivals too
  uint32_t done = 0;
  while (!done) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  }
}
// TMA load issued by either CTA of a pair; the transaction bytes are counted on the barrier at cluster address `bar`
// (the leader CTA's), the data lands in the issuing CTA's own shared memory.
__device__ __forceinline__ void tma_load_2d_pair(void* smem_dst, const void* desc, uint32_t bar_cluster_addr, int32_t c0,
                                                 int32_t c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(desc)), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
      : "memory");
}
template <int kHint>
__device__ __forceinline__ void tma_load_2d_pair_h(void* smem_dst, const void* desc, uint32_t bar_cluster_addr, int32_t c0, int32_t c1) {
  if constexpr (kHint == 0) {
    tma_load_2d_pair(smem_dst, desc, bar_cluster_addr, c0, c1);
  } else {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], [%2], %5;"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(desc)), "r"(bar_cluster_addr), "r"(c0), "r"(c1), "l"(l2_policy(kHint))
        : "memory");
  }
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* smem_dst, uint32_t ncols) {  // one warp in EACH CTA of the pair
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish_pair() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// M = 256 MMA across the pair: each CTA supplies its 128 rows of A and its half of B's N rows from the SAME shared-memory
// offsets; each CTA's TMEM receives its 128 rows x N columns.  Issued by one thread of the leader CTA only.
__device__ __forceinline__ void umma_f16_pair_lohi(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                                   uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      ".reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %2};\n\t"
      "mov.b64 db, {%3, %4};\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %5, p;\n\t"
      "}"
      ::"r"(tmem_d), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on the barrier at this shared-memory offset in BOTH CTAs of the pair once all MMAs issued so far have completed
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
  const uint16_t mask = 3;
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"(mask)
               : "memory");
}

// Shared-memory matrix descriptor (64-bit), K-major operand tile stored as rows of 128 bytes with
// the 128-byte swizzle TMA applied (CU_TENSOR_MAP_SWIZZLE_128B):
//   [ 0,14) start address >> 4          [16,30) leading-dim byte offset >> 4 (unused for swizzled K-major)
//   [32,46) stride-dim byte offset >> 4 = 1024 B between 8-row groups
//   [46,48) version = 1 (sm_100)        [49,52) base offset   [61,64) layout: 2 = SWIZZLE_128B
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t smem_addr, uint32_t base_offset = 0) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr >> 4) & 0x3FFF);
  d |= static_cast<uint64_t>(1) << 16;                 // LBO (ignored)
  d |= static_cast<uint64_t>(1024 >> 4) << 32;         // SBO
  d |= static_cast<uint64_t>(1) << 46;                 // version
  d |= static_cast<uint64_t>(base_offset & 7) << 49;
  d |= static_cast<uint64_t>(2) << 61;                 // SWIZZLE_128B
  return d;
}

// Instruction descriptor (32-bit) for kind::f16, BF16 x BF16 -> FP32, both operands K-major:
//   [4,6) D format: 1 = F32   [7,10) A format: 1 = BF16   [10,13) B format: 1 = BF16
//   [15] A major (0 = K)      [16] B major (0 = K)        [17,23) N >> 3      [24,29) M >> 4
__host__ __device__ constexpr uint32_t make_idesc_bf16(int m, int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(n >> 3) << 17) |
         (static_cast<uint32_t>(m >> 4) << 24);
}
// the same with the operand format chosen at run time: A / B format 0 = F16 (IEEE half), 1 = BF16 (kind::f16 takes either at the
// same rate; P3_PRECISION_FP16 runs the tower with fp16 operands, the reference's production precision)
__host__ __device__ constexpr uint32_t make_idesc_op(int m, int n, int f16) {
  return (1u << 4) | (f16 ? 0u : ((1u << 7) | (1u << 10))) | (static_cast<uint32_t>(n >> 3) << 17) |
         (static_cast<uint32_t>(m >> 4) << 24);
}

// ---- tcgen05: TMEM -> registers ------------------------------------------------------------------------------
// 32 lanes x 32 consecutive 32-bit columns: thread i of the warp receives lane (base_lane + i).
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
}
// 32 lanes x 8 consecutive 32-bit columns
__device__ __forceinline__ void tmem_ld_32x8(uint32_t taddr, uint32_t* v) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr));
}
// 32 lanes x 16 consecutive 32-bit columns
__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

}  // namespace ptx
}  // namespace p3
