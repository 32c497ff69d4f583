// Microbenchmark: does a [rows, 256] fp16 matrix read/written as 64-column (128-byte) slabs through TMA reach copy bandwidth?
//   layout A ("row-major"):  global [rows][256] halfs, a box = 32 rows x 64 cols -> 32 separate 128-byte lines, 512 B apart
//   layout B ("slab-major"): global [4][rows][64] halfs, a box = 32 rows x 64 cols -> 4 KB contiguous
// Persistent CTAs, one warp streams boxes global -> smem ring -> global (in place or to a second buffer), in the order the
// fused boundary kernel visits them (per 128-row tile: slab 0 of quarters 0-3, slab 1 of quarters 0-3, ...).
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>
#include "../../p3achygo_b200/csrc/common.cuh"
namespace p3 { int fail(int code, const std::string& msg) { std::fprintf(stderr, "fail %d %s\n", code, msg.c_str()); return code; } void set_error(const std::string&) {} }
#include "../../p3achygo_b200/csrc/tc_util.cuh"
using namespace p3;

constexpr int kBoxBytes = 32 * 128;

template <int kRing>
__global__ void __launch_bounds__(256, 1) copy_kernel(const __grid_constant__ CUtensorMap map_in, const __grid_constant__ CUtensorMap map_out,
                                                      int rows, int slabs, int slab_major, int rows_per_box, int n_warps, int ahead, int outstanding) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int box_bytes = kBoxBytes * (rows_per_box / 32);
  uint64_t* full_all = reinterpret_cast<uint64_t*>(smem + n_warps * kRing * box_bytes);
  if (threadIdx.x == 0) {
    for (int i = 0; i < n_warps * kRing; ++i) ptx::mbar_init(&full_all[i], 1);
    ptx::fence_mbar_init();
  }
  __syncthreads();
  const int w = threadIdx.x >> 5;
  if ((threadIdx.x & 31) != 0 || w >= n_warps) return;
  smem += w * kRing * box_bytes;
  uint64_t* full = full_all + w * kRing;
  const int tiles = rows / 128;
  const int boxes_per_tile = slabs * (128 / rows_per_box);
  // box ordinal o of this CTA -> (tile, slab, sub)
  long long n_all = 0;
  for (int t = blockIdx.x; t < tiles; t += gridDim.x) n_all += boxes_per_tile;
  const long long n_boxes = (n_all - w + n_warps - 1) / n_warps;   // this warp takes the CTA's boxes w, w + n_warps, ...
  auto coords = [&](long long ow, int& c0, int& c1, int& c2) {
    const long long o = ow * n_warps + w;
    const int t = blockIdx.x + static_cast<int>(o / boxes_per_tile) * gridDim.x;
    const int r = static_cast<int>(o % boxes_per_tile);
    const int slab = r / (128 / rows_per_box), sub = r % (128 / rows_per_box);
    const int row = t * 128 + sub * rows_per_box;
    if (slab_major) { c0 = 0; c1 = row; c2 = slab; } else { c0 = slab * 64; c1 = row; c2 = 0; }
  };
  auto load = [&](long long o) {
    int c0, c1, c2; coords(o, c0, c1, c2);
    const int s = static_cast<int>(o % kRing);
    ptx::mbar_arrive_expect_tx(&full[s], box_bytes);
    if (slab_major)
      asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                   ::"r"(ptx::smem_u32(smem + s * box_bytes)), "l"(reinterpret_cast<uint64_t>(&map_in)), "r"(ptx::smem_u32(&full[s])), "r"(c0), "r"(c1), "r"(c2) : "memory");
    else
      ptx::tma_load_2d(smem + s * box_bytes, &map_in, &full[s], c0, c1);
  };
  for (long long o = 0; o < ahead && o < n_boxes; ++o) load(o);
  for (long long o = 0; o < n_boxes; ++o) {
    const int s = static_cast<int>(o % kRing);
    ptx::mbar_wait(&full[s], static_cast<uint32_t>((o / kRing) & 1));
    int c0, c1, c2; coords(o, c0, c1, c2);
    if (slab_major)
      asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
                   ::"l"(reinterpret_cast<uint64_t>(&map_out)), "r"(ptx::smem_u32(smem + s * box_bytes)), "r"(c0), "r"(c1), "r"(c2) : "memory");
    else
      ptx::tma_store_2d(&map_out, nullptr, 0, 0, ptx::smem_u32(smem + s * box_bytes), c0, c1);
    ptx::bulk_commit();
    if (outstanding >= 6) ptx::bulk_wait_read<6>(); else if (outstanding >= 3) ptx::bulk_wait_read<3>(); else ptx::bulk_wait_read<1>();
    if (o + ahead < n_boxes) load(o + ahead);
  }
  ptx::bulk_wait_all();
}

static int make_map(CUtensorMap* m, void* base, int rows, int slabs, int slab_major, int rows_per_box) {
  EncodeTiledFn fn = tc_encode_fn();
  if (slab_major) {
    cuuint64_t gdim[3] = {64, (cuuint64_t)rows, (cuuint64_t)slabs};
    cuuint64_t gstride[2] = {128, (cuuint64_t)rows * 128};
    cuuint32_t box[3] = {64, (cuuint32_t)rows_per_box, 1};
    cuuint32_t es[3] = {1, 1, 1};
    return fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, base, gdim, gstride, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
              CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  }
  return tc_make_map_2d(m, base, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, (uint64_t)slabs * 64, rows, 64, rows_per_box, CU_TENSOR_MAP_SWIZZLE_128B);
}

int main(int argc, char** argv) {
  const int rows = 1024 * 400, slabs = 4;
  const size_t bytes = (size_t)rows * slabs * 128;
  void *a, *b;
  cudaMalloc(&a, bytes); cudaMalloc(&b, bytes);
  cudaMemset(a, 1, bytes); cudaMemset(b, 0, bytes);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  // reference: cudaMemcpy D2D
  for (int i = 0; i < 3; ++i) cudaMemcpy(b, a, bytes, cudaMemcpyDeviceToDevice);
  cudaEventRecord(e0); for (int i = 0; i < 10; ++i) cudaMemcpyAsync(b, a, bytes, cudaMemcpyDeviceToDevice); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  std::printf("cudaMemcpy D2D: %.1f us per %.0f MB -> %.0f GB/s (read+write)\n", ms * 100, bytes / 1e6, 2.0 * bytes / (ms / 10 * 1e-3) / 1e9);
  struct Cfg { int rpb, n_warps, ring, ahead, outstanding; };
  const Cfg cfgs[] = {{32, 1, 16, 8, 6}, {32, 4, 12, 5, 3}, {32, 4, 12, 8, 3}, {32, 4, 5, 3, 1}, {32, 4, 8, 5, 1}, {32, 8, 6, 3, 1}, {64, 4, 6, 3, 1}, {128, 1, 12, 6, 3}, {128, 2, 6, 3, 1}, {128, 4, 3, 1, 1}};
  for (const Cfg& c : cfgs)
    for (int slab_major = 0; slab_major < 2; ++slab_major) {
      const int inplace = 1, rpb = c.rpb;
      CUtensorMap mi, mo;
      if (make_map(&mi, a, rows, slabs, slab_major, rpb) || make_map(&mo, inplace ? a : b, rows, slabs, slab_major, rpb)) { std::printf("map failed\n"); return 1; }
      constexpr int kRing = 16;
      (void)kRing;
      const size_t smem = (size_t)c.n_warps * c.ring * kBoxBytes * (rpb / 32) + 1024 + 1024;
      if (smem > 227 * 1024) { std::printf("skip (smem)\n"); continue; }
      auto launch = [&]() {
        switch (c.ring) {
          case 16: copy_kernel<16><<<148, 256, smem>>>(mi, mo, rows, slabs, slab_major, rpb, c.n_warps, c.ahead, c.outstanding); break;
          case 12: copy_kernel<12><<<148, 256, smem>>>(mi, mo, rows, slabs, slab_major, rpb, c.n_warps, c.ahead, c.outstanding); break;
          case 8: copy_kernel<8><<<148, 256, smem>>>(mi, mo, rows, slabs, slab_major, rpb, c.n_warps, c.ahead, c.outstanding); break;
          case 6: copy_kernel<6><<<148, 256, smem>>>(mi, mo, rows, slabs, slab_major, rpb, c.n_warps, c.ahead, c.outstanding); break;
          case 5: copy_kernel<5><<<148, 256, smem>>>(mi, mo, rows, slabs, slab_major, rpb, c.n_warps, c.ahead, c.outstanding); break;
          case 3: copy_kernel<3><<<148, 256, smem>>>(mi, mo, rows, slabs, slab_major, rpb, c.n_warps, c.ahead, c.outstanding); break;
        }
      };
      cudaFuncSetAttribute(copy_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
      cudaFuncSetAttribute(copy_kernel<12>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
      cudaFuncSetAttribute(copy_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
      cudaFuncSetAttribute(copy_kernel<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
      cudaFuncSetAttribute(copy_kernel<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
      cudaFuncSetAttribute(copy_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
      for (int i = 0; i < 3; ++i) launch();
      cudaEventRecord(e0);
      for (int i = 0; i < 10; ++i) launch();
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      cudaError_t err = cudaGetLastError();
      cudaEventElapsedTime(&ms, e0, e1);
      std::printf("%s rows/box %3d warps %d ring %2d ahead %d outst %d (%3zu KB): %.1f us -> %.0f GB/s  [%s]\n", slab_major ? "slab-major" : "row-major ", rpb, c.n_warps, c.ring,
                  c.ahead, c.outstanding, smem / 1024, ms * 100, 2.0 * bytes / (ms / 10 * 1e-3) / 1e9, cudaGetErrorString(err));
    }
  return 0;
}
