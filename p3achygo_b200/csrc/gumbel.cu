// Gumbel root sampling (cc/mcts/gumbel.cc:283-321) as a warp-shuffle kernel: one warp per root.
//
//   for i in 0..361 (index order): illegal -> score -10000, no noise, NO PRNG draw
//                                  legal   -> score = logit[i] + noise_scaling * Gumbel(0,1)   (one PCG32 draw)
//   top-k by score (the reference std::sort's all 362 and takes the first min(k, k_valid)).
//
// The serial PRNG stream is parallelised exactly: the r-th legal move (r = its rank among legal
// moves, a ballot prefix sum) uses the PCG state after r LCG steps, reached by an O(log r) jump
// (cc/core/rand.cc:32-43 pcg32; cc/core/probability.cc:12-30 Uniform/GumbelSample).  The uniform is
// bit-exact, and so is -logf(-logf(u)): logf_glibc below is the C library routine the reference calls
// (GNU libc 2.39, sysdeps/ieee754/flt-32/e_logf.c: 16-entry {1/c, log c} table, degree-3 polynomial, all in
// double, one rounding to float), restated in oracle/features_oracle.c::orc_logf and checked there against libm
// over every positive float.  Its double arithmetic is IEEE-exact on the GPU as on the host, fused or not,
// so scores AND the selected set are compared with == in the tests.
#include "common.cuh"

namespace p3 {
namespace {

constexpr unsigned long long kPcgMult = 6364136223846793005ULL;
constexpr unsigned long long kPcgInc = 1442695040888963407ULL;
constexpr int kRounds = 12;  // 12 * 32 = 384 >= 362
constexpr float kSmallLogit = -10000.0f;  // gumbel.cc:28

__device__ __forceinline__ unsigned long long pcg_advance(unsigned long long state, unsigned delta) {
  unsigned long long acc_mult = 1, acc_plus = 0, cur_mult = kPcgMult, cur_plus = kPcgInc;
  while (delta) {
    if (delta & 1u) {
      acc_mult *= cur_mult;
      acc_plus = acc_plus * cur_mult + cur_plus;
    }
    cur_plus = (cur_mult + 1) * cur_plus;
    cur_mult *= cur_mult;
    delta >>= 1;
  }
  return acc_mult * state + acc_plus;
}

__device__ __forceinline__ uint32_t pcg_output(unsigned long long x) {  // rand.cc:32-43, on the pre-step state
  const unsigned count = static_cast<unsigned>(x >> 59);
  x ^= x >> 18;
  const uint32_t v = static_cast<uint32_t>(x >> 27);
  return v >> count | v << ((-count) & 31);
}

// glibc 2.39 logf (e_logf.c / e_logf_data.c): T[i] = {invc, logc}, poly A, ln2; see oracle/features_oracle.c::orc_logf
__constant__ double kLogfTab[16][2] = {
    {0x1.661ec79f8f3bep+0, -0x1.57bf7808caadep-2}, {0x1.571ed4aaf883dp+0, -0x1.2bef0a7c06ddbp-2},
    {0x1.49539f0f010bp+0, -0x1.01eae7f513a67p-2},  {0x1.3c995b0b80385p+0, -0x1.b31d8a68224e9p-3},
    {0x1.30d190c8864a5p+0, -0x1.6574f0ac07758p-3}, {0x1.25e227b0b8eap+0, -0x1.1aa2bc79c81p-3},
    {0x1.1bb4a4a1a343fp+0, -0x1.a4e76ce8c0e5ep-4}, {0x1.12358f08ae5bap+0, -0x1.1973c5a611cccp-4},
    {0x1.0953f419900a7p+0, -0x1.252f438e10c1ep-5}, {0x1p+0, 0x0p+0},
    {0x1.e608cfd9a47acp-1, 0x1.aa5aa5df25984p-5},  {0x1.ca4b31f026aap-1, 0x1.c5e53aa362eb4p-4},
    {0x1.b2036576afce6p-1, 0x1.526e57720db08p-3},  {0x1.9c2d163a1aa2dp-1, 0x1.bc2860d22477p-3},
    {0x1.886e6037841edp-1, 0x1.1058bc8a07ee1p-2},  {0x1.767dcf5534862p-1, 0x1.4043057b6ee09p-2}};

__device__ __forceinline__ float logf_glibc(float x) {
  uint32_t ix = __float_as_uint(x);
  if (ix == 0x3f800000u) return 0.0f;
  if (ix - 0x00800000u >= 0x7f800000u - 0x00800000u) {  // zero, subnormal, negative, inf, nan
    if (ix * 2u == 0u) return -INFINITY;
    if (ix == 0x7f800000u) return x;
    if ((ix & 0x80000000u) || ix * 2u >= 0xff000000u) return __uint_as_float(0x7fc00000u);
    ix = __float_as_uint(__fmul_rn(x, 8388608.0f)) - (23u << 23);  // subnormal: normalise (no flush: __fmul_rn)
  }
  const uint32_t tmp = ix - 0x3f330000u;
  const int i = static_cast<int>((tmp >> 19) & 15u);
  const int k = static_cast<int>(tmp) >> 23;
  const double z = static_cast<double>(__uint_as_float(ix - (tmp & 0xff800000u)));
  const double r = fma(z, kLogfTab[i][0], -1.0);
  const double y0 = fma(static_cast<double>(k), 0x1.62e42fefa39efp-1, kLogfTab[i][1]);
  const double r2 = r * r;
  double y = fma(0x1.5575b0be00b6ap-2, r, -0x1.ffffef20a4123p-2);
  y = fma(-0x1.00ea348b88334p-2, r2, y);
  y = fma(y, r2, y0 + r);
  return __double2float_rn(y);
}

__device__ __forceinline__ float gumbel_from_bits(uint32_t r) {
  const float u = __uint_as_float((127u << 23) | (r >> 9)) - 1.0f;  // probability.cc:17-30
  return -logf_glibc(-logf_glibc(u));                                // probability.cc:12-15
}

__global__ void __launch_bounds__(128)
gumbel_kernel(const float* __restrict__ logits, const uint8_t* __restrict__ legal,
              unsigned long long* __restrict__ prng_state, int n, float noise_scaling, int k,
              int32_t* __restrict__ out_moves, float* __restrict__ out_scores, int32_t* __restrict__ out_kvalid,
              size_t logit_stride, const int32_t* __restrict__ slots, bool legal_by_slot) {
  const int root = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (root >= n) return;
  const size_t src = slots ? static_cast<size_t>(slots[root]) : static_cast<size_t>(root);
  const float* lg = logits + src * logit_stride;
  const uint8_t* lm = legal + (legal_by_slot ? src : static_cast<size_t>(root)) * P3_MAX_MOVES;
  const unsigned long long s0 = prng_state[root];

  float score[kRounds];
  int enc[kRounds];
  int legal_before = 0;
#pragma unroll
  for (int j = 0; j < kRounds; ++j) {
    const int i = j * 32 + lane;
    const bool in = i < P3_MAX_MOVES;
    const bool ok = in && lm[i] != 0;
    const unsigned mask = __ballot_sync(0xffffffffu, ok);
    const int rank = legal_before + __popc(mask & ((1u << lane) - 1u));
    legal_before += __popc(mask);
    if (ok) {
      // two roundings as in the reference (noise is stored, then added: gumbel.cc:296-300): no FMA contraction
      const float noise = __fmul_rn(noise_scaling, gumbel_from_bits(pcg_output(pcg_advance(s0, rank))));
      score[j] = __fadd_rn(lg[i], noise);  // + qtransform (0 at the root before any visit)
      enc[j] = i;
    } else {
      score[j] = in ? kSmallLogit : -INFINITY;
      enc[j] = -1;
    }
  }
  const int k_valid = legal_before;
  const int k_out = min(k, k_valid);

  for (int sel = 0; sel < k; ++sel) {
    // lane-local best (ties -> lower move index), then warp argmax
    float best = -INFINITY;
    int best_j = 0;
#pragma unroll
    for (int j = 0; j < kRounds; ++j)
      if (score[j] > best) {
        best = score[j];
        best_j = j;
      }
    int best_idx = best_j * 32 + lane;
    float wbest = best;
    int widx = best_idx;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ob = __shfl_xor_sync(0xffffffffu, wbest, o);
      const int oi = __shfl_xor_sync(0xffffffffu, widx, o);
      if (ob > wbest || (ob == wbest && oi < widx)) {
        wbest = ob;
        widx = oi;
      }
    }
    const int owner = widx & 31, oj = widx >> 5;
    int wenc = -1;
#pragma unroll
    for (int j = 0; j < kRounds; ++j)
      if (j == oj) {
        wenc = enc[j];
        if (lane == owner) score[j] = -INFINITY;
      }
    wenc = __shfl_sync(0xffffffffu, wenc, owner);
    if (lane == 0) {
      const bool keep = sel < k_out;
      out_moves[static_cast<size_t>(root) * k + sel] = keep ? wenc : -1;
      out_scores[static_cast<size_t>(root) * k + sel] = keep ? wbest : 0.0f;
    }
  }
  if (lane == 0) {
    out_kvalid[root] = k_valid;
    prng_state[root] = pcg_advance(s0, k_valid);
  }
}

}  // namespace

int gumbel_launch(const float* logits, const uint8_t* legal, uint64_t* prng_state, int n, float noise_scaling, int k,
                  int32_t* out_moves, float* out_scores, int32_t* out_kvalid, cudaStream_t stream, size_t logit_stride,
                  const int32_t* slots, bool legal_by_slot) {
  if (n <= 0) return P3_OK;
  if (k <= 0 || k > 64) return fail(P3_ERR_INVALID_ARG, "gumbel: k must be in [1, 64]");
  const int warps_per_block = 4;
  gumbel_kernel<<<(n + warps_per_block - 1) / warps_per_block, 32 * warps_per_block, 0, stream>>>(
      logits, legal, reinterpret_cast<unsigned long long*>(prng_state), n, noise_scaling, k, out_moves, out_scores,
      out_kvalid, logit_stride, slots, legal_by_slot);
  P3_CUDA(cudaGetLastError());
  return P3_OK;
}

}  // namespace p3
