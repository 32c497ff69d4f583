// Gumbel root sampling (cc/mcts/gumbel.cc:283-321) as a warp-shuffle kernel: one warp per root.
//
//   for i in 0..361 (index order): illegal -> score -10000, no noise, NO PRNG draw
//                                  legal   -> score = logit[i] + noise_scaling * Gumbel(0,1)   (one PCG32 draw)
//   top-k by score (the reference std::sort's all 362 and takes the first min(k, k_valid)).
//
// The serial PRNG stream is parallelised exactly: the r-th legal move (r = its rank among legal
// moves, a ballot prefix sum) uses the PCG state after r LCG steps, reached by an O(log r) jump
// (cc/core/rand.cc:32-43 pcg32; cc/core/probability.cc:12-30 Uniform/GumbelSample).  The uniform is
// bit-exact; -logf(-logf(u)) is evaluated through fp64 log and rounded to fp32, which matches glibc's
// logf except for rare last-ulp cases, so scores are tolerance-checked (1 ulp) and the selected set is
// exact whenever the top-k is not tied within that ulp.
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

__device__ __forceinline__ float gumbel_from_bits(uint32_t r) {
  const float u = __uint_as_float((127u << 23) | (r >> 9)) - 1.0f;  // probability.cc:17-30
  const float inner = static_cast<float>(log(static_cast<double>(u)));
  return -static_cast<float>(log(static_cast<double>(-inner)));      // probability.cc:12-15
}

__global__ void __launch_bounds__(128)
gumbel_kernel(const float* __restrict__ logits, const uint8_t* __restrict__ legal,
              unsigned long long* __restrict__ prng_state, int n, float noise_scaling, int k,
              int32_t* __restrict__ out_moves, float* __restrict__ out_scores, int32_t* __restrict__ out_kvalid) {
  const int root = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (root >= n) return;
  const float* lg = logits + static_cast<size_t>(root) * P3_MAX_MOVES;
  const uint8_t* lm = legal + static_cast<size_t>(root) * P3_MAX_MOVES;
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
      const float noise = noise_scaling * gumbel_from_bits(pcg_output(pcg_advance(s0, rank)));
      score[j] = lg[i] + noise;  // + qtransform (0 at the root before any visit)
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
                  int32_t* out_moves, float* out_scores, int32_t* out_kvalid, cudaStream_t stream) {
  if (n <= 0) return P3_OK;
  if (k <= 0 || k > 64) return fail(P3_ERR_INVALID_ARG, "gumbel: k must be in [1, 64]");
  const int warps_per_block = 4;
  gumbel_kernel<<<(n + warps_per_block - 1) / warps_per_block, 32 * warps_per_block, 0, stream>>>(
      logits, legal, reinterpret_cast<unsigned long long*>(prng_state), n, noise_scaling, k, out_moves, out_scores,
      out_kvalid);
  P3_CUDA(cudaGetLastError());
  return P3_OK;
}

}  // namespace p3
