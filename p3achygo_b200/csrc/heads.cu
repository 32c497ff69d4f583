// Policy / pass / value / score / ownership heads (python/model.py:783-812, 887-979, 691-702, 643-647)
// plus everything the C++ consumer derives per leaf: softmax of the optimistic logits
// (cc/nn/engine/trt_engine.cc:347) and the value / E[score] / Var[score] statistics of
// mcts::LeafEvaluator InitFields (cc/mcts/leaf_evaluator.cc:83-112).
//
// Input: `pgv` = the three head 1x1 convs (conv_p | conv_g | conv_v, python/model.py:784-785,889)
// evaluated as ONE GEMM over the raw trunk output, fp32 [n*400, 3*Ch] in the padded board-row layout.
// One CTA per position; HBM-bound (reads 361*3Ch*4 B, writes ~9.7 KB), warp-shuffle reductions, fp32.
#include "common.cuh"
#include "math.cuh"

namespace p3 {
namespace {

constexpr int kThreads = 256;
constexpr int kMaxCh = 64;
constexpr int kMaxCv = 128;

__device__ __forceinline__ float block_reduce(float v, bool is_max, float* s_scratch) {
  v = is_max ? warp_max(v) : warp_sum(v);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();
  if (lane == 0) s_scratch[warp] = v;
  __syncthreads();
  float r = s_scratch[0];
  for (int i = 1; i < kThreads / 32; ++i) r = is_max ? fmaxf(r, s_scratch[i]) : r + s_scratch[i];
  return r;
}

// in-place softmax over s_in[0..n) -> out[0..n) (global), accurate expf as core::Softmax (vmath.h:169-178)
__device__ void block_softmax(const float* s_in, int n, float* out, float* out2, float* s_scratch) {
  float m = -INFINITY;
  for (int i = threadIdx.x; i < n; i += kThreads) m = fmaxf(m, s_in[i]);
  m = block_reduce(m, true, s_scratch);
  float s = 0.0f;
  for (int i = threadIdx.x; i < n; i += kThreads) s += expf(s_in[i] - m);
  s = block_reduce(s, false, s_scratch);
  for (int i = threadIdx.x; i < n; i += kThreads) {
    const float p = expf(s_in[i] - m) / s;
    out[i] = p;
    if (out2) out2[i] = p;
  }
}

// kAccurate: libm-grade mish / exp (the fp32 parity engine).  The bf16 engine uses the ex2 / rcp forms (rel. err ~1e-6,
// far below the bf16 rounding of its inputs): the 800 x Cv mish evaluations of the score head dominate this kernel.
template <bool kAccurate>
__global__ void __launch_bounds__(kThreads)
heads_kernel(const float* __restrict__ pgv, HeadWeights hw, p3_infer_result* __restrict__ results,
             p3_aux_result* __restrict__ auxs) {
  __shared__ float s_logits[4][P3_MAX_MOVES];  // main, aux, soft, optimistic
  __shared__ float s_score[P3_NUM_SCORE_LOGITS];
  __shared__ float s_part[2][kThreads];
  __shared__ float s_gp[2 * kMaxCh], s_vp[2 * kMaxCh], s_pbias[kMaxCh];
  __shared__ float s_e[kMaxCv], s_g2[kMaxCv], s_base[kMaxCv], s_ws[kMaxCv];
  __shared__ float s_o[16], s_mcts[64];
  __shared__ float s_scratch[kThreads / 32];
  __shared__ float s_gamma_mult, s_gamma;

  const int b = blockIdx.x, tid = threadIdx.x;
  const int Ch = hw.Ch, Cv = hw.Cv, W3 = 3 * Ch;
  const float* base = pgv + static_cast<size_t>(b) * kRowsPerPos * W3;
  p3_infer_result& res = results[b];
  p3_aux_result& aux = auxs[b];

  // ---- pass 1: global pools.  g -> mish(BN(g)) (GlobalPoolBias, model.py:697-699), v raw (model.py:890)
  {
    const int cols = 2 * Ch, groups = kThreads / cols;
    const int col = tid % cols, grp = tid / cols;
    float sum = 0.0f, mx = -INFINITY;
    if (grp < groups) {
      const bool is_g = col < Ch;
      const float sc = is_g ? hw.gp_scale[col] : 1.0f, sh = is_g ? hw.gp_shift[col] : 0.0f;
      for (int p = grp; p < P3_NUM_BOARD_LOCS; p += groups) {
        float x = base[static_cast<size_t>(board_row(p)) * W3 + Ch + col];
        if (is_g) x = mish_f32<kAccurate>(fmaf(x, sc, sh));
        sum += x;
        mx = fmaxf(mx, x);
      }
    }
    s_part[0][tid] = sum;
    s_part[1][tid] = mx;
    __syncthreads();
    if (tid < cols) {
      float s = 0.0f, m = -INFINITY;
      for (int g = 0; g < groups; ++g) {
        s += s_part[0][g * cols + tid];
        m = fmaxf(m, s_part[1][g * cols + tid]);
      }
      const float mean = s / static_cast<float>(P3_NUM_BOARD_LOCS);
      if (tid < Ch) {  // GlobalPool: concat(mean, max) (model.py:643-647)
        s_gp[tid] = mean;
        s_gp[Ch + tid] = m;
      } else {
        s_vp[tid - Ch] = mean;
        s_vp[Ch + tid - Ch] = m;
      }
    }
    __syncthreads();
  }

  // ---- small dense layers on the pooled vectors
  if (tid < Ch) {  // g_biases = dense(g_pooled) (model.py:700)
    float a = hw.gp_dense_b[tid];
    for (int i = 0; i < 2 * Ch; ++i) a = fmaf(s_gp[i], hw.gp_dense_w[i * Ch + tid], a);
    s_pbias[tid] = a;
  }
  if (tid >= 64 && tid < 68) {  // pass logits: dense(g_pooled) - 3 (model.py:795,803,805); bias holds the -3
    const int k = tid - 64;
    float a = hw.pass_b[k];
    for (int i = 0; i < 2 * Ch; ++i) a = fmaf(s_gp[i], hw.pass_w[i * 4 + k], a);
    s_logits[k][P3_NUM_BOARD_LOCS] = a;
  }
  if (tid >= 128 && tid < 128 + Cv) {  // value embeddings (model.py:893-894, 908-909, 935)
    const int j = tid - 128;
    float e = hw.outcome_pre_b[j], g2 = hw.gamma_pre_b[j], sb = hw.score_pre_b[j];
    for (int i = 0; i < 2 * Ch; ++i) {
      const float v = s_vp[i];
      e = fmaf(v, hw.outcome_pre_w[i * Cv + j], e);
      g2 = fmaf(v, hw.gamma_pre_w[i * Cv + j], g2);
      sb = fmaf(v, hw.score_pre_w[i * Cv + j], sb);
    }
    s_e[j] = mish_f32<kAccurate>(e);
    s_g2[j] = mish_f32<kAccurate>(g2);
    s_base[j] = sb;                               // W_v . v_pooled + b : the per-position part of score_pre
    s_ws[j] = hw.score_pre_w[2 * Ch * Cv + j];    // weight of the score-bin input
  }
  __syncthreads();
  if (tid < 14) {  // outcome_q_output (model.py:895)
    float a = hw.outcome_b[tid];
    for (int j = 0; j < Cv; ++j) a = fmaf(s_e[j], hw.outcome_w[j * 14 + tid], a);
    s_o[tid] = a;
  } else if (tid >= 32 && tid < 32 + 51) {  // mcts value distribution logits (model.py:903)
    const int k = tid - 32;
    float a = hw.mcts_b[k];
    for (int j = 0; j < Cv; ++j) a = fmaf(s_e[j], hw.mcts_w[j * 51 + k], a);
    s_mcts[k] = a;
  } else if (tid == 96) {  // gamma (model.py:908-910) and its multiplier min(softplus(gamma), 10) (model.py:949-951)
    float a = hw.gamma_b[0];
    for (int j = 0; j < Cv; ++j) a = fmaf(s_g2[j], hw.gamma_w[j], a);
    s_gamma = a;
    s_gamma_mult = fminf(softplus_f32(a), 10.0f);
  }

  // ---- pass 2: per-point policy logits (model.py:787-812) and ownership (model.py:906-907)
  for (int p = tid; p < P3_NUM_BOARD_LOCS; p += kThreads) {
    const float* row = base + static_cast<size_t>(board_row(p)) * W3;
    float l0 = 0.0f, l1 = 0.0f, l2 = 0.0f, l3 = 0.0f, own = 0.0f;
    for (int c = 0; c < Ch; ++c) {
      const float a = mish_f32<kAccurate>(row[c] + s_pbias[c]);
      l0 = fmaf(a, hw.moves_w[c], l0);
      l1 = fmaf(a, hw.moves_w[Ch + c], l1);
      l2 = fmaf(a, hw.moves_w[2 * Ch + c], l2);
      l3 = fmaf(a, hw.moves_w[3 * Ch + c], l3);
      own = fmaf(row[2 * Ch + c], hw.own_w[c], own);
    }
    s_logits[0][p] = l0;
    s_logits[1][p] = l1;
    s_logits[2][p] = l2;
    s_logits[3][p] = l3;
    aux.ownership[p] = tanhf(own);
  }
  __syncthreads();

  // ---- score distribution logits (model.py:925-951), factored: mish(base + w_s * s_i) . w_out + b
  for (int i = tid; i < P3_NUM_SCORE_LOGITS; i += kThreads) {
    const float si = hw.scores[i];
    float a = hw.score_b[0];
    for (int j = 0; j < Cv; ++j) a = fmaf(mish_f32<kAccurate>(fmaf(s_ws[j], si, s_base[j])), hw.score_w[j], a);
    s_score[i] = s_gamma_mult * a;
  }
  __syncthreads();

  // ---- outputs
  for (int i = tid; i < P3_MAX_MOVES; i += kThreads) {
    res.move_logits[i] = s_logits[0][i];
    aux.pi_logits_aux[i] = s_logits[1][i];
    aux.pi_logits_soft[i] = s_logits[2][i];
    aux.pi_logits_optimistic[i] = s_logits[3][i];
  }
  for (int i = tid; i < P3_NUM_SCORE_LOGITS; i += kThreads) aux.score_logits[i] = s_score[i];
  block_softmax(s_logits[0], P3_MAX_MOVES, res.move_probs, nullptr, s_scratch);          // 01:pi
  block_softmax(s_logits[3], P3_MAX_MOVES, res.opt_move_probs, nullptr, s_scratch);      // trt_engine.cc:347
  block_softmax(s_score, P3_NUM_SCORE_LOGITS, res.score_probs, nullptr, s_scratch);      // 06:score_probs
  block_softmax(s_mcts, 51, aux.mcts_dist_probs, nullptr, s_scratch);                    // 24
  if (tid < 51) aux.mcts_dist_logits[tid] = s_mcts[tid];

  // leaf statistics over the score distribution (leaf_evaluator.cc:95-107)
  __syncthreads();
  {
    float m = -INFINITY;
    for (int i = tid; i < P3_NUM_SCORE_LOGITS; i += kThreads) m = fmaxf(m, s_score[i]);
    m = block_reduce(m, true, s_scratch);
    float z = 0.0f, e1 = 0.0f, e2 = 0.0f;
    for (int i = tid; i < P3_NUM_SCORE_LOGITS; i += kThreads) {
      const float w = expf(s_score[i] - m), s = static_cast<float>(i - 400) + 0.5f;
      z += w;
      e1 += w * s;
      e2 += w * s * s;
    }
    z = block_reduce(z, false, s_scratch);
    e1 = block_reduce(e1, false, s_scratch);
    e2 = block_reduce(e2, false, s_scratch);
    if (tid == 0) {
      const float mean = e1 / z;
      aux.score_mean = mean;
      aux.score_var = e2 / z - mean * mean;
    }
  }
  if (tid == 0) {
    // outcome = softmax(o[0:2]) (model.py:1266); [0] = loss, [1] = win (leaf_evaluator.cc:92-93)
    const float m = fmaxf(s_o[0], s_o[1]);
    const float e0 = expf(s_o[0] - m), e1 = expf(s_o[1] - m);
    res.value_probs[0] = e0 / (e0 + e1);
    res.value_probs[1] = e1 / (e0 + e1);
    aux.value = res.value_probs[1] - res.value_probs[0];
    aux.outcome_logits[0] = s_o[0];
    aux.outcome_logits[1] = s_o[1];
    aux.gamma = s_gamma;
    for (int k = 0; k < 3; ++k) {
      aux.q[k] = tanhf(s_o[2 + k]);                       // model.py:899-901
      aux.q_err[k] = 4.0f * sigmoid_f32(s_o[5 + k]);      // model.py:955-957
      aux.q_score[k] = s_o[8 + k];
      aux.q_score_err[k] = fabsf(s_o[11 + k]);            // model.py:961-963
    }
    res.err2_outcome = aux.q_err[0];                      // 12:q6_err (trt_names.h:19)
  }
}

}  // namespace

int heads_launch(const float* pgv, int n, const HeadWeights& hw, p3_infer_result* results, p3_aux_result* aux,
                 cudaStream_t stream, bool accurate) {
  if (hw.Ch > kMaxCh || hw.Cv > kMaxCv || 2 * hw.Ch > kThreads)
    return fail(P3_ERR_UNSUPPORTED, "heads: head channels / c_val too large");
  if (accurate) heads_kernel<true><<<n, kThreads, 0, stream>>>(pgv, hw, results, aux);
  else heads_kernel<false><<<n, kThreads, 0, stream>>>(pgv, hw, results, aux);
  P3_CUDA(cudaGetLastError());
  return P3_OK;
}

}  // namespace p3
