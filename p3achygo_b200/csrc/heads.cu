// Policy / pass / value / score / ownership heads (python/model.py:783-812, 887-979, 691-702, 643-647)
// plus everything the C++ consumer derives per leaf: softmax of the optimistic logits
// (cc/nn/engine/trt_engine.cc:347) and the value / E[score] / Var[score] statistics of
// mcts::LeafEvaluator InitFields (cc/mcts/leaf_evaluator.cc:83-112).
//
// Input: `pgv` = the three head 1x1 convs (conv_p | conv_g | conv_v, python/model.py:784-785,889)
// evaluated as ONE GEMM over the raw trunk output, fp32 [n*400, 3*Ch] in the padded board-row layout.
// Persistent CTAs, 8 positions in flight per CTA (one per 128-thread group), head weights staged once in shared memory;
// reads 361*3Ch*4 B and writes ~9.7 KB per position, warp-shuffle reductions, fp32.
#include <algorithm>
#include <string>

#include "common.cuh"
#include "math.cuh"

namespace p3 {
namespace {

constexpr int kGroups = 8;            // positions in flight per CTA (148 x 8 >= 1024: the BASELINE batch is one round)
constexpr int kGT = 128;              // threads per group (one position)
constexpr int kThreads = kGroups * kGT;
constexpr int kMaxCh = 64;
constexpr int kMaxCv = 128;

__device__ __forceinline__ void group_sync(int g) { asm volatile("bar.sync %0, %1;" ::"r"(1 + g), "r"(kGT) : "memory"); }

__device__ __forceinline__ float group_reduce(float v, bool is_max, float* s_scratch, int g, int gtid) {
  v = is_max ? warp_max(v) : warp_sum(v);
  const int warp = gtid >> 5, lane = gtid & 31;
  group_sync(g);
  if (lane == 0) s_scratch[warp] = v;
  group_sync(g);
  float r = s_scratch[0];
  for (int i = 1; i < kGT / 32; ++i) r = is_max ? fmaxf(r, s_scratch[i]) : r + s_scratch[i];
  return r;
}

// softmax over s_in[0..n) -> out[0..n) (global), accurate expf as core::Softmax (vmath.h:169-178)
// sym != 0: out[i] = p[T(sym, i)] for board points (ApplyInverse, cc/game/symmetry.h:53-62), the pass entry stays in place
__device__ void group_softmax(const float* s_in, int n, float* out, float* s_scratch, int g, int gtid, int sym = 0,
                              float* out2 = nullptr) {
  float m = -INFINITY;
  for (int i = gtid; i < n; i += kGT) m = fmaxf(m, s_in[i]);
  m = group_reduce(m, true, s_scratch, g, gtid);
  float s = 0.0f;
  for (int i = gtid; i < n; i += kGT) s += expf(s_in[i] - m);
  s = group_reduce(s, false, s_scratch, g, gtid);
  for (int i = gtid; i < n; i += kGT) {
    const int src = (sym != 0 && i < P3_NUM_BOARD_LOCS) ? sym_transform_index(sym, i) : i;
    const float pr = expf(s_in[src] - m) / s;
    out[i] = pr;
    if (out2) out2[i] = pr;
  }
}

// Head weights staged once per CTA (shared by its 4 position groups): every dense layer then reads shared memory instead of
// running a chain of dependent L2 loads per position.
struct SmemWeights {
  float *gp_scale, *gp_shift, *gp_dense_w, *gp_dense_b, *moves_w, *pass_w, *pass_b, *outcome_pre_w, *outcome_pre_b, *outcome_w,
      *outcome_b, *mcts_w, *mcts_b, *own_w, *gamma_pre_w, *gamma_pre_b, *gamma_w, *gamma_b, *score_pre_w, *score_pre_b, *score_w,
      *score_b, *scores;
};
__host__ __device__ inline int heads_weight_floats(int Ch, int Cv) {
  return 2 * Ch + 2 * Ch * Ch + Ch + 4 * Ch + 2 * Ch * 4 + 4 + 2 * Ch * Cv + Cv + Cv * 14 + 14 + Cv * 51 + 51 + Ch + 2 * Ch * Cv + Cv +
         Cv + 1 + (2 * Ch + 1) * Cv + Cv + Cv + 1 + P3_NUM_SCORE_LOGITS;
}
constexpr int kGroupFloats = 4 * P3_MAX_MOVES + P3_NUM_SCORE_LOGITS + 2 * kGT + 4 * kMaxCh + kMaxCh + 4 * kMaxCv + 16 + 64 + 8 + 8;

// kAccurate: libm-grade mish / exp (the fp32 parity engine).  The bf16 engine uses the ex2 / rcp forms (rel. err ~1e-6,
// far below the bf16 rounding of its inputs): the 800 x Cv mish evaluations of the score head dominate the arithmetic.
// Persistent CTAs of 8 x 128 threads; a 128-thread group evaluates one position at a time, synchronising on its own
// named barrier (the kernel is bound by the latency of a position's serial phases, so positions in flight matter).
template <bool kAccurate>
__global__ void __launch_bounds__(kThreads, 1)
heads_kernel(const float* __restrict__ pgv, HeadWeights hw, p3_infer_result* __restrict__ results,
             p3_aux_result* __restrict__ auxs, int n, const int8_t* __restrict__ syms, p3_leaf_result* __restrict__ leafs) {
  const size_t R = static_cast<size_t>(n) * kRowsPerPos;  // pgv is channel-major: element (row, c) at pgv[c * R + row]
  extern __shared__ __align__(16) float hsm[];
  const int Ch = hw.Ch, Cv = hw.Cv;
  SmemWeights w;
  {
    float* p = hsm;
    auto take = [&](int cnt) { float* r = p; p += cnt; return r; };
    w.gp_scale = take(Ch); w.gp_shift = take(Ch); w.gp_dense_w = take(2 * Ch * Ch); w.gp_dense_b = take(Ch);
    w.moves_w = take(4 * Ch); w.pass_w = take(2 * Ch * 4); w.pass_b = take(4);
    w.outcome_pre_w = take(2 * Ch * Cv); w.outcome_pre_b = take(Cv); w.outcome_w = take(Cv * 14); w.outcome_b = take(14);
    w.mcts_w = take(Cv * 51); w.mcts_b = take(51); w.own_w = take(Ch);
    w.gamma_pre_w = take(2 * Ch * Cv); w.gamma_pre_b = take(Cv); w.gamma_w = take(Cv); w.gamma_b = take(1);
    w.score_pre_w = take((2 * Ch + 1) * Cv); w.score_pre_b = take(Cv); w.score_w = take(Cv); w.score_b = take(1);
    w.scores = take(P3_NUM_SCORE_LOGITS);
  }
  {
    auto copy = [&](float* dst, const float* src, int cnt) {
      for (int i = threadIdx.x; i < cnt; i += kThreads) dst[i] = src[i];
    };
    copy(w.gp_scale, hw.gp_scale, Ch); copy(w.gp_shift, hw.gp_shift, Ch); copy(w.gp_dense_w, hw.gp_dense_w, 2 * Ch * Ch);
    copy(w.gp_dense_b, hw.gp_dense_b, Ch); copy(w.moves_w, hw.moves_w, 4 * Ch); copy(w.pass_w, hw.pass_w, 2 * Ch * 4);
    copy(w.pass_b, hw.pass_b, 4); copy(w.outcome_pre_w, hw.outcome_pre_w, 2 * Ch * Cv); copy(w.outcome_pre_b, hw.outcome_pre_b, Cv);
    copy(w.outcome_w, hw.outcome_w, Cv * 14); copy(w.outcome_b, hw.outcome_b, 14); copy(w.mcts_w, hw.mcts_w, Cv * 51);
    copy(w.mcts_b, hw.mcts_b, 51); copy(w.own_w, hw.own_w, Ch); copy(w.gamma_pre_w, hw.gamma_pre_w, 2 * Ch * Cv);
    copy(w.gamma_pre_b, hw.gamma_pre_b, Cv); copy(w.gamma_w, hw.gamma_w, Cv); copy(w.gamma_b, hw.gamma_b, 1);
    copy(w.score_pre_w, hw.score_pre_w, (2 * Ch + 1) * Cv); copy(w.score_pre_b, hw.score_pre_b, Cv); copy(w.score_w, hw.score_w, Cv);
    copy(w.score_b, hw.score_b, 1); copy(w.scores, hw.scores, P3_NUM_SCORE_LOGITS);
  }
  __syncthreads();

  const int g = threadIdx.x / kGT, tid = threadIdx.x % kGT;
  float* gs = hsm + ((heads_weight_floats(Ch, Cv) + 3) & ~3) + g * kGroupFloats;
  float (*s_logits)[P3_MAX_MOVES] = reinterpret_cast<float (*)[P3_MAX_MOVES]>(gs);  // main, aux, soft, optimistic
  float* s_score = gs + 4 * P3_MAX_MOVES;
  float* s_part = s_score + P3_NUM_SCORE_LOGITS;  // [2][kGT]
  float* s_gp = s_part + 2 * kGT;                 // [2 kMaxCh]
  float* s_vp = s_gp + 2 * kMaxCh;                // [2 kMaxCh]
  float* s_pbias = s_vp + 2 * kMaxCh;             // [kMaxCh]
  float* s_e = s_pbias + kMaxCh;                  // [kMaxCv] x 4
  float* s_g2 = s_e + kMaxCv;
  float* s_base = s_g2 + kMaxCv;
  float* s_ws = s_base + kMaxCv;
  float* s_o = s_ws + kMaxCv;                     // [16]
  float* s_mcts = s_o + 16;                       // [64]
  float* s_scratch = s_mcts + 64;                 // [8]
  float* s_misc = s_scratch + 8;                  // gamma_mult, gamma

  for (int b = blockIdx.x * kGroups + g; b < n; b += gridDim.x * kGroups) {
    const float* base = pgv + static_cast<size_t>(b) * kRowsPerPos;  // + c * R + padded row
    p3_infer_result& res = results[b];
    p3_aux_result& aux = auxs[b];
    const int sym = syms ? syms[b] : 0;

    // ---- pass 1: global pools.  g -> mish(BN(g)) (GlobalPoolBias, model.py:697-699), v raw (model.py:890).  A warp owns a
    // channel at a time and strides over the position's padded rows (contiguous in the channel-major layout).
    {
      const int warp = tid >> 5, lane = tid & 31;
      for (int col = warp; col < 2 * Ch; col += kGT / 32) {  // col < Ch: g channel, else v channel
        const bool is_g = col < Ch;
        const float sc = is_g ? w.gp_scale[col] : 1.0f, sh = is_g ? w.gp_shift[col] : 0.0f;
        const float* src = base + static_cast<size_t>(Ch + col) * R;
        // all 12 loads of the lane are issued before the first use (the kernel is bound by load latency, not bandwidth)
        constexpr int kIters = (kRowsPerPos - kRowBase + 31) / 32;  // 12
        float xv[kIters];
#pragma unroll
        for (int k = 0; k < kIters; ++k) {
          const int q = kRowBase + lane + 32 * k;
          xv[k] = q < kRowsPerPos ? __ldg(src + q) : 0.0f;
        }
        float sum = 0.0f, mx = -INFINITY;
#pragma unroll
        for (int k = 0; k < kIters; ++k) {
          const int q = kRowBase + lane + 32 * k;
          if (q < kRowsPerPos && (q - kRowBase) % kRowPitch != 19) {  // the zero column of the layout is not a board point
            const float x = is_g ? mish_f32<kAccurate>(fmaf(xv[k], sc, sh)) : xv[k];
            sum += x;
            mx = fmaxf(mx, x);
          }
        }
        sum = warp_sum(sum);
        mx = warp_max(mx);
        if (lane == 0) {  // GlobalPool: concat(mean, max) (model.py:643-647)
          const float mean = sum / static_cast<float>(P3_NUM_BOARD_LOCS);
          if (is_g) {
            s_gp[col] = mean;
            s_gp[Ch + col] = mx;
          } else {
            s_vp[col - Ch] = mean;
            s_vp[Ch + col - Ch] = mx;
          }
        }
      }
      group_sync(g);
    }

    // ---- small dense layers on the pooled vectors: one work item per thread
    for (int idx = tid; idx < Ch + 4 + Cv; idx += kGT) {
      if (idx < Ch) {  // g_biases = dense(g_pooled) (model.py:700)
        float a = w.gp_dense_b[idx];
        for (int i = 0; i < 2 * Ch; ++i) a = fmaf(s_gp[i], w.gp_dense_w[i * Ch + idx], a);
        s_pbias[idx] = a;
      } else if (idx < Ch + 4) {  // pass logits: dense(g_pooled) - 3 (model.py:795,803,805); bias holds the -3
        const int k = idx - Ch;
        float a = w.pass_b[k];
        for (int i = 0; i < 2 * Ch; ++i) a = fmaf(s_gp[i], w.pass_w[i * 4 + k], a);
        s_logits[k][P3_NUM_BOARD_LOCS] = a;
      } else {  // value embeddings (model.py:893-894, 908-909, 935)
        const int j = idx - Ch - 4;
        float e = w.outcome_pre_b[j], g2 = w.gamma_pre_b[j], sb = w.score_pre_b[j];
        for (int i = 0; i < 2 * Ch; ++i) {
          const float v = s_vp[i];
          e = fmaf(v, w.outcome_pre_w[i * Cv + j], e);
          g2 = fmaf(v, w.gamma_pre_w[i * Cv + j], g2);
          sb = fmaf(v, w.score_pre_w[i * Cv + j], sb);
        }
        s_e[j] = mish_f32<kAccurate>(e);
        s_g2[j] = mish_f32<kAccurate>(g2);
        s_base[j] = sb;                              // W_v . v_pooled + b : the per-position part of score_pre
        s_ws[j] = w.score_pre_w[2 * Ch * Cv + j];    // weight of the score-bin input
      }
    }
    group_sync(g);
    for (int idx = tid; idx < 14 + 51 + 1; idx += kGT) {
      if (idx < 14) {  // outcome_q_output (model.py:895)
        float a = w.outcome_b[idx];
        for (int j = 0; j < Cv; ++j) a = fmaf(s_e[j], w.outcome_w[j * 14 + idx], a);
        s_o[idx] = a;
      } else if (idx < 14 + 51) {  // mcts value distribution logits (model.py:903)
        const int k = idx - 14;
        float a = w.mcts_b[k];
        for (int j = 0; j < Cv; ++j) a = fmaf(s_e[j], w.mcts_w[j * 51 + k], a);
        s_mcts[k] = a;
      } else {  // gamma (model.py:908-910) and its multiplier min(softplus(gamma), 10) (model.py:949-951)
        float a = w.gamma_b[0];
        for (int j = 0; j < Cv; ++j) a = fmaf(s_g2[j], w.gamma_w[j], a);
        s_misc[1] = a;
        s_misc[0] = fminf(softplus_f32(a), 10.0f);
      }
    }

    // ---- pass 2: per-point policy logits (model.py:787-812) and ownership (model.py:906-907); a thread owns a point, and
    // neighbouring threads read neighbouring rows of each channel
    for (int p = tid; p < P3_NUM_BOARD_LOCS; p += kGT) {
      const float* row = base + board_row(p);
      float l0 = 0.0f, l1 = 0.0f, l2 = 0.0f, l3 = 0.0f, own = 0.0f;
#pragma unroll 8
      for (int c = 0; c < Ch; ++c) {
        const float pv = __ldg(row + static_cast<size_t>(c) * R), vv = __ldg(row + static_cast<size_t>(2 * Ch + c) * R);
        const float a = mish_f32<kAccurate>(pv + s_pbias[c]);
        l0 = fmaf(a, w.moves_w[c], l0);
        l1 = fmaf(a, w.moves_w[Ch + c], l1);
        l2 = fmaf(a, w.moves_w[2 * Ch + c], l2);
        l3 = fmaf(a, w.moves_w[3 * Ch + c], l3);
        own = fmaf(vv, w.own_w[c], own);
      }
      s_logits[0][p] = l0;
      s_logits[1][p] = l1;
      s_logits[2][p] = l2;
      s_logits[3][p] = l3;
      // un-rotated like the policies (out[i] = in[T(sym, i)]): the thread of rotated point p writes original point T^-1(p)
      aux.ownership[sym != 0 ? sym_transform_inv(sym, p) : p] = tanhf(own);
    }
    group_sync(g);

    // ---- score distribution logits (model.py:925-951), factored: mish(base + w_s * s_i) . w_out + b.  A thread evaluates
    // two bins per pass (i, i + 400) so the per-channel constants are read once for both and the two chains interleave.
    {
      const float gm = s_misc[0], sb0 = w.score_b[0];
      for (int i = tid; i < P3_NUM_SCORE_LOGITS / 2; i += kGT) {
        const float si0 = w.scores[i], si1 = w.scores[i + P3_NUM_SCORE_LOGITS / 2];
        float a0 = sb0, a1 = sb0;
#pragma unroll 4
        for (int j = 0; j < Cv; ++j) {
          const float ws = s_ws[j], bs = s_base[j], wo = w.score_w[j];
          a0 = fmaf(mish_f32<kAccurate>(fmaf(ws, si0, bs)), wo, a0);
          a1 = fmaf(mish_f32<kAccurate>(fmaf(ws, si1, bs)), wo, a1);
        }
        s_score[i] = gm * a0;
        s_score[i + P3_NUM_SCORE_LOGITS / 2] = gm * a1;
      }
    }
    group_sync(g);

    // ---- outputs
    p3_leaf_result* leaf = leafs ? leafs + b : nullptr;
    for (int i = tid; i < P3_MAX_MOVES; i += kGT) {
      const float lg = s_logits[0][(sym != 0 && i < P3_NUM_BOARD_LOCS) ? sym_transform_index(sym, i) : i];
      res.move_logits[i] = lg;
      if (leaf) leaf->move_logits[i] = lg;
      aux.pi_logits_aux[i] = s_logits[1][i];
      aux.pi_logits_soft[i] = s_logits[2][i];
      aux.pi_logits_optimistic[i] = s_logits[3][i];
    }
    for (int i = tid; i < P3_NUM_SCORE_LOGITS; i += kGT) aux.score_logits[i] = s_score[i];
    group_softmax(s_logits[0], P3_MAX_MOVES, res.move_probs, s_scratch, g, tid, sym, leaf ? leaf->move_probs : nullptr);  // 01:pi
    group_softmax(s_logits[3], P3_MAX_MOVES, res.opt_move_probs, s_scratch, g, tid, sym, leaf ? leaf->opt_move_probs : nullptr);  // trt_engine.cc:347
    group_softmax(s_score, P3_NUM_SCORE_LOGITS, res.score_probs, s_scratch, g, tid);      // 06:score_probs
    group_softmax(s_mcts, 51, aux.mcts_dist_probs, s_scratch, g, tid);                    // 24
    if (tid < 51) aux.mcts_dist_logits[tid] = s_mcts[tid];

    // leaf statistics over the score distribution (leaf_evaluator.cc:95-107)
    group_sync(g);
    {
      float m = -INFINITY;
      for (int i = tid; i < P3_NUM_SCORE_LOGITS; i += kGT) m = fmaxf(m, s_score[i]);
      m = group_reduce(m, true, s_scratch, g, tid);
      float z = 0.0f, e1 = 0.0f, e2 = 0.0f;
      for (int i = tid; i < P3_NUM_SCORE_LOGITS; i += kGT) {
        const float wgt = expf(s_score[i] - m), s = static_cast<float>(i - 400) + 0.5f;
        z += wgt;
        e1 += wgt * s;
        e2 += wgt * s * s;
      }
      z = group_reduce(z, false, s_scratch, g, tid);
      e1 = group_reduce(e1, false, s_scratch, g, tid);
      e2 = group_reduce(e2, false, s_scratch, g, tid);
      if (tid == 0) {
        const float mean = e1 / z;
        aux.score_mean = mean;
        aux.score_var = e2 / z - mean * mean;
        if (leaf) {
          leaf->score_mean = mean;
          leaf->score_var = e2 / z - mean * mean;
        }
      }
    }
    if (tid == 0) {
      // outcome = softmax(o[0:2]) (model.py:1266); [0] = loss, [1] = win (leaf_evaluator.cc:92-93)
      const float m = fmaxf(s_o[0], s_o[1]);
      const float e0 = expf(s_o[0] - m), e1 = expf(s_o[1] - m);
      const float p_loss = e0 / (e0 + e1), p_win = e1 / (e0 + e1);  // (`res` may be mapped host memory: never read it back)
      res.value_probs[0] = p_loss;
      res.value_probs[1] = p_win;
      aux.value = p_win - p_loss;
      if (leaf) {
        leaf->value = p_win - p_loss;                                  // init_outcome_est, leaf_evaluator.cc:92-93
        leaf->err = sqrtf(4.0f * sigmoid_f32(s_o[5]));                 // init_err_est = sqrt(err2_outcome), :108
      }
      aux.outcome_logits[0] = s_o[0];
      aux.outcome_logits[1] = s_o[1];
      aux.gamma = s_misc[1];
      for (int k = 0; k < 3; ++k) {
        aux.q[k] = tanhf(s_o[2 + k]);                       // model.py:899-901
        aux.q_err[k] = 4.0f * sigmoid_f32(s_o[5 + k]);      // model.py:955-957
        aux.q_score[k] = s_o[8 + k];
        aux.q_score_err[k] = fabsf(s_o[11 + k]);            // model.py:961-963
      }
      res.err2_outcome = 4.0f * sigmoid_f32(s_o[5]);        // 12:q6_err (trt_names.h:19)
    }
    group_sync(g);  // the group's shared arrays are reused by its next position
  }
}

}  // namespace

int heads_launch(const float* pgv, int n, const HeadWeights& hw, p3_infer_result* results, p3_aux_result* aux,
                 cudaStream_t stream, bool accurate, const int8_t* sym, p3_leaf_result* leafs) {
  if (hw.Ch > kMaxCh || hw.Cv > kMaxCv || 2 * hw.Ch > kGT || hw.Ch % 4 != 0)
    return fail(P3_ERR_UNSUPPORTED, "heads: head channels / c_val not supported");
  const size_t smem = (static_cast<size_t>((heads_weight_floats(hw.Ch, hw.Cv) + 3) & ~3) + static_cast<size_t>(kGroups) * kGroupFloats) * sizeof(float);
  if (smem > 227 * 1024) return fail(P3_ERR_UNSUPPORTED, "heads: weights do not fit in shared memory");
  cudaError_t e = accurate ? cudaFuncSetAttribute(heads_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem))
                           : cudaFuncSetAttribute(heads_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  if (e != cudaSuccess) return fail(P3_ERR_CUDA, std::string("heads smem attribute: ") + cudaGetErrorString(e));
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int grid = std::min(sms, (n + kGroups - 1) / kGroups);
  if (accurate) heads_kernel<true><<<grid, kThreads, smem, stream>>>(pgv, hw, results, aux, n, sym, leafs);
  else heads_kernel<false><<<grid, kThreads, smem, stream>>>(pgv, hw, results, aux, n, sym, leafs);
  P3_CUDA(cudaGetLastError());
  return P3_OK;
}

}  // namespace p3
