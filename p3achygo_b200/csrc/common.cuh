// Shared declarations of libp3b200: error plumbing, the padded board-row layout every trunk
// activation uses in HBM, and the launch signatures of the kernels in this directory.
#pragma once

#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

#include "p3_b200.h"

namespace p3 {

// ---- error plumbing ------------------------------------------------------------------------
void set_error(const std::string& msg);
int fail(int code, const std::string& msg);

#define P3_CUDA(call)                                                                         \
  do {                                                                                        \
    cudaError_t _e = (call);                                                                  \
    if (_e != cudaSuccess) {                                                                  \
      return ::p3::fail(P3_ERR_CUDA, std::string(#call) + " -> " + cudaGetErrorString(_e) +    \
                                         " (" + __FILE__ + ":" + std::to_string(__LINE__) + ")"); \
    }                                                                                         \
  } while (0)

// ---- padded board-row layout ---------------------------------------------------------------
// A trunk activation of a batch is a 2-D matrix [B * kRowsPerPos, C] (row-major, C contiguous).
// Point (r, c) of position b lives in row  b*400 + 20 + r*20 + c.  Rows 0..19 of each position
// and column 19 of every board row hold ZEROS, so a 3x3 tap (dy, dx) of a "same"-padded
// convolution is the pure row shift  dy*20 + dx : the zero rows/columns are the padding
// (applied after BN + mish, as the reference's conv(mish(BN(x))) with padding="same" requires,
// python/model.py:276-281).  361 of 400 rows are live.
constexpr int kRowsPerPos = 400;
constexpr int kRowPitch = 20;
constexpr int kRowBase = 20;

__host__ __device__ inline int board_row(int point) {  // point = r*19 + c  ->  padded row
  return kRowBase + (point / 19) * kRowPitch + (point % 19);
}
__host__ __device__ inline bool row_is_live(int q) {  // q in [0, 400)
  return q >= kRowBase && ((q - kRowBase) % kRowPitch) != 19;
}
__host__ __device__ inline int row_point(int q) {  // live padded row -> point index
  return ((q - kRowBase) / kRowPitch) * 19 + ((q - kRowBase) % kRowPitch);
}

// Position-aligned pair tiles (CTA-pair kernels: 3x3, fused boundary, stand-alone 1x1).  A pair's 256-row tile is tile k (of 3)
// of two consecutive positions, one per CTA: CTA `rank` of tile `mt` covers rows
//     [(2 * (mt / 3) + rank) * 400 + 20 + (mt % 3) * 128, + 128),
// i.e. a position is 3 x 128 = 384 rows starting at its first board row.  Its 20 leading padding rows are never loaded,
// multiplied or stored (they stay the zeros the buffers were allocated with; every other writer writes zeros there), which
// removes 4 % of the rows of a flat 256-row tiling.  The last tile of a position spills 4 rows into the next position's leading
// padding (written as zeros: not live); rows beyond the batch are zero-filled / clipped by TMA.  `rows` is a multiple of 400.
__host__ __device__ inline int pair_tile_count(int rows) { return ((rows / kRowsPerPos + 1) / 2) * 3; }
__host__ __device__ inline int pair_tile_row0(int mt, int rank) {
  return ((mt / 3) * 2 + rank) * kRowsPerPos + kRowBase + (mt % 3) * 128;
}
// Tile order of a launch: consecutive launches walk the batch in opposite directions (rev_last = m_tiles - 1, else -1), so that a
// kernel starts on the rows its predecessor wrote last - the part of that output the 126 MB L2 still holds - instead of on the
// rows it wrote first, which are long evicted (every tensor of the b12c256btl3 step at batch 1024 is 105 - 210 MB).
__host__ __device__ inline int pair_tile_row0(int mt, int rank, int rev_last) {
  return pair_tile_row0(rev_last >= 0 ? rev_last - mt : mt, rank);
}

// What a conv epilogue writes besides the raw sum.
enum ActMode : int {
  kActNone = 0,      // no activated copy
  kActMishBN = 1,    // act = mish(raw * scale[c] + shift[c])   (the NEXT layer's BN folded in)
  kActIdentity = 2,  // act = raw (cast to the operand type); input of the head convs
  kActMish = 3       // act = mish(raw); input of the broadcast mix (python/model.py:574)
};

struct ConvEpilogue {
  const void* residual = nullptr;   // [rows, cout] fp32 (fp16 if raw_f16), added to the accumulator (may alias raw_out)
  void* raw_out = nullptr;          // [rows, cout] fp32 (fp16 if raw_f16) raw sum (residual stream), or null
  bool raw_f16 = false;             // the residual stream is stored as IEEE fp16 (bf16 engine)
  bool raw_transposed = false;      // raw_out is written channel-major, [cout, rows] fp32 (head conv -> heads kernel)
  void* act_out = nullptr;          // [rows, cout] activated copy in the operand type, or null
  const float* scale = nullptr;     // [cout] next layer's folded BN (kActMishBN)
  const float* shift = nullptr;
  int act_mode = kActNone;
  bool op_f16 = false;              // tensor path: operands (input, weights, act_out) are IEEE fp16 instead of bf16
};

// ---- fp32 CUDA-core path (conv_fp32.cu) ------------------------------------------------------
// in [rows, cin] fp32; w [taps][cin][cout] fp32; tap_off[taps] row shifts.
int conv_fp32_launch(const float* in, const float* w, int rows, int cin, int cout, int taps,
                     const int* tap_off_host, const ConvEpilogue& ep, cudaStream_t stream);

// ---- tcgen05 path (conv_tc.cu) -----------------------------------------------------------------
struct TcConvPlan;  // TMA maps + tile config for one layer (opaque; built once per layer)
int tc_conv_plan_create(const __nv_bfloat16* in, const __nv_bfloat16* w, int rows, int cin, int cout,
                        int taps, const int* tap_off_host, const ConvEpilogue& ep, TcConvPlan** out);
void tc_conv_plan_destroy(TcConvPlan* plan);
bool tc_conv_plan_set_reverse(TcConvPlan* p, bool reverse);  // pair-kernel plans only (returns false otherwise)
int tc_conv_launch(const TcConvPlan* plan, cudaStream_t stream);
bool tc_conv_supported(int cin, int cout);

// fused block boundary (chain_tc.cu): x' = x + W1*in ; u = mish(BN1(x')) ; out2 = act2(W2*u), all on one CTA pair.
// in [rows, k1] bf16, w1 [n1][k1] bf16, w2 [n2][n1] bf16, residual / raw [rows, n1] fp16 (may alias), out2 [rows, n2] bf16.
struct TcChainPlan;
bool tc_chain_supported(int k1, int n1, int n2);
int tc_chain_plan_create(const __nv_bfloat16* in, const __nv_bfloat16* w1, const __nv_bfloat16* w2, int rows, int k1, int n1,
                         int n2, const void* residual_f16, void* raw_f16, const float* scale1, const float* shift1, void* out2,
                         const float* scale2, const float* shift2, int act2_mode, TcChainPlan** out, bool op_f16 = false,
                         float* tail_out_t = nullptr, int tail_out_ld = 0, int tail_n2_valid = 0);
// tail form (tail_out_t != nullptr): u = x' (identity), x' is not stored, out2 is ignored and act2(W2*u) (identity) is written
// as fp32, channel-major [tail_n2_valid, tail_out_ld] - the tower's last expand fused with the heads' 1x1 conv
void tc_chain_plan_destroy(TcChainPlan* p);
void tc_chain_plan_set_reverse(TcChainPlan* p, bool reverse);  // walk the tiles from the last position to the first
int tc_chain_launch(const TcChainPlan* p, cudaStream_t stream);

// stand-alone 1x1 trunk layers of the bf16 engine (pw_tc.cu): CTA-pair GEMM with resident weights, in-place residual boxes,
// per-quarter I/O warps.  ep.residual / ep.raw_out must be the fp16 stream (ep.raw_f16), ep.act_out bf16.
struct TcPwPlan;
bool tc_pw_supported(int k1, int n1);
int tc_pw_plan_create(const __nv_bfloat16* in, const __nv_bfloat16* w, int rows, int k1, int n1, const ConvEpilogue& ep,
                      TcPwPlan** out);
void tc_pw_plan_destroy(TcPwPlan* p);
void tc_pw_plan_set_reverse(TcPwPlan* p, bool reverse);
int tc_pw_launch(const TcPwPlan* p, cudaStream_t stream);

// ---- encode (encode.cu) --------------------------------------------------------------------------
// feats: device copy of p3_go_features[n]. planes [n,361,P] fp32, scalars [n,S] fp32,
// masks [n,361] uint16 (bit ch set <=> planes[...,ch] == 1).
// Optional extra outputs for the tensor-core first layer (init_tc.cu): the plane masks in a zero-bordered 23 x 24 grid per
// position (point (r, c) at (r + 2) * 24 + c + 2, so a 5x5 neighbourhood needs no bounds checks) and the game-state bias.
constexpr int kMaskPadW = 24, kMaskPadH = 23, kMaskPadElems = kMaskPadW * kMaskPadH;  // 552 uint16 = 1104 B (16-byte multiple)
struct EncodeExtra {
  uint16_t* masks_padded = nullptr;  // [n][kMaskPadElems], borders pre-zeroed
  const float* gs_w = nullptr;       // [S][C]
  const float* gs_b = nullptr;       // [C]
  int C = 0;
  float* gs_out = nullptr;           // [n][C]
  const int8_t* sym = nullptr;       // [n] game::Symmetry to apply to each slot's (identity-orientation) features, or null
};

// Dihedral-8 index maps of cc/game/symmetry.cc:11-80 on the 19x19 grid (0 identity, 1-3 rot90/180/270, 4 flip, 5-7 flip then rot).
__host__ __device__ inline int sym_rot(int idx, int r) {  // r: 0 = 90, 1 = 180, 2 = 270 (symmetry.cc:20-34)
  const int i = idx / 19, j = idx % 19;
  return r == 0 ? j * 19 + (18 - i) : r == 1 ? (18 - i) * 19 + (18 - j) : (18 - j) * 19 + i;
}
__host__ __device__ inline int sym_flip(int idx) { return (idx / 19) * 19 + (18 - idx % 19); }
__host__ __device__ inline int sym_transform_index(int sym, int idx) {  // TransformIndex, symmetry.cc:36-57
  if (sym == 0) return idx;
  if (sym <= 3) return sym_rot(idx, sym - 1);
  if (sym == 4) return sym_flip(idx);
  return sym_rot(sym_flip(idx), sym - 5);
}
__host__ __device__ inline int sym_transform_inv(int sym, int idx) {  // TransformInv, symmetry.cc:59-80
  if (sym == 0) return idx;
  if (sym <= 3) return sym_rot(idx, 3 - sym);
  if (sym == 4) return sym_flip(idx);
  return sym_flip(sym_rot(idx, 7 - sym));
}
int encode_launch(const p3_go_features* feats, int n, int version, float* planes, float* scalars,
                  uint16_t* masks, cudaStream_t stream, const EncodeExtra* extra = nullptr);
int liberties_launch(const int8_t* boards, int n, int8_t* out, cudaStream_t stream);
int legal_mask_launch(const int8_t* boards, const int8_t* colors, const int8_t* forbidden, int n,
                      uint8_t* out, cudaStream_t stream);

// ---- init conv (init_conv.cu) ------------------------------------------------------------------------
// 5x5 conv over the binary planes as a sparse gather-add of weight rows + game-state dense.
// wt [25][P][C] fp32, gs_w [S][C], gs_b [C]; writes raw fp32 [n*400, C] and the activated copy.
int init_conv_launch(const uint16_t* masks, const float* scalars, int n, int nplanes, int nscalars, int C,
                     const float* wt, const float* gs_w, const float* gs_b, void* raw_out, void* act_out,
                     bool act_bf16, const float* scale, const float* shift, cudaStream_t stream);  // act_bf16 => raw is fp16

// bf16-mode variant with the (bf16-rounded) weight table resident in shared memory; C <= 256.
bool init_conv_smem_supported(int nplanes, int C);
int init_conv_smem_launch(const uint16_t* masks, const float* scalars, int n, int nplanes, int nscalars, int C,
                          const __nv_bfloat16* wt_bf16, const float* gs_w, const float* gs_b, __half* raw_out,
                          __nv_bfloat16* act_out, const float* scale, const float* shift, cudaStream_t stream);

// bf16-mode tensor-core variant (init_tc.cu): implicit GEMM over the plane masks, weights repacked by init_tc_pack_weights.
bool init_tc_supported(int nplanes, int nscalars, int C);
struct InitTcPlan;
int init_tc_plan_create(const uint16_t* masks_padded, const float* gs, int n, int C, const __nv_bfloat16* w_packed,
                        __half* raw_out, __nv_bfloat16* act_out, const float* scale, const float* shift, InitTcPlan** out,
                        bool op_f16 = false);
void init_tc_plan_destroy(InitTcPlan* p);
int init_tc_launch(const InitTcPlan* p, cudaStream_t stream);

// second form (init_tc2.cu): the planes expanded once per position, the 25 taps as row-shifted views, 4-D TMA output map
bool init_tc2_supported(int nplanes, int nscalars, int C);
struct InitTc2Plan;
int init_tc2_plan_create(const uint16_t* masks_padded, const float* gs, int n, int C, const __nv_bfloat16* w_packed,
                         __half* raw_out, __nv_bfloat16* act_out, const float* scale, const float* shift, InitTc2Plan** out,
                         bool op_f16 = false);
void init_tc2_plan_destroy(InitTc2Plan* p);
int init_tc2_launch(const InitTc2Plan* p, cudaStream_t stream);

// ---- broadcast mix (broadcast.cu) -------------------------------------------------------------------
// y[b,q,c] = sum_p W[p,q] * x[b,p,c] + bias[q], then act = mish(BN(y)); x, act in the operand type.
int broadcast_launch(const void* x, const float* w, const float* bias, int n, int C, void* act_out,
                     bool bf16, const float* scale, const float* shift, cudaStream_t stream);

// tcgen05 version (broadcast_tc.cu): W^T as the K-major A operand, the NHWC activation as an MN-major B operand.
struct TcBcastPlan;
bool tc_broadcast_supported(int C);
int tc_broadcast_plan_create(const float* w_host, const float* bias_host, const void* x, void* act_out, int B, int C,
                             const float* scale, const float* shift, TcBcastPlan** out, bool op_f16 = false);
void tc_broadcast_plan_destroy(TcBcastPlan* p);
void tc_broadcast_plan_set_reverse(TcBcastPlan* p, bool reverse);
int tc_broadcast_launch(const TcBcastPlan* p, cudaStream_t stream);

// ---- heads (heads.cu) ------------------------------------------------------------------------------------
struct HeadWeights {
  int Ch, Cv;
  // policy
  const float *gp_scale, *gp_shift;        // BN of g (folded)
  const float *gp_dense_w, *gp_dense_b;    // [2Ch, Ch], [Ch]
  const float* moves_w;                    // [4][Ch]: main, aux, soft, optimistic
  const float *pass_w, *pass_b;            // [2Ch][4], [4]  (bias includes the -3)
  // value
  const float *outcome_pre_w, *outcome_pre_b;  // [2Ch, Cv]
  const float *outcome_w, *outcome_b;          // [Cv, 14]
  const float *mcts_w, *mcts_b;                // [Cv, 51]
  const float* own_w;                          // [Ch]
  const float *gamma_pre_w, *gamma_pre_b;      // [2Ch, Cv]
  const float *gamma_w, *gamma_b;              // [Cv], [1]
  const float *score_pre_w, *score_pre_b;      // [2Ch + 1, Cv]
  const float *score_w, *score_b;              // [Cv], [1]
  const float* scores;                         // [800]
};
// pgv: the head conv output, channel-major [3*Ch, n*400] fp32 (p | g | v), so that both the per-channel pooling and the
// per-point policy pass read it coalesced; results/aux device arrays of n.
// sym (optional, [n]): the symmetry each slot's input was rotated by; move_logits / move_probs / opt_move_probs come back
// un-rotated (ApplyInverse, cc/nn/nn_interface.h:263-287), the pass entry untouched.
// leafs (optional, [n]): the compact per-leaf record of mcts::LeafEvaluator InitFields (p3_leaf_result).
int heads_launch(const float* pgv, int n, const HeadWeights& hw, p3_infer_result* results, p3_aux_result* aux,
                 cudaStream_t stream, bool accurate = true, const int8_t* sym = nullptr, p3_leaf_result* leafs = nullptr);

// ---- gumbel (gumbel.cu) --------------------------------------------------------------------------------------
// ladder.cu: replay of move lists -> boards, laddered stones (board.cc:692-899), exact legal masks (board.cc:595-644); device pointers
int ladder_run(const int16_t* d_moves, const int32_t* d_num_moves, int max_moves, const int8_t* d_forbidden, const int8_t* d_colors,
               int n, int8_t* d_boards, int8_t* d_laddered, uint8_t* d_legal, int32_t* d_status, cudaStream_t stream);
struct LadderWorkspace;   // preallocated buffers of the game-record kernels for batches up to n, move lists up to max_moves
int ladder_workspace_create(int n, int max_moves, LadderWorkspace** out);
void ladder_workspace_destroy(LadderWorkspace* w);
void ladder_workspace_set_unsplit(LadderWorkspace* w, bool on);  // watchdog retry: searches stay on the warp that claimed them
// asynchronous form of ladder_run on a workspace; ev (optional) = 4 events recorded around the three kernels
int ladder_enqueue(LadderWorkspace* w, const int16_t* d_moves, const int32_t* d_num_moves, const int8_t* d_forbidden,
                   const int8_t* d_colors, int n, int8_t* d_boards, int8_t* d_laddered, uint8_t* d_legal, int32_t* d_status,
                   cudaStream_t stream, cudaEvent_t* ev);
// derived grids + move list -> the GoFeatures records of the slots whose num_moves >= 0 (NNInterface::LoadBatch, nn_interface.cc:245-277)
int assemble_features_launch(const int16_t* d_moves, const int32_t* d_num_moves, int max_moves, const int8_t* d_boards, const int8_t* d_libs,
                             const int8_t* d_laddered, int n, p3_go_features* d_feats, cudaStream_t stream);
// logits: root r reads logits + (slots ? slots[r] : r) * logit_stride floats (logit_stride = 362 for a packed [n,362] array,
// sizeof(p3_infer_result) / 4 when sampling straight from a result array); legal [n,362] packed by root, or (legal_by_slot) a per-slot array indexed like the logits.
int gumbel_launch(const float* logits, const uint8_t* legal, uint64_t* prng_state, int n, float noise_scaling,
                  int k, int32_t* out_moves, float* out_scores, int32_t* out_kvalid, cudaStream_t stream,
                  size_t logit_stride = P3_MAX_MOVES, const int32_t* slots = nullptr, bool legal_by_slot = false);

}  // namespace p3
