/*
 * p3_b200.h — C ABI of libp3b200.so, the B200-native (sm_100a) batched leaf
 * evaluator that sits behind p3achygo's `nn::Engine` boundary.
 *
 * Every entry point below replaces (or is what a reference-side binding would
 * call from) one member of the reference interface; the reference location is
 * cited on each declaration as `cc/...:line`.  Signatures are plain C: opaque
 * handle, POD structs, pointers and sizes.  No torch / C++ types cross this
 * boundary.  All functions return 0 on success and a non-zero code on failure;
 * `p3_last_error()` returns a thread-local description.  The reference treats
 * engine errors as fatal (CHECK / LOG(FATAL), cc/nn/engine/trt_engine.cc:27-35),
 * so the C++ adapter in INTEGRATION.md CHECKs every return code.
 *
 * There is NO CPU fallback: every compute entry point fails with
 * P3_ERR_NO_DEVICE when no CUDA device is usable.
 */
#ifndef P3_B200_H_
#define P3_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- constants: cc/constants/constants.h:36-72 ------------------------------------------ */
#define P3_BOARD_LEN 19
#define P3_NUM_BOARD_LOCS 361          /* kNumBoardLocs */
#define P3_MAX_MOVES 362               /* kMaxMovesPerPosition (361 + pass) */
#define P3_NUM_VALUE_LOGITS 2          /* kNumValueLogits */
#define P3_NUM_SCORE_LOGITS 800        /* kNumScoreLogits */
#define P3_NUM_LAST_MOVES 5            /* kNumLastMoves */
#define P3_NUM_PLANES_V1 15            /* kNumInputFeaturePlanesV1 */
#define P3_NUM_SCALARS_V1 8            /* kNumInputFeatureScalarsV1 */
#define P3_NUM_PLANES_V0 13            /* kNumInputFeaturePlanesV0 */
#define P3_NUM_SCALARS_V0 7            /* kNumInputFeatureScalarsV0 */
#define P3_BLACK 1
#define P3_WHITE (-1)
#define P3_EMPTY 0
#define P3_PASS_ENCODING 361           /* kPassMoveEncoding */

/* ---- error codes ------------------------------------------------------------------------ */
enum {
  P3_OK = 0,
  P3_ERR_INVALID_ARG = 1,
  P3_ERR_NO_DEVICE = 2,     /* no CUDA device / driver: there is no CPU fallback */
  P3_ERR_CUDA = 3,          /* a CUDA runtime / driver call failed */
  P3_ERR_IO = 4,            /* weight file missing / malformed */
  P3_ERR_UNSUPPORTED = 5    /* model shape / version not supported by the kernels */
};

/* Arithmetic the tower runs in.  Both accumulate in fp32. */
enum {
  P3_PRECISION_FP32 = 0,  /* fp32 operands, CUDA-core FFMA: parity mode (max-abs 1e-3 vs fp32 oracle) */
  P3_PRECISION_BF16 = 1,  /* bf16 operands, tcgen05.mma with fp32 TMEM accumulators: throughput mode   */
  P3_PRECISION_FP16 = 2   /* IEEE fp16 operands (tcgen05.mma kind::f16, same rate as bf16), fp32 accumulators: the reference's
                             production precision (whole-graph fp16, python/rl_loop/model_utils.py:181) with 3 more mantissa bits
                             than bf16; activations and weights beyond +-65504 saturate (p3_engine_range_check) */
};

/* ---- POD mirrors of the reference structs ------------------------------------------------ */

/* game::Loc, cc/game/loc.h:15-21.  kNoopLoc = {-1,-1}, kPassLoc = {19,0} (loc.h:46-47). */
typedef struct p3_loc {
  int32_t i;
  int32_t j;
} p3_loc;

/* nn::GoFeatures, cc/nn/engine/go_features.h:12-22 (1860 bytes; game::Color = int8_t). */
typedef struct p3_go_features {
  int32_t bsize;
  int8_t color;
  float komi;
  int8_t board[P3_NUM_BOARD_LOCS];
  p3_loc last_moves[P3_NUM_LAST_MOVES];
  int8_t stones_atari[P3_NUM_BOARD_LOCS];
  int8_t stones_two_liberties[P3_NUM_BOARD_LOCS];
  int8_t stones_three_liberties[P3_NUM_BOARD_LOCS];
  int8_t stones_laddered[P3_NUM_BOARD_LOCS];
} p3_go_features;

/* nn::NNInferResult, cc/nn/engine/engine.h:12-20 (7568 bytes, 16-byte aligned). */
typedef struct p3_infer_result {
  float move_logits[P3_MAX_MOVES];
  float move_probs[P3_MAX_MOVES];
  float value_probs[P3_NUM_VALUE_LOGITS];   /* [0] = P(loss), [1] = P(win), side to move */
  float score_probs[P3_NUM_SCORE_LOGITS];
#if defined(__cplusplus)
  alignas(16)
#else
  _Alignas(16)
#endif
  float opt_move_probs[P3_MAX_MOVES];
  float err2_outcome;
} p3_infer_result;

/* Outputs of python/model.py:1267-1293 that the reference C++ does not read
 * (cc/nn/engine/trt_names.h:9-21) but that parity tests compare. */
typedef struct p3_aux_result {
  float pi_logits_aux[P3_MAX_MOVES];      /* 08 */
  float pi_logits_soft[P3_MAX_MOVES];     /* 21 */
  float pi_logits_optimistic[P3_MAX_MOVES]; /* 22 */
  float outcome_logits[2];                /* 02 */
  float score_logits[P3_NUM_SCORE_LOGITS]; /* 05 */
  float gamma;                            /* 07 */
  float q[3];                             /* 09-11: q6, q16, q50 */
  float q_err[3];                         /* 12-14 */
  float q_score[3];                       /* 15-17 */
  float q_score_err[3];                   /* 18-20 */
  float mcts_dist_logits[51];             /* 23 */
  float mcts_dist_probs[51];              /* 24 */
  float ownership[P3_NUM_BOARD_LOCS];     /* 04 */
  /* leaf statistics of mcts::LeafEvaluator::InitFields, cc/mcts/leaf_evaluator.cc:83-112 */
  float value;                            /* p[1] - p[0] */
  float score_mean;                       /* sum p_i (i - 400 + .5) */
  float score_var;                        /* E[s^2] - E[s]^2 */
} p3_aux_result;

/* What mcts::LeafEvaluator's InitFields keeps of an NNInferResult (cc/mcts/leaf_evaluator.cc:83-112): the three policy
 * arrays it copies into the TreeNode (:85-90) and the four scalars it derives (:92-108).  4360 bytes instead of 7568: the
 * 800 score_probs the TensorRT engine ships per leaf (cc/nn/engine/trt_engine.cc:281-298) stay on the GPU. */
typedef struct p3_leaf_result {
  float move_logits[P3_MAX_MOVES];     /* TreeNode::move_logits */
  float move_probs[P3_MAX_MOVES];      /* TreeNode::move_probs */
  float opt_move_probs[P3_MAX_MOVES];  /* TreeNode::opt_probs */
  float value;                         /* init_outcome_est = p[1] - p[0] */
  float score_mean;                    /* init_score_est   = sum p_i (i - 400 + .5) */
  float score_var;                     /* init_score_var   = E[s^2] - E[s]^2 */
  float err;                           /* init_err_est     = sqrt(err2_outcome) */
} p3_leaf_result;

typedef struct p3_engine p3_engine;

/* ---- engine lifecycle: nn::CreateEngine, cc/nn/engine/engine_factory.cc:56-73 ------------ */

/* Construct an evaluator.  Replaces `TrtEngine::Create(path, batch_size, version)`
 * (cc/nn/engine/trt_engine.cc:85-166, :358-362).  `weights_path` is a flat P3W1 weight
 * file (p3achygo_b200/weights.py; tag tree of python/export_weights.py:16-90).
 * `feat_version`: 1 = 15 planes / 8 scalars (cc/nn/engine/trt_engine.cc:89-97). */
int p3_engine_create(const char* weights_path, int device, int batch_size, int feat_version,
                     int precision, p3_engine** out);
void p3_engine_destroy(p3_engine* e);

/* nn::Engine::LoadBatch, cc/nn/engine/engine.h:35 (TRT impl cc/nn/engine/trt_engine.cc:222-236).
 * Thread-safe for distinct `batch_id`; takes no lock; may overlap p3_engine_run_inference
 * (cc/nn/nn_interface.cc:276). Copies the 1860-byte game state into pinned staging. */
int p3_engine_load_batch(p3_engine* e, int batch_id, const p3_go_features* features);

/* NNInterface::LoadBatch + GetBatch symmetry handling moved onto the GPU (cc/nn/nn_interface.cc:245-277,
 * cc/nn/nn_interface.h:263-287): `features` are in the game's own (identity) orientation and `sym` is the
 * game::Symmetry (0..7, cc/game/symmetry.h) the caller would have applied.  The encode kernel applies it to the board,
 * the derived grids and the last moves (passes / no-ops untouched); the heads kernel un-applies it on the 361 board
 * entries of move_logits, move_probs and opt_move_probs, so GetBatch returns what NNInterface::GetBatch would. */
int p3_engine_load_batch_sym(p3_engine* e, int batch_id, const p3_go_features* features, int sym);

/* nn::Engine::RunInference, cc/nn/engine/engine.h:36 (cc/nn/engine/trt_engine.cc:238-304).
 * H2D(game state) -> encode -> tower -> heads -> D2H(results) -> stream sync. Always the full batch. */
int p3_engine_run_inference(p3_engine* e);

/* nn::Engine::GetBatch, cc/nn/engine/engine.h:37 (cc/nn/engine/trt_engine.cc:306-351).
 * Fills every field (opt_move_probs is the softmax of 22:pi_logits_optimistic, computed on device
 * instead of the host core::Softmax of trt_engine.cc:347). Thread-safe for distinct `batch_id`. */
int p3_engine_get_batch(p3_engine* e, int batch_id, p3_infer_result* result);

/* nn::Engine::GetOwnership, cc/nn/engine/engine.h:38-39 (TF impl cc/nn/engine/tf_engine.cc:292-299). */
int p3_engine_get_ownership(p3_engine* e, int batch_id, float own[P3_NUM_BOARD_LOCS]);

/* ---- pipelined form: two slot banks ---------------------------------------------
 * The reference serialises LoadBatch x B -> RunInference -> GetBatch x B with mu_ held across RunInference
 * (cc/nn/nn_interface.cc:286-371), so the GPU idles while workers load and read their slots and while the
 * copies run.  Here the engine owns P3_NUM_BANKS independent sets of `batch_size` slots: a caller fills one
 * bank while the other is in flight.  Bank 0 is the slot set of the serial calls above.
 *   p3_engine_load_batch_bank = nn::Engine::LoadBatch on a bank (`sym` as in p3_engine_load_batch_sym, 0 = none);
 *   p3_engine_submit          = the asynchronous half of nn::Engine::RunInference (trt_engine.cc:238-300):
 *                               H2D on a copy stream -> the captured step -> D2H on a second copy stream; returns at once;
 *   p3_engine_wait            = its stream sync (trt_engine.cc:302-304) for that bank only;
 *   p3_engine_get_batch_bank  = nn::Engine::GetBatch on a bank (valid after p3_engine_wait, until the next submit).
 * Kernels of both banks run back to back on one stream; results are identical to the serial calls.
 * A bank must be waited before it is submitted again (P3_ERR_INVALID_ARG otherwise). */
#define P3_NUM_BANKS 2
int p3_engine_load_batch_bank(p3_engine* e, int bank, int batch_id, const p3_go_features* features, int sym);
int p3_engine_submit(p3_engine* e, int bank);
int p3_engine_wait(p3_engine* e, int bank);
int p3_engine_get_batch_bank(p3_engine* e, int bank, int batch_id, p3_infer_result* result);

/* NNInterface::LoadBatch from the game record itself (cc/nn/nn_interface.cc:245-277): instead of a filled GoFeatures the
 * slot receives the game's move list (encoding of p3_game_derive below), the colour to move, komi, the optional pass-alive
 * grid and the symmetry; the next run of that bank replays the game on the GPU and derives the board, the liberty grids
 * (Board::GetStonesWithLiberties), the laddered stones (Board::GetLadderedStones) and the last five moves before the encode
 * kernel runs.  Results are bit-identical to loading the GoFeatures the reference would have built.  Works with both
 * p3_engine_run_inference (bank 0) and p3_engine_submit; slots of either kind can be mixed in one batch. */
#define P3_MAX_GAME_MOVES 1024
int p3_engine_load_game_bank(p3_engine* e, int bank, int batch_id, const int16_t* moves, int num_moves, int color, float komi,
                             const int8_t* forbidden, int sym);

/* ---- compact leaf results -------------------------------------------------------------------------
 * P3_RESULT_LEAF: runs copy back one p3_leaf_result per slot (4360 B) instead of one NNInferResult (7568 B); read them with
 * p3_engine_get_leaf / p3_engine_get_leaf_bank.  p3_engine_get_batch* fail with P3_ERR_INVALID_ARG in this mode (their
 * host copy is not refreshed).  The mode applies to runs started after the call; do not switch with a bank in flight. */
enum { P3_RESULT_FULL = 0, P3_RESULT_LEAF = 1 };
int p3_engine_set_result_mode(p3_engine* e, int mode);
int p3_engine_get_leaf(p3_engine* e, int batch_id, p3_leaf_result* leaf);                 /* after p3_engine_run_inference */
int p3_engine_get_leaf_bank(p3_engine* e, int bank, int batch_id, p3_leaf_result* leaf);  /* after p3_engine_wait(bank) */

/* Gumbel root sampling (cc/mcts/gumbel.cc:283-321, see p3_gumbel_topk below) straight from the results of `bank`'s last
 * completed run, which are still in HBM: root i samples from move_logits of slot slots[i]; only the legal masks
 * (legal [n,362], HOST; Game::IsValidMove as the caller knows it), the PRNG states and the k winners cross PCIe. */
int p3_engine_gumbel_topk_bank(p3_engine* e, int bank, const int32_t* slots, int n, const uint8_t* legal, uint64_t* prng_state,
                               float noise_scaling, int k, int32_t* out_moves, float* out_scores, int32_t* out_kvalid);

/* nn::Engine::kind()/path(), cc/nn/engine/engine.h:33-34. */
const char* p3_engine_path(const p3_engine* e);
int p3_engine_batch_size(const p3_engine* e);

/* ---- parity / measurement hooks (no reference counterpart; used by tests and bench) ------ */

/* The feature planes the encode kernel produced for `batch_id` in the last run, in the reference's
 * host layout: planes NHWC float[19*19*C] and scalars float[S] (cc/nn/engine/go_features.cc:10-68). */
int p3_engine_get_planes(p3_engine* e, int batch_id, float* planes, float* scalars);
/* The non-consumed model outputs + leaf statistics for `batch_id` of the last SERIAL run (p3_engine_run_inference /
 * p3_engine_run_device); like p3_engine_get_ownership it reads the step's own buffer, which the next run of either bank
 * overwrites.  After p3_engine_submit use the _bank forms: each bank keeps its own copy, valid from p3_engine_wait(bank)
 * until that bank's next submit.  Ownership comes back in the game's orientation (symmetry un-applied, like the policies). */
int p3_engine_get_aux(p3_engine* e, int batch_id, p3_aux_result* aux);
int p3_engine_get_aux_bank(p3_engine* e, int bank, int batch_id, p3_aux_result* aux);
int p3_engine_get_ownership_bank(p3_engine* e, int bank, int batch_id, float own[P3_NUM_BOARD_LOCS]);
/* Run encode -> tower -> heads on the game state already resident in HBM (last H2D), without any
 * host<->device copy; returns device time in ms measured with CUDA events on the engine's stream. */
int p3_engine_run_device(p3_engine* e, float* ms_total);
/* H2D of the pinned game-state staging area only (then stream sync): makes the inputs of the next
 * p3_engine_run_device resident in HBM outside its timed region. */
int p3_engine_upload(p3_engine* e);
/* One eager (non-graph) pass with a CUDA event around every launch; accumulates device ms and launch counts
 * per kernel class: [0] encode [1] init conv [2] stand-alone conv 1x1 [3] conv 3x3 [4] broadcast mix [5] head conv [6] heads
 * [7] fused block boundaries (expand 1x1 + residual + next reduce 1x1, chain_tc.cu).
 * flops[c] = algorithmic FLOPs (2*MAC) those launches performed for the whole batch;
 * bytes[c] (may be NULL) = algorithmic HBM bytes: every input / output tensor of the launch read / written once. */
#define P3_NUM_KERNEL_CLASSES 8
int p3_engine_profile(p3_engine* e, float ms[P3_NUM_KERNEL_CLASSES], int launches[P3_NUM_KERNEL_CLASSES],
                      double flops[P3_NUM_KERNEL_CLASSES], double bytes[P3_NUM_KERNEL_CLASSES]);
/* Dynamic range of the residual stream for the inputs resident in HBM (one eager pass, a scan after every launch): the tensor
 * engines keep it in IEEE fp16 and pack with cvt.rn.satfinite, so a net whose trunk exceeds +-65504 would be clamped silently.
 * max_abs = largest |x| any block left in the stream, n_saturated = values at the clamp (or NaN).  With env P3_RANGE_CHECK=1
 * every p3_engine_run_inference runs this check and fails with P3_ERR_UNSUPPORTED when n_saturated > 0 (validation mode for a
 * new checkpoint: slower, one extra pass per run).  The fp32 engine has an fp32 stream: it reports max_abs and non-finite values. */
int p3_engine_range_check(p3_engine* e, float* max_abs, long long* n_saturated);
/* Test hook for the first layer alone (python/model.py:1230-1237: 5x5 conv of the planes + dense of the game state): runs encode +
 * first layer on the inputs resident in HBM and copies its two outputs to the host, both in the padded board-row layout
 * [batch * 400, C] (point (r, c) of slot b = row b * 400 + 20 + r * 20 + c): stream_out = the residual stream x (IEEE fp16),
 * act_out = mish(BN_0(x)) in the engine's operand format (bf16 or fp16).  bytes = size of each buffer (>= batch * 400 * C * 2).
 * 16-bit engines only. */
int p3_engine_first_layer(p3_engine* e, void* stream_out, void* act_out, size_t bytes);
/* Per-stage device times of one eager pass (ms): [0] encode, [1] tower, [2] heads. */
int p3_engine_stage_ms(p3_engine* e, float ms[3]);
/* Number of kernel launches one p3_engine_run_inference issues (for bench.py's gpu_launches). */
int p3_engine_launches_per_run(const p3_engine* e);
/* Algorithmic FLOPs per position of the loaded net (2*MAC of all convs / denses). */
double p3_engine_flops_per_position(const p3_engine* e);
/* Enable (1) / disable (0) CUDA-graph replay of the device-side sequence (default on; env P3_CUDA_GRAPH=0). */
int p3_engine_set_cuda_graph(p3_engine* e, int enabled);

/* ---- stand-alone kernels ---------------------------------------------------------------- */

/* nn::LoadGoFeatures after the zero fill, cc/nn/engine/go_features.cc:62-68 + trt_engine.cc:230-233,
 * for `n` positions held in HOST memory: planes float[n,19,19,P], scalars float[n,S] in HOST memory. */
int p3_encode_features(int device, const p3_go_features* features, int n, int feat_version,
                       float* planes, float* scalars);

/* Board::GetStonesWithLiberties(1|2|3), cc/game/board.cc:670-690, from raw boards (int8 [n,361], HOST):
 * out[n,3,361] = colour of every stone whose group has exactly 1 / 2 / 3 liberties. */
int p3_board_liberties(int device, const int8_t* boards, int n, int8_t* out);

/* Legal-move mask of Game::IsValidMove without history (cc/game/board.cc:595-644 minus superko and
 * pass-alive): out[n,362] = 1 for pass, and for empty points that are not suicide for `colors[b]`.
 * `forbidden` (optional, int8 [n,361], non-zero = host-known superko / pass-alive prohibited point). */
int p3_legal_mask(int device, const int8_t* boards, const int8_t* colors, const int8_t* forbidden,
                  int n, uint8_t* out);

/* Rules from a game record, entirely on the GPU: for `n` games given as move lists
 * (game::Game::moves(), cc/game/game.h; pad moves of game.cc:20 excluded), replay them from the empty board and return
 *   boards   [n,361]  the position (Board::position(), +1 black / -1 white),
 *   laddered [n,361]  Board::GetLadderedStones(), cc/game/board.cc:692-899 (what NNInterface::LoadBatch puts in
 *                     GoFeatures::stones_laddered, cc/nn/nn_interface.cc:268-270),
 *   legal    [n,362]  Game::IsValidMove for colors[b] over all encodings, cc/game/game.cc:45-51 -> Board::PlayMoveDry,
 *                     cc/game/board.cc:595-644: occupied, pass-alive, self-capture AND positional superko,
 *   status   [n]      0 = ok, bit 0 = the move list is not a legal game, bit 1 = candidate list overflow in the reader,
 *                     bit 3 (on position 0) = the reader's watchdog fired (a warp polled ~10 s for another warp's part of a
 *                     split search): p3_game_derive, p3_engine_run_inference and p3_engine_wait then run the batch ONCE more
 *                     with splitting off (every search on the warp that claimed it; cannot stall, ~10 ms for a hard batch)
 *                     and only report the bit / P3_ERR_CUDA if that run raises it too.
 * moves [n,max_moves] int16: board point 0..360 or 361 = pass, + P3_MOVE_WHITE for a white move; entries beyond
 * num_moves[b] are ignored.  The reference's pass-alive regions (GroupTracker::BensonSolver, board.cc:246-462, computed at
 * every pass from the game's third on, board.cc:582-593, and prohibited for both colours, :607) are derived from the record
 * as well; `forbidden` (optional, [n,361], non-zero = prohibited) adds caller-known prohibited points on top.
 * boards / laddered / legal may each be NULL (legal needs colors).  All pointers are HOST memory. */
#define P3_MOVE_WHITE 512
int p3_game_derive(int device, const int16_t* moves, const int32_t* num_moves, int max_moves, const int8_t* forbidden,
                   const int8_t* colors, int n, int8_t* boards, int8_t* laddered, uint8_t* legal, int32_t* status);

/* Gumbel root sampling, cc/mcts/gumbel.cc:283-321 with core::Probability::GumbelSample
 * (cc/core/probability.cc:12-30) over PCG32 (cc/core/rand.cc:32-71), for `n` independent roots.
 *   logits [n,362], legal [n,362] (0 = masked: logit -10000, no noise, no PRNG draw),
 *   prng_state[n]: PRng state_[0] before the call; updated to the state after k_valid draws,
 *   noise_scaling, k (<= 64).
 *   out_moves [n,k] (move encodings, -1 padded), out_scores [n,k] (logit + noise), out_kvalid [n]. */
int p3_gumbel_topk(int device, const float* logits, const uint8_t* legal, uint64_t* prng_state, int n,
                   float noise_scaling, int k, int32_t* out_moves, float* out_scores, int32_t* out_kvalid);

/* One conv layer on the tcgen05 path against the fp32 CUDA-core path, for kernel unit tests:
 * x [n,361,cin] fp32 NHWC (HOST), w OIHW [cout,cin,ks,ks] fp32 (HOST), y [n,361,cout] fp32 (HOST).
 * precision selects the kernel. Applies no BN / activation. */
int p3_conv_test(int device, int precision, const float* x, const float* w, int n, int cin, int cout,
                 int ksize, float* y);

/* The broadcast mix of a BroadcastResidualBlock (python/model.py:570-581) for kernel unit tests:
 * x [n,361,C] fp32 (HOST, already activated), w [361,361] (in, out), bias [361]; y [n,361,C] = mish(W^T x + bias). */
int p3_broadcast_test(int device, int precision, const float* x, const float* w, const float* bias, int n, int C, float* y);

/* The boundary between two bottleneck blocks of the bf16 engine (python/model.py:372-427), for kernel unit tests:
 *   x'  = x + W1 t                    (expand 1x1, k1 -> n1, plus the residual stream x)
 *   u   = mish(x' * scale1 + shift1)  (folded BN of the next block's first conv)
 *   out = mish((W2 u) * scale2 + shift2)   (reduce 1x1, n1 -> n2, and the following conv's folded BN)
 * t [n,361,k1], x [n,361,n1] fp32 NHWC (HOST; rounded to bf16 / fp16 on upload), w1 [n1,k1], w2 [n2,n1] fp32,
 * scale / shift fp32 vectors; xprime [n,361,n1] (the fp16 stream) and out [n,361,n2] (bf16) come back as fp32.
 * fused != 0 runs the single fused launch (chain_tc.cu), fused == 0 the two stand-alone 1x1 launches (pw_tc.cu). */
int p3_block_boundary_test(int device, int fused, const float* t, const float* x, const float* w1, const float* w2,
                           const float* scale1, const float* shift1, const float* scale2, const float* shift2, int n,
                           int k1, int n1, int n2, float* xprime, float* out);

const char* p3_last_error(void);
const char* p3_version(void);

#ifdef __cplusplus
}
#endif
#endif /* P3_B200_H_ */
