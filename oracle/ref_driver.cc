// TEST INFRASTRUCTURE — NOT PRODUCT CODE.
//
// Thin extern "C" driver around the UNMODIFIED reference sources
// (/root/reference/cc/{game,core,nn/engine/go_features.cc}), compiled where they
// lie by oracle/Makefile into oracle/_ref/libp3ref.so.  Nothing here is copied
// from the reference: it only #includes its headers and calls its functions, so
// that tests can pin oracle/features_oracle.c and the CUDA kernels against the
// reference's own implementation of
//   - Board/Game rules, liberties, ladders, legal moves   (cc/game/board.cc)
//   - symmetry transforms                                  (cc/game/symmetry.cc)
//   - feature-plane fill                                   (cc/nn/engine/go_features.cc)
//   - PCG32 / Gumbel sampling                              (cc/core/{rand,probability}.cc)
//   - scalar softmax                                       (cc/core/vmath.h:169-178)
// Two reference functions live in translation units that need abseil's
// synchronisation / the whole MCTS tree and cannot be compiled here; they are
// restated below against the reference's own classes, citing the lines followed:
//   - NNInterface::LoadBatch        cc/nn/nn_interface.cc:245-277
//   - Gumbel root top-k sampling    cc/mcts/gumbel.cc:283-321 (+ :44-47 comparator)
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
// legs may load this library.
#include <algorithm>
#include <array>
#include <cstdint>
#include <cstring>

#include "cc/constants/constants.h"
#include "cc/core/probability.h"
#include "cc/core/rand.h"
#include "cc/core/vmath.h"
#include "cc/game/board.h"
#include "cc/game/game.h"
#include "cc/game/symmetry.h"
#include "cc/nn/engine/go_features.h"

using game::Color;
using game::Game;
using game::Loc;
using game::Symmetry;

static_assert(sizeof(nn::GoFeatures) == 1860, "GoFeatures layout changed");

extern "C" {

// ---- Game / Board ------------------------------------------------------------------------
void* ref_game_new(float komi, int prohibit_pass_alive) {
  Game* g = new Game(prohibit_pass_alive != 0);
  g->SetKomi(komi);
  return g;
}
void ref_game_free(void* g) { delete static_cast<Game*>(g); }
int ref_game_play(void* gp, int i, int j, int color) {
  Game* g = static_cast<Game*>(gp);
  Loc loc{i, j};
  if (loc == game::kPassLoc) return g->Pass(static_cast<Color>(color)) ? 1 : 0;
  return g->PlayMove(loc, static_cast<Color>(color)) ? 1 : 0;
}
int ref_game_num_moves(void* gp) { return static_cast<Game*>(gp)->num_moves(); }
int ref_game_is_over(void* gp) { return static_cast<Game*>(gp)->IsGameOver() ? 1 : 0; }
void ref_game_board(void* gp, int8_t* out) {
  const auto& pos = static_cast<Game*>(gp)->board().position();
  std::memcpy(out, pos.data(), 361);
}
// Board::GetStonesWithLiberties(n), cc/game/board.cc:670-690
void ref_game_liberties(void* gp, int n, int8_t* out) {
  auto grid = static_cast<Game*>(gp)->board().GetStonesWithLiberties(n);
  std::memcpy(out, grid.data(), 361);
}
// Board::GetLadderedStones, cc/game/board.cc:692-899
void ref_game_laddered(void* gp, int8_t* out) {
  auto grid = static_cast<Game*>(gp)->board().GetLadderedStones();
  std::memcpy(out, grid.data(), 361);
}
// Game::IsValidMove over all 362 encodings, cc/game/game.cc:45-51
void ref_game_legal_mask(void* gp, int color, uint8_t* out) {
  Game* g = static_cast<Game*>(gp);
  for (int i = 0; i < constants::kMaxMovesPerPosition; ++i)
    out[i] = g->IsValidMove(i, static_cast<Color>(color)) ? 1 : 0;
}

// Game::moves() without the kMoveOffset pad (cc/game/game.cc:20,30-32), encoded as include/p3_b200.h's p3_game_derive
// expects: point 0..360 or 361 = pass, + 512 for WHITE.  Returns the number of moves.
int ref_game_moves(void* gp, int16_t* out, int cap) {
  Game& g = *static_cast<Game*>(gp);
  const int n = std::min(g.num_moves(), cap);
  for (int i = 0; i < n; ++i) {
    const game::Move mv = g.move(i);
    const int point = mv.loc == game::kPassLoc ? 361 : mv.loc.i * BOARD_LEN + mv.loc.j;
    out[i] = static_cast<int16_t>(point + (mv.color == WHITE ? 512 : 0));
  }
  return g.num_moves();
}
// Board::PlayMoveDry(loc, color).status for every board point (cc/game/board.cc:595-644): 0 kValid, 3 kLocNotEmpty,
// 4 kPassAliveRegion, 5 kSelfCapture, 6 kRepeatedPosition.
void ref_game_move_status(void* gp, int color, uint8_t* out) {
  const game::Board& b = static_cast<Game*>(gp)->board();
  for (int p = 0; p < 361; ++p)
    out[p] = static_cast<uint8_t>(b.PlayMoveDry(Loc{p / BOARD_LEN, p % BOARD_LEN}, static_cast<Color>(color)).status);
}

// NNInterface::LoadBatch restated (cc/nn/nn_interface.cc:245-277): gather the last five
// moves oldest->newest (noop pad, pass kept untransformed), apply `sym` to the board and
// the four derived grids, fill GoFeatures.
void ref_game_features(void* gp, int color_to_move, int sym_i, void* out) {
  Game& game = *static_cast<Game*>(gp);
  Symmetry sym = static_cast<Symmetry>(sym_i);
  nn::GoFeatures f;
  std::memset(&f, 0, sizeof(f));
  f.bsize = BOARD_LEN;
  f.color = static_cast<Color>(color_to_move);
  int num_moves = game.num_moves();
  for (int i = 0; i < constants::kNumLastMoves; ++i) {
    int mv_offset = num_moves - constants::kNumLastMoves + i;
    if (mv_offset < 0) {
      f.last_moves[i] = game::kNoopLoc;
    } else if (game.move(mv_offset).loc == game::kPassLoc) {
      f.last_moves[i] = game::kPassLoc;
    } else {
      f.last_moves[i] = game::ApplySymmetry(sym, game.move(mv_offset).loc, BOARD_LEN);
    }
  }
  f.board = game::ApplySymmetry(sym, game.board().position(), BOARD_LEN);
  f.stones_atari = game::ApplySymmetry(sym, game.board().GetStonesInAtari(), BOARD_LEN);
  f.stones_two_liberties = game::ApplySymmetry(sym, game.board().GetStonesWithLiberties(2), BOARD_LEN);
  f.stones_three_liberties = game::ApplySymmetry(sym, game.board().GetStonesWithLiberties(3), BOARD_LEN);
  f.stones_laddered = game::ApplySymmetry(sym, game.board().GetLadderedStones(), BOARD_LEN);
  f.komi = game.komi();
  std::memcpy(out, &f, sizeof(f));
}

// ---- feature planes: zero fill (trt_engine.cc:230-233) + nn::LoadGoFeatures ---------------
void ref_load_go_features(const void* feats, int n, int version, float* planes, float* scalars) {
  const int np = version == 0 ? constants::kNumInputFeaturePlanesV0 : constants::kNumInputFeaturePlanesV1;
  const int ns = version == 0 ? constants::kNumInputFeatureScalarsV0 : constants::kNumInputFeatureScalarsV1;
  std::array<int, 4> pshape{n, BOARD_LEN, BOARD_LEN, np};
  std::array<int, 2> fshape{n, ns};
  std::fill(planes, planes + static_cast<size_t>(n) * 361 * np, 0.0f);
  std::fill(scalars, scalars + static_cast<size_t>(n) * ns, 0.0f);
  const nn::GoFeatures* f = static_cast<const nn::GoFeatures*>(feats);
  for (int b = 0; b < n; ++b) nn::LoadGoFeatures(planes, scalars, pshape, fshape, f[b], b, version);
}

// ---- symmetry ----------------------------------------------------------------------------
int ref_transform_index(int sym, int index) {
  return game::TransformIndex(static_cast<Symmetry>(sym), index, BOARD_LEN);
}
int ref_transform_inv(int sym, int index) {
  return game::TransformInv(static_cast<Symmetry>(sym), index, BOARD_LEN);
}

// ---- PRNG / probability ------------------------------------------------------------------
void* ref_prob_new(uint64_t seed) { return new core::Probability(seed); }
void ref_prob_free(void* p) { delete static_cast<core::Probability*>(p); }
float ref_prob_gumbel(void* p) { return static_cast<core::Probability*>(p)->GumbelSample(); }
float ref_prob_uniform(void* p) { return static_cast<core::Probability*>(p)->Uniform(); }
uint32_t ref_prob_next(void* p) { return static_cast<core::Probability*>(p)->prng().next(); }
int ref_prob_rand_range(void* p, int lo, int hi) {
  return core::RandRange(static_cast<core::Probability*>(p)->prng(), lo, hi);
}
int ref_prob_random_symmetry(void* p) {
  return static_cast<int>(game::GetRandomSymmetry(static_cast<core::Probability*>(p)->prng()));
}

// core::Softmax<362>, cc/core/vmath.h:169-178 (what trt_engine.cc:347 applies to the optimistic logits)
void ref_softmax362(const float* logits, float* out) {
  core::Softmax<constants::kMaxMovesPerPosition>(logits, out);
}

// Gumbel root sampling restated from cc/mcts/gumbel.cc:283-321 using the reference's own
// Probability; `legal` is the mask Game::IsValidMove produced (and pass-disable folded in).
// Returns k_valid; writes min(k, k_valid) move encodings + their (logit + noise) scores.
int ref_gumbel_topk(void* prob, const float* move_logits, const uint8_t* legal, float noise_scaling,
                    int k, int32_t* out_moves, float* out_scores) {
  struct Info { float logit = 0, noise = 0, q = 0; int enc = -1; };
  constexpr float kSmallLogit = -10000;  // gumbel.cc:28
  core::Probability& probability = *static_cast<core::Probability*>(prob);
  Info info[constants::kMaxMovesPerPosition]{};
  int k_valid = 0;
  for (int i = 0; i < constants::kMaxMovesPerPosition; ++i) {
    if (!legal[i]) {
      info[i].logit = kSmallLogit;
      continue;
    }
    info[i].logit = move_logits[i];
    info[i].noise = noise_scaling * probability.GumbelSample();
    info[i].enc = i;
    ++k_valid;
  }
  k = std::min(k_valid, k);
  std::sort(info, info + constants::kMaxMovesPerPosition,
            [](const Info& x, const Info& y) { return x.logit + x.noise + x.q > y.logit + y.noise + y.q; });
  for (int i = 0; i < k; ++i) {
    out_moves[i] = info[i].enc;
    out_scores[i] = info[i].logit + info[i].noise;
  }
  return k_valid;
}

}  // extern "C"
