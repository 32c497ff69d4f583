/*
 * TEST INFRASTRUCTURE — NOT PRODUCT CODE.
 *
 * CPU restatement (plain C) of the integer / byte / PRNG algorithms on p3achygo's
 * leaf-evaluation hot path.  Each function cites the reference file:line it follows.
 * It is the checker the CUDA kernels are compared with on the GPU box (where
 * /root/reference does not exist).  PARITY PINNED: tests/test_oracle_vs_ref.py checks every
 * function here bit-for-bit against the reference's own sources compiled unmodified
 * (oracle/_ref/libp3ref.so) and against golden vectors generated from them
 * (tests/golden/, generator tests/golden/make_golden.py).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load this; the product library never links it.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define BOARD_LEN 19
#define NLOCS 361
#define NMOVES 362

typedef struct { int32_t i, j; } orc_loc;
typedef struct {
  int32_t bsize;
  int8_t color;
  float komi;
  int8_t board[NLOCS];
  orc_loc last_moves[5];
  int8_t stones_atari[NLOCS];
  int8_t stones_two_liberties[NLOCS];
  int8_t stones_three_liberties[NLOCS];
  int8_t stones_laddered[NLOCS];
} orc_go_features; /* nn::GoFeatures, cc/nn/engine/go_features.h:12-22 */

_Static_assert(sizeof(orc_go_features) == 1860, "GoFeatures mirror must be 1860 bytes");

/* ---- symmetry: cc/game/symmetry.cc:11-80 -------------------------------------------------- */
static int inv_(int i) { return BOARD_LEN - i - 1; }
static int flip_(int idx) { return (idx / BOARD_LEN) * BOARD_LEN + inv_(idx % BOARD_LEN); }
static int rot_(int idx, int r) { /* r: 0=90, 1=180, 2=270; symmetry.cc:20-34 */
  int i = idx / BOARD_LEN, j = idx % BOARD_LEN;
  if (r == 0) return j * BOARD_LEN + inv_(i);
  if (r == 1) return inv_(i) * BOARD_LEN + inv_(j);
  return inv_(j) * BOARD_LEN + i;
}
int orc_transform_index(int sym, int idx) { /* TransformIndex, symmetry.cc:36-57 */
  switch (sym) {
    case 0: return idx;
    case 1: return rot_(idx, 0);
    case 2: return rot_(idx, 1);
    case 3: return rot_(idx, 2);
    case 4: return flip_(idx);
    case 5: return rot_(flip_(idx), 0);
    case 6: return rot_(flip_(idx), 1);
    case 7: return rot_(flip_(idx), 2);
  }
  return idx;
}
int orc_transform_inv(int sym, int idx) { /* TransformInv, symmetry.cc:59-80 */
  switch (sym) {
    case 0: return idx;
    case 1: return rot_(idx, 2);
    case 2: return rot_(idx, 1);
    case 3: return rot_(idx, 0);
    case 4: return flip_(idx);
    case 5: return flip_(rot_(idx, 2));
    case 6: return flip_(rot_(idx, 1));
    case 7: return flip_(rot_(idx, 0));
  }
  return idx;
}
/* ApplySymmetry<T,N>, cc/game/symmetry.h:42-51: sym_grid[T(i)] = grid[i] */
void orc_apply_symmetry_i8(int sym, const int8_t* grid, int8_t* out) {
  for (int i = 0; i < NLOCS; ++i) out[orc_transform_index(sym, i)] = grid[i];
}
/* ApplyInverse<T,N>, cc/game/symmetry.h:53-62: inv_grid[Tinv(i)] = grid[i] */
void orc_apply_inverse_f32(int sym, const float* grid, float* out) {
  for (int i = 0; i < NLOCS; ++i) out[orc_transform_inv(sym, i)] = grid[i];
}

/* ---- feature planes: cc/nn/engine/go_features.cc:10-68, buf_utils.h:57-87 ----------------- */
static void fill_plane_pair(float* planes, int np, int b, int ours, int theirs, const int8_t* grid,
                            int8_t color) { /* FillPlanePair, buf_utils.h:57-76 */
  for (int p = 0; p < NLOCS; ++p) {
    int8_t c = grid[p];
    if (c == color) planes[((size_t)b * NLOCS + p) * np + ours] = 1.0f;
    else if (c == (int8_t)-color) planes[((size_t)b * NLOCS + p) * np + theirs] = 1.0f;
  }
}
/* zero fill (cc/nn/engine/trt_engine.cc:230-233) + LoadPlanes + LoadFeatures */
void orc_load_go_features(const orc_go_features* f, int n, int version, float* planes, float* scalars) {
  const int np = version == 0 ? 13 : 15, ns = version == 0 ? 7 : 8;
  memset(planes, 0, sizeof(float) * (size_t)n * NLOCS * np);
  memset(scalars, 0, sizeof(float) * (size_t)n * ns);
  for (int b = 0; b < n; ++b) {
    const orc_go_features* g = &f[b];
    fill_plane_pair(planes, np, b, 0, 1, g->board, g->color);              /* go_features.cc:12-13 */
    fill_plane_pair(planes, np, b, 7, 8, g->stones_atari, g->color);       /* :14-15 */
    fill_plane_pair(planes, np, b, 9, 10, g->stones_two_liberties, g->color);   /* :16-18 */
    fill_plane_pair(planes, np, b, 11, 12, g->stones_three_liberties, g->color); /* :19-21 */
    if (version >= 1) fill_plane_pair(planes, np, b, 13, 14, g->stones_laddered, g->color); /* :22-26 */
    for (int i = 0; i < 5; ++i) { /* :27-36 one-hot of last moves, skipping noop / pass */
      orc_loc lm = g->last_moves[i];
      if ((lm.i == -1 && lm.j == -1) || (lm.i == 19 && lm.j == 0)) continue;
      planes[((size_t)b * NLOCS + lm.i * BOARD_LEN + lm.j) * np + (i + 2)] = 1.0f;
    }
    scalars[(size_t)b * ns + (g->color == 1 ? 0 : 1)] = 1.0f;               /* :41-43 */
    for (int i = 0; i < 5; ++i) {                                           /* :44-52 pass flags */
      orc_loc lm = g->last_moves[i];
      if (lm.i == 19 && lm.j == 0) scalars[(size_t)b * ns + i + 2] = 1.0f;
    }
    if (version >= 1) /* :54-59 komi from the mover's perspective / 15 */
      scalars[(size_t)b * ns + 7] = (g->color == 1 ? -1.0f : 1.0f) * g->komi / 15.0f;
  }
}

/* ---- groups & liberties: Board::GetStonesWithLiberties, cc/game/board.cc:670-690 ----------- */
/* Restated as a flood fill from the raw position (the reference reads its incremental
 * GroupTracker; the liberty count of a group is the number of DISTINCT empty neighbours,
 * cc/game/board.cc GroupTracker::Move / CoalesceGroups).  Empty points never match. */
static const int DI[4] = {-1, 1, 0, 0}, DJ[4] = {0, 0, -1, 1};
static int group_liberties(const int8_t* board, int start, int* members, int* n_members, uint8_t* seen) {
  int8_t color = board[start];
  uint8_t lib_seen[NLOCS];
  memset(lib_seen, 0, sizeof lib_seen);
  int stack[NLOCS], sp = 0, libs = 0;
  *n_members = 0;
  stack[sp++] = start;
  seen[start] = 1;
  while (sp) {
    int p = stack[--sp];
    members[(*n_members)++] = p;
    int i = p / BOARD_LEN, j = p % BOARD_LEN;
    for (int d = 0; d < 4; ++d) {
      int ni = i + DI[d], nj = j + DJ[d];
      if (ni < 0 || ni >= BOARD_LEN || nj < 0 || nj >= BOARD_LEN) continue;
      int q = ni * BOARD_LEN + nj;
      if (board[q] == 0) {
        if (!lib_seen[q]) { lib_seen[q] = 1; ++libs; }
      } else if (board[q] == color && !seen[q]) {
        seen[q] = 1;
        stack[sp++] = q;
      }
    }
  }
  return libs;
}
void orc_stones_with_liberties(const int8_t* board, int liberties, int8_t* out) {
  uint8_t seen[NLOCS];
  int members[NLOCS], n;
  memset(seen, 0, sizeof seen);
  memset(out, 0, NLOCS);
  for (int p = 0; p < NLOCS; ++p) {
    if (board[p] == 0 || seen[p]) continue;
    int libs = group_liberties(board, p, members, &n, seen);
    if (libs == liberties)
      for (int k = 0; k < n; ++k) out[members[k]] = board[members[k]];
  }
}

/* ---- legal moves without history: Board::PlayMoveDry, cc/game/board.cc:595-644 ------------- */
/* pass legal (:596-599); point must be empty (:605); not in a host-supplied forbidden set
 * (pass-alive :607 and positional superko :636-640 need game history, so the caller passes them);
 * capture of an adjacent opponent group in atari makes the move legal (:611-613); otherwise the
 * move is illegal iff it is self-capture (:616, IsSelfCapture :901-915): no empty neighbour and
 * every adjacent own group has exactly one liberty (the point itself). */
void orc_legal_mask_nohist(const int8_t* board, int8_t color, const int8_t* forbidden, uint8_t* out) {
  int lib_of[NLOCS];
  uint8_t seen[NLOCS];
  int members[NLOCS], n;
  memset(seen, 0, sizeof seen);
  for (int p = 0; p < NLOCS; ++p) lib_of[p] = 0;
  for (int p = 0; p < NLOCS; ++p) {
    if (board[p] == 0 || seen[p]) continue;
    int libs = group_liberties(board, p, members, &n, seen);
    for (int k = 0; k < n; ++k) lib_of[members[k]] = libs;
  }
  for (int p = 0; p < NLOCS; ++p) {
    out[p] = 0;
    if (board[p] != 0) continue;
    if (forbidden && forbidden[p]) continue;
    int i = p / BOARD_LEN, j = p % BOARD_LEN, ok = 0;
    for (int d = 0; d < 4 && !ok; ++d) {
      int ni = i + DI[d], nj = j + DJ[d];
      if (ni < 0 || ni >= BOARD_LEN || nj < 0 || nj >= BOARD_LEN) continue;
      int q = ni * BOARD_LEN + nj;
      if (board[q] == 0) ok = 1;                                   /* an empty neighbour */
      else if (board[q] == (int8_t)-color && lib_of[q] == 1) ok = 1; /* captures */
      else if (board[q] == color && lib_of[q] > 1) ok = 1;          /* joins a group that keeps a liberty */
    }
    out[p] = (uint8_t)ok;
  }
  out[NLOCS] = 1;
}

/* ---- PCG32 / Probability: cc/core/rand.cc:7-71,100-121, cc/core/probability.cc:12-30 ------- */
#define PCG_MULT 6364136223846793005ULL
#define PCG_INC0 1442695040888963407ULL
uint64_t orc_prng_seed(uint64_t seed) { return seed + PCG_INC0; } /* PRng(seed), rand.cc:55-62 */
uint32_t orc_prng_next(uint64_t* state) {                          /* pcg32, rand.cc:32-43 */
  uint64_t x = *state;
  unsigned count = (unsigned)(x >> 59);
  *state = x * PCG_MULT + PCG_INC0;
  x ^= x >> 18;
  uint32_t v = (uint32_t)(x >> 27);
  return v >> count | v << (-count & 31);
}
float orc_uniform(uint64_t* state) { /* Probability::Uniform, probability.cc:17-30 */
  uint32_t x = (127u << 23) | (orc_prng_next(state) >> 9);
  float r;
  memcpy(&r, &x, sizeof r);
  return r - 1.0f;
}
float orc_gumbel(uint64_t* state) { /* Probability::GumbelSample, probability.cc:12-15 */
  float cdf = orc_uniform(state);
  return -logf(-logf(cdf));
}
int orc_rand_range(uint64_t* state, int lo, int hi) { /* RandRange, rand.cc:100-121 */
  if (lo == hi) return lo;
  uint32_t width = (uint32_t)hi - (uint32_t)lo, bit_mask = 0, shift = 1;
  while (width >> shift) { bit_mask = bit_mask << 1 | 1u; ++shift; }
  bit_mask = bit_mask << 1 | 1u;
  uint32_t r = orc_prng_next(state);
  while ((r & bit_mask) >= width) r = orc_prng_next(state);
  return (int)(r & bit_mask) + lo;
}

/* ---- Gumbel root top-k: cc/mcts/gumbel.cc:283-321 (comparator :44-47) ---------------------- */
/* Ties are broken by move index here; std::sort in the reference is unstable, so tests only
 * compare cases without ties among the top-k (ties have probability ~0 with continuous noise). */
typedef struct { float score; int enc; } orc_cand;
static int cand_greater(const void* a, const void* b) {
  const orc_cand* x = (const orc_cand*)a; const orc_cand* y = (const orc_cand*)b;
  if (x->score > y->score) return -1;
  if (x->score < y->score) return 1;
  return x->enc - y->enc;
}
int orc_gumbel_topk(uint64_t* state, const float* logits, const uint8_t* legal, float noise_scaling, int k,
                    int32_t* out_moves, float* out_scores) {
  orc_cand c[NMOVES];
  int k_valid = 0;
  for (int i = 0; i < NMOVES; ++i) {
    if (!legal[i]) { c[i].score = -10000.0f + 0.0f + 0.0f; c[i].enc = -1 - i; continue; }
    float noise = noise_scaling * orc_gumbel(state);
    c[i].score = logits[i] + noise + 0.0f; /* logit + gumbel_noise + qtransform(0) */
    c[i].enc = i;
    ++k_valid;
  }
  qsort(c, NMOVES, sizeof(orc_cand), cand_greater);
  if (k > k_valid) k = k_valid;
  for (int i = 0; i < k; ++i) { out_moves[i] = c[i].enc; out_scores[i] = c[i].score; }
  return k_valid;
}

/* ---- scalar softmax: core::Softmax<N>, cc/core/vmath.h:169-178 ----------------------------- */
void orc_softmax(int n, const float* logits, float* out) {
  float m = logits[0];
  for (int i = 1; i < n; ++i) m = logits[i] > m ? logits[i] : m;
  float s = 0.0f;
  for (int i = 0; i < n; ++i) { out[i] = expf(logits[i] - m); s += out[i]; }
  for (int i = 0; i < n; ++i) out[i] /= s;
}

/* ---- leaf statistics: InitFields, cc/mcts/leaf_evaluator.cc:83-112 -------------------------- */
void orc_init_fields(const float* value_probs, const float* score_probs, float* out3) {
  float score_est = 0.0f, score_sq_est = 0.0f;
  for (int i = 0; i < 800; ++i) {
    float s = (float)(i - 400) + .5f, p = score_probs[i];
    score_est += p * s;
    score_sq_est += p * s * s;
  }
  out3[0] = value_probs[0] * -1 + value_probs[1] * 1;
  out3[1] = score_est;
  out3[2] = score_sq_est - score_est * score_est;
}
