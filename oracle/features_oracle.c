/*
 * TEST INFRASTRUCTURE — NOT PRODUCT CODE.
 *
 * CPU restatement (plain C) of the integer / byte / PRNG algorithms on p3achygo's
 * leaf-evaluation hot path (feature fill, symmetry, liberties, legality, PCG32 / Gumbel, softmax, leaf statistics, and the
 * rules from a game record: replay, ladder reader, exact legal mask).  Each function cites the reference file:line it follows.
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
/* logf as the reference executes it.  probability.cc:12-15 calls the C library's logf; that algorithm lives in a third-party
 * dependency absent from /root/reference: GNU libc 2.39 (Ubuntu 2.39-0ubuntu8.5 in this image),
 * sysdeps/ieee754/flt-32/e_logf.c + e_logf_data.c (Szabolcs Nagy's ARM optimized-routines logf: 16-entry {1/c, log c}
 * table indexed by the top 4 mantissa bits around OFF = 0x3f330000, degree-3 polynomial in r = z/c - 1, everything in
 * double, ONE final rounding to float).  Restated here so that the oracle - and the CUDA kernel that mirrors it
 * (csrc/gumbel.cu) - do not depend on the libm of the box.  Pinned: tests/test_oracle_vs_ref.py compares it with this
 * machine's libm logf (what the compiled reference calls) on > 10^6 arguments; checked once exhaustively over all
 * 2 139 095 039 positive finite floats, with and without fused multiply-adds: 0 mismatches either way. */
static const double kLogfTab[16][2] = {
    {0x1.661ec79f8f3bep+0, -0x1.57bf7808caadep-2}, {0x1.571ed4aaf883dp+0, -0x1.2bef0a7c06ddbp-2},
    {0x1.49539f0f010bp+0, -0x1.01eae7f513a67p-2},  {0x1.3c995b0b80385p+0, -0x1.b31d8a68224e9p-3},
    {0x1.30d190c8864a5p+0, -0x1.6574f0ac07758p-3}, {0x1.25e227b0b8eap+0, -0x1.1aa2bc79c81p-3},
    {0x1.1bb4a4a1a343fp+0, -0x1.a4e76ce8c0e5ep-4}, {0x1.12358f08ae5bap+0, -0x1.1973c5a611cccp-4},
    {0x1.0953f419900a7p+0, -0x1.252f438e10c1ep-5}, {0x1p+0, 0x0p+0},
    {0x1.e608cfd9a47acp-1, 0x1.aa5aa5df25984p-5},  {0x1.ca4b31f026aap-1, 0x1.c5e53aa362eb4p-4},
    {0x1.b2036576afce6p-1, 0x1.526e57720db08p-3},  {0x1.9c2d163a1aa2dp-1, 0x1.bc2860d22477p-3},
    {0x1.886e6037841edp-1, 0x1.1058bc8a07ee1p-2},  {0x1.767dcf5534862p-1, 0x1.4043057b6ee09p-2}};
float orc_logf(float x) {
  const double A0 = -0x1.00ea348b88334p-2, A1 = 0x1.5575b0be00b6ap-2, A2 = -0x1.ffffef20a4123p-2, Ln2 = 0x1.62e42fefa39efp-1;
  uint32_t ix;
  memcpy(&ix, &x, 4);
  if (ix == 0x3f800000u) return 0.0f;
  if (ix - 0x00800000u >= 0x7f800000u - 0x00800000u) { /* zero, subnormal, negative, inf, nan */
    if (ix * 2 == 0) return -INFINITY;
    if (ix == 0x7f800000u) return x;
    if ((ix & 0x80000000u) || ix * 2 >= 0xff000000u) return NAN;
    float xs = x * 0x1p23f; /* subnormal: normalise */
    memcpy(&ix, &xs, 4);
    ix -= 23u << 23;
  }
  uint32_t tmp = ix - 0x3f330000u;
  int i = (int)((tmp >> 19) % 16);
  int k = (int32_t)tmp >> 23; /* arithmetic shift */
  uint32_t iz = ix - (tmp & 0xff800000u);
  float zf;
  memcpy(&zf, &iz, 4);
  double z = zf, r = z * kLogfTab[i][0] - 1, y0 = kLogfTab[i][1] + (double)k * Ln2, r2 = r * r;
  double y = A1 * r + A2;
  y = A0 * r2 + y;
  y = y * r2 + (y0 + r);
  return (float)y;
}
float orc_libm_logf(float x) { return logf(x); } /* this box's libm, for the pinning test only */
float orc_gumbel_from_uniform(float cdf) { return -orc_logf(-orc_logf(cdf)); }
float orc_gumbel(uint64_t* state) { /* Probability::GumbelSample, probability.cc:12-15 */
  float cdf = orc_uniform(state);
  return -orc_logf(-orc_logf(cdf));
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

/* ==== rules from a game record: replay, superko history, ladder reader, exact legal mask ====================
 * Restates Board::PlayMove / PlayMoveDry (cc/game/board.cc:536-644), Board::IsSelfCapture (:901-915) and
 * Board::GetLadderedStones with its Solver (:692-899) on plain arrays, with the recursion and the Board copies of the
 * reference kept as they are.  PARITY PINNED: tests/test_oracle_golden.py compares it with tests/golden/ladder_games.npz
 * (outputs of the compiled reference for 1297 game records, among them the positions of the reference's own ladder tests,
 * cc/game/__tests__/board_test.cc "LadderTest") and tests/test_oracle_vs_ref.py with the compiled reference directly.
 * Hashes: any injective-in-practice position hash reproduces seen_states_ membership; the reference's Zobrist table is
 * seeded from the clock (cc/game/zobrist.cc), so a splitmix64 table is used here. */
#define ORC_MAX_SEEN 2048
#define ORC_WHITE_BIT 512

typedef struct {
  int8_t at[NLOCS];
  uint64_t hash;
  uint64_t seen[ORC_MAX_SEEN]; /* Board::seen_states_ (board.h:372): copied with the board, as the reference does */
  int n_seen;
} orc_board;

static uint64_t orc_zobrist(int p, int is_white) {
  uint64_t x = ((uint64_t)p * 2 + (uint64_t)is_white + 1) * 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}

static int orc_adjacent(int p, int out[4]) {
  int n = 0, i = p / BOARD_LEN, j = p % BOARD_LEN;
  if (i > 0) out[n++] = p - BOARD_LEN;
  if (i < BOARD_LEN - 1) out[n++] = p + BOARD_LEN;
  if (j > 0) out[n++] = p - 1;
  if (j < BOARD_LEN - 1) out[n++] = p + 1;
  return n;
}

/* GroupTracker::ExpandGroup + LibertiesForGroup (board.cc:150-168): stones of p's group, its distinct liberties */
static int orc_group(const int8_t* at, int p, int* stones, int* n_stones, int* libs) {
  uint8_t mark[NLOCS];
  int stack[NLOCS], sp = 0, ns = 0, nl = 0, nb[4];
  memset(mark, 0, sizeof mark);
  const int8_t c = at[p];
  stack[sp++] = p;
  mark[p] = 1;
  while (sp) {
    const int q = stack[--sp];
    stones[ns++] = q;
    const int k = orc_adjacent(q, nb);
    for (int t = 0; t < k; ++t) {
      const int r = nb[t];
      if (mark[r]) continue;
      if (at[r] == c) {
        mark[r] = 1;
        stack[sp++] = r;
      } else if (at[r] == 0) {
        mark[r] = 1;
        if (libs) libs[nl] = r;
        ++nl;
      }
    }
  }
  *n_stones = ns;
  return nl;
}

/* Board::PlayMoveDry + PlayMove for a board point (board.cc:536-580, 595-644).  check = 0: replay of a recorded move. */
static int orc_play(orc_board* b, int p, int8_t color, int check, const int8_t* forbidden) {
  int nb[4], stones[NLOCS], ns, captured[NLOCS], nc = 0;
  if (check && (b->at[p] != 0 || (forbidden && forbidden[p]))) return 0; /* kLocNotEmpty / kPassAliveRegion */
  const int k = orc_adjacent(p, nb);
  b->at[p] = color;
  for (int t = 0; t < k; ++t) { /* GetCapturedGroups: opposing neighbours whose only liberty was p */
    const int q = nb[t];
    if (b->at[q] != -color) continue;
    int dup = 0;
    for (int c = 0; c < nc; ++c) dup |= captured[c] == q;
    if (dup) continue;
    if (orc_group(b->at, q, stones, &ns, NULL) == 0)
      for (int s = 0; s < ns; ++s) captured[nc++] = stones[s];
  }
  if (nc == 0 && check && orc_group(b->at, p, stones, &ns, NULL) == 0) { /* IsSelfCapture, board.cc:901-915 */
    b->at[p] = 0;
    return 0;
  }
  uint64_t h = b->hash ^ orc_zobrist(p, color < 0);
  for (int c = 0; c < nc; ++c) h ^= orc_zobrist(captured[c], -color < 0);
  if (check)
    for (int s = 0; s < b->n_seen; ++s)
      if (b->seen[s] == h) { /* kRepeatedPosition, board.cc:636-640 */
        b->at[p] = 0;
        return 0;
      }
  for (int c = 0; c < nc; ++c) b->at[captured[c]] = 0;
  b->hash = h;
  if (b->n_seen < ORC_MAX_SEEN) b->seen[b->n_seen++] = h;
  return 1;
}

/* Solver::Solve, board.cc:776-840.  `board` is this call's own copy. */
static int orc_solve(orc_board* board, int8_t g_color, int8_t color_to_move, int group_root, int last_move, int call_depth,
                     const int8_t* forbidden) {
  if (call_depth > 300) return 0;
  if (!orc_play(board, last_move, (int8_t)-color_to_move, 1, forbidden)) return g_color != color_to_move;
  int stones[NLOCS], ns, libs[NLOCS];
  const int liberties = orc_group(board->at, group_root, stones, &ns, libs);
  orc_board* copy = (orc_board*)malloc(sizeof(orc_board));
  int result;
  if (g_color != color_to_move) { /* try to capture (:800-812) */
    if (liberties > 2) result = 0;
    else if (liberties <= 1) result = 1;
    else {
      *copy = *board;
      result = orc_solve(copy, g_color, (int8_t)-color_to_move, group_root, libs[0], call_depth + 1, forbidden);
      if (!result) {
        *copy = *board;
        result = orc_solve(copy, g_color, (int8_t)-color_to_move, group_root, libs[1], call_depth + 1, forbidden);
      }
    }
  } else { /* try to refute (:813-839) */
    if (liberties > 1) result = 0;
    else {
      *copy = *board;
      if (!orc_solve(copy, g_color, (int8_t)-color_to_move, group_root, libs[0], call_depth + 1, forbidden)) {
        result = 0;
      } else {
        result = 1;
        /* FindSurroundingStonesInAtari (:744-770): opposing groups next to the group with one liberty; capture them */
        uint8_t done[NLOCS];
        memset(done, 0, sizeof done);
        int nb[4], s2[NLOCS], n2, l2[NLOCS];
        for (int s = 0; s < ns && result; ++s) {
          const int k = orc_adjacent(stones[s], nb);
          for (int t = 0; t < k && result; ++t) {
            const int q = nb[t];
            if (board->at[q] != -g_color || done[q]) continue;
            const int nl = orc_group(board->at, q, s2, &n2, l2);
            for (int u = 0; u < n2; ++u) done[s2[u]] = 1;
            if (nl != 1) continue;
            *copy = *board;
            if (!orc_solve(copy, g_color, (int8_t)-color_to_move, group_root, l2[0], call_depth + 1, forbidden)) result = 0;
          }
        }
      }
    }
  }
  free(copy);
  return result;
}

/* GroupTracker::BensonSolver::CalculatePassAliveRegionForColor, board.cc:246-462, on plain arrays: marks pass_alive[p] = color
 * for the stones of the surviving groups and every point of the surviving small regions. */
static void orc_benson_color(const int8_t* at, int8_t color, int8_t* pass_alive) {
  int gid[NLOCS], rid[NLOCS], stones[NLOCS], ns, nb[4];
  int n_groups = 0, n_regions = 0;
  static _Thread_local uint8_t vital[NLOCS][NLOCS], adj[NLOCS][NLOCS]; /* [region][group] */
  int g_alive[NLOCS], r_alive[NLOCS], g_vital[NLOCS];
  for (int p = 0; p < NLOCS; ++p) gid[p] = rid[p] = -1;
  for (int p = 0; p < NLOCS; ++p) { /* GetGroupMap (:278-296) */
    if (at[p] != color || gid[p] >= 0) continue;
    orc_group(at, p, stones, &ns, NULL);
    for (int s = 0; s < ns; ++s) gid[stones[s]] = n_groups;
    ++n_groups;
  }
  for (int p = 0; p < NLOCS; ++p) { /* GetRegionMap (:298-357): flood over empty / opposing points from every unseen empty point */
    if (at[p] != 0 || rid[p] != -1) continue;
    int stack[NLOCS], sp = 0, members[NLOCS], nm = 0, small = 1;
    stack[sp++] = p;
    rid[p] = -2;
    while (sp) {
      const int q = stack[--sp];
      members[nm++] = q;
      int is_liberty = at[q] != 0;
      const int k = orc_adjacent(q, nb);
      for (int t = 0; t < k; ++t) {
        if (at[nb[t]] == color) {
          if (at[q] == 0) is_liberty = 1;
          continue;
        }
        if (rid[nb[t]] == -1) {
          rid[nb[t]] = -2;
          stack[sp++] = nb[t];
        }
      }
      if (!is_liberty) small = 0;
    }
    for (int m = 0; m < nm; ++m) rid[members[m]] = small ? n_regions : -3; /* -3: seen, not small */
    if (small) ++n_regions;
  }
  for (int r = 0; r < n_regions; ++r)
    for (int g = 0; g < n_groups; ++g) vital[r][g] = 1, adj[r][g] = 0;
  for (int p = 0; p < NLOCS; ++p) { /* PopulateAdjacentRegions (:359-374), PopulateVitalRegions (:376-418) */
    if (rid[p] < 0) continue;
    uint8_t touches[NLOCS];
    memset(touches, 0, (size_t)n_groups);
    const int k = orc_adjacent(p, nb);
    for (int t = 0; t < k; ++t)
      if (gid[nb[t]] >= 0) touches[gid[nb[t]]] = 1, adj[rid[p]][gid[nb[t]]] = 1;
    if (at[p] == 0)
      for (int g = 0; g < n_groups; ++g)
        if (!touches[g]) vital[rid[p]][g] = 0;
  }
  for (int g = 0; g < n_groups; ++g) g_alive[g] = 1, g_vital[g] = 0;
  for (int r = 0; r < n_regions; ++r) {
    r_alive[r] = 1;
    for (int g = 0; g < n_groups; ++g) g_vital[g] += vital[r][g];
  }
  for (int changed = 1; changed;) { /* RunBenson (:420-462) */
    changed = 0;
    for (int g = 0; g < n_groups; ++g) {
      if (!g_alive[g] || g_vital[g] >= 2) continue;
      changed = 1;
      g_alive[g] = 0;
      for (int r = 0; r < n_regions; ++r) {
        if (!r_alive[r] || !adj[r][g]) continue;
        r_alive[r] = 0;
        for (int v = 0; v < n_groups; ++v) g_vital[v] -= vital[r][v];
      }
    }
  }
  for (int p = 0; p < NLOCS; ++p)
    if ((gid[p] >= 0 && g_alive[gid[p]]) || (rid[p] >= 0 && r_alive[rid[p]])) pass_alive[p] = color;
}

/* Replays `moves` (codes: point 0..360 or 361 = pass, + 512 for WHITE; Game::moves()) from the empty board and returns
 * board [361], laddered [361] (Board::GetLadderedStones, board.cc:842-899) and, when legal != NULL, Game::IsValidMove for
 * `color` over all 362 encodings (game.cc:45-51).  forbidden: pass-alive points (optional).  Returns 0, or 1 for an
 * impossible record. */
int orc_game_derive(const int16_t* moves, int num_moves, const int8_t* forbidden, int8_t color, int8_t* board, int8_t* laddered,
                    uint8_t* legal) {
  orc_board* b = (orc_board*)calloc(1, sizeof(orc_board));
  orc_board* copy = (orc_board*)malloc(sizeof(orc_board));
  b->seen[b->n_seen++] = 0; /* the empty board (board.cc:505-513) */
  int rc = 0, passes = 0, consecutive = 0, have_snapshot = 0;
  int8_t snapshot[NLOCS], fb[NLOCS];
  for (int m = 0; m < num_moves && !rc; ++m) {
    const int code = moves[m], p = code & (ORC_WHITE_BIT - 1);
    if (code < 0) continue;
    if (p >= NLOCS) { /* Board::Pass (board.cc:582-593): seen_states_ untouched; Benson from the third pass on unless the game ends */
      ++passes;
      ++consecutive;
      if (consecutive != 2 && passes >= 3) {
        memcpy(snapshot, b->at, NLOCS);
        have_snapshot = 1;
      }
      continue;
    }
    consecutive = 0;
    if (b->at[p] != 0) rc = 1;
    else orc_play(b, p, (code & ORC_WHITE_BIT) ? -1 : 1, 0, NULL);
  }
  memcpy(board, b->at, NLOCS);
  /* pass-alive points: the caller's grid and / or GroupTracker::CalculatePassAliveRegions (board.cc:223-233) at the last such pass */
  memset(fb, 0, NLOCS);
  if (forbidden) memcpy(fb, forbidden, NLOCS);
  if (have_snapshot) {
    orc_benson_color(snapshot, 1, fb);
    orc_benson_color(snapshot, -1, fb);
  }
  forbidden = fb;
  if (laddered) {
    memset(laddered, 0, NLOCS);
    uint8_t visited[NLOCS];
    memset(visited, 0, sizeof visited);
    int stones[NLOCS], ns, libs[NLOCS], nb[4];
    for (int p = 0; p < NLOCS && !rc; ++p) { /* groups in atari (:874-882) */
      if (b->at[p] == 0 || visited[p]) continue;
      const int nl = orc_group(b->at, p, stones, &ns, libs);
      for (int s = 0; s < ns; ++s) visited[stones[s]] = 1;
      if (nl != 1) continue;
      int empties = 0; /* IsLaddered's quick reject: GroupTracker::LibertiesAt(liberty) >= 3 (:857-860) */
      const int k = orc_adjacent(libs[0], nb);
      for (int t = 0; t < k; ++t) empties += b->at[nb[t]] == 0;
      if (empties >= 3) continue;
      const int8_t g_color = b->at[p];
      *copy = *b;
      if (orc_solve(copy, g_color, (int8_t)-g_color, p, libs[0], 0, forbidden))
        for (int s = 0; s < ns; ++s) laddered[stones[s]] = g_color;
    }
  }
  if (legal) {
    for (int p = 0; p < NLOCS; ++p) {
      *copy = *b;
      legal[p] = (uint8_t)orc_play(copy, p, color, 1, forbidden);
    }
    legal[NLOCS] = 1; /* pass (board.cc:516-518) */
  }
  free(copy);
  free(b);
  return rc;
}
