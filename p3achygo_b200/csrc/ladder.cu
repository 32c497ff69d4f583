// Go rules on the GPU from a game's move list: position replay with positional-superko history, the ladder reader
// (Board::GetLadderedStones, cc/game/board.cc:692-899) and the exact legal-move mask (Board::PlayMoveDry,
// cc/game/board.cc:595-644), so the derived feature grids no longer have to come from the host.
//
// Representation: ONE WARP PER BOARD.  Lane r (0..18) holds row r of the black stones and of the white stones as 19-bit
// masks in two registers; lanes 19..31 hold zeros.  Everything the rules need is then a handful of warp-uniform bit
// operations:
//   * the 4-neighbourhood of a point set is (x << 1 | x >> 1) within a row plus one shuffle up and one shuffle down;
//   * a group is the fix point of  x |= nbrs(x) & colour  (flood), its liberties are nbrs(group) & empty;
//   * the stones a move captures are the part of the adjacent opposing groups that cannot reach an empty point once the
//     stone stands; a move without captures whose own group then has no liberty is self-capture
//     (Board::IsSelfCapture, board.cc:901-915, is exactly that condition);
//   * positional superko: a 64-bit Zobrist hash over (point, colour), updated by the played stone and the captured
//     stones, looked up in the list of the game's earlier positions plus the hashes on the current search path
//     (Board::seen_states_ travels with every Board copy the reference's Solve makes, board.cc:794-798).
// The history list comes from replaying the game's moves from the empty board (the reference's hashes are clock-seeded
// per process, so they could not be passed in anyway).  Points the reference prohibits as pass-alive regions (Benson,
// only computed after three passes, board.cc:587-590) are supplied by the caller as a grid, as for p3_legal_mask.
//
// The reader itself is the reference's recursion turned into an explicit stack of frames in global memory (one stack per
// resident warp, 301 frames = the reference's call_depth <= 300): an attacker node is an OR over the two liberties, a
// defender node an AND over "extend at the liberty" and "capture an adjacent group in atari"; an illegal move loses for
// its mover; results do not depend on the order children are tried (no depth-capped branch is reachable on 19 x 19), so
// duplicate candidate points are tried once.  Searches (one per group in atari that passes the reference's quick reject,
// board.cc:857-860) are handed to persistent warps through an atomic counter, since their lengths vary wildly.
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

#include "common.cuh"

namespace p3 {
namespace {

constexpr uint32_t kRowFull = 0x7FFFFu;   // 19 columns
constexpr int kMaxDepth = 301;            // frames 0..300 (Solve returns false beyond call_depth 300, board.cc:778-780)
constexpr int kMaxCand = 40;              // candidate moves of one node (1 + adjacent groups in atari)
constexpr int kWhiteBit = 512;            // move encoding: point (0..360) or 361 = pass, + 512 for WHITE
constexpr unsigned kAll = 0xFFFFFFFFu;

struct Board {
  uint32_t bk, wh;  // this lane's row
};

__device__ __forceinline__ uint64_t zobrist(int point, int is_white) {  // splitmix64 of (point, colour)
  uint64_t x = (static_cast<uint64_t>(point) * 2 + is_white + 1) * 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}

__device__ __forceinline__ uint32_t row_mask(int lane) { return lane < P3_BOARD_LEN ? kRowFull : 0u; }

__device__ __forceinline__ uint32_t nbrs(uint32_t x, int lane) {
  uint32_t up = __shfl_up_sync(kAll, x, 1);      // row above (lane - 1)
  uint32_t dn = __shfl_down_sync(kAll, x, 1);    // row below (lane + 1); lanes >= 19 hold zeros
  if (lane == 0) up = 0;
  if (lane == 31) dn = 0;
  return ((x << 1) | (x >> 1) | up | dn) & row_mask(lane);
}

// connected closure of `seed` inside `mask`
__device__ __forceinline__ uint32_t flood(uint32_t seed, uint32_t mask, int lane) {
  uint32_t x = seed & mask;
  while (true) {
    uint32_t y = x | (((x << 1) | (x >> 1)) & mask);   // two horizontal steps per vertical one: rows are cheap
    y |= ((y << 1) | (y >> 1)) & mask;
    y |= nbrs(y, lane) & mask;
    const bool changed = __any_sync(kAll, y != x);
    x = y;
    if (!changed) break;
  }
  return x;
}

__device__ __forceinline__ int warp_count(uint32_t x) { return static_cast<int>(__reduce_add_sync(kAll, static_cast<unsigned>(__popc(x)))); }

__device__ __forceinline__ uint64_t warp_xor64(uint64_t h) {
  const uint32_t lo = __reduce_xor_sync(kAll, static_cast<uint32_t>(h));
  const uint32_t hi = __reduce_xor_sync(kAll, static_cast<uint32_t>(h >> 32));
  return (static_cast<uint64_t>(hi) << 32) | lo;
}

// first point (row-major) of a non-empty row-mask set, or -1; warp-uniform
__device__ __forceinline__ int first_point(uint32_t x) {
  const uint32_t b = __ballot_sync(kAll, x != 0);
  if (!b) return -1;
  const int src = __ffs(b) - 1;
  const uint32_t row = __shfl_sync(kAll, x, src);
  return src * P3_BOARD_LEN + (__ffs(row) - 1);
}

__device__ __forceinline__ uint32_t point_bit(int point, int lane) {
  return lane == point / P3_BOARD_LEN ? 1u << (point % P3_BOARD_LEN) : 0u;
}

__device__ __forceinline__ uint64_t hash_of(uint32_t rows, int is_white, int lane) {
  uint64_t h = 0;
  while (rows) {
    const int c = __ffs(rows) - 1;
    rows &= rows - 1;
    h ^= zobrist(lane * P3_BOARD_LEN + c, is_white);
  }
  return warp_xor64(h);
}

struct SeenSet {
  const uint64_t* hist;  // hashes of the game's positions up to and including the current one
  int n_hist;
  const uint64_t* path;  // hashes of the search path's positions
  int n_path;
};

__device__ __forceinline__ bool seen_contains(const SeenSet& s, uint64_t h, int lane) {
  bool hit = false;
  for (int i = lane; i < s.n_hist; i += 32) hit |= s.hist[i] == h;
  for (int i = lane; i < s.n_path; i += 32) hit |= s.path[i] == h;
  return __any_sync(kAll, hit);
}

// Board::PlayMove (board.cc:536-580) for a board point.  `check` = false replays a recorded (legal) move: captures only.
// Returns false (board untouched) when the reference's PlayMoveDry would not return kValid.
__device__ __forceinline__ bool play(Board& b, uint64_t& hash, int point, int color, bool check, uint32_t forbidden_row,
                                     const SeenSet& seen, int lane) {
  const uint32_t bit = point_bit(point, lane);
  if (check && __any_sync(kAll, (bit & (b.bk | b.wh | forbidden_row)) != 0)) return false;  // kLocNotEmpty / kPassAliveRegion
  const bool black = color == P3_BLACK;
  uint32_t own = (black ? b.bk : b.wh) | bit;
  uint32_t opp = black ? b.wh : b.bk;
  uint32_t empty = ~(own | opp) & row_mask(lane);
  uint32_t captured = 0;
  const uint32_t seeds = nbrs(bit, lane) & opp;
  if (__any_sync(kAll, seeds != 0)) {
    const uint32_t touched = flood(seeds, opp, lane);                      // the adjacent opposing groups
    const uint32_t alive = flood(nbrs(empty, lane) & touched, touched, lane);  // ... that still reach an empty point
    captured = touched & ~alive;
  }
  const bool any_capture = __any_sync(kAll, captured != 0);
  if (any_capture) {
    opp &= ~captured;
  } else if (check) {  // IsSelfCapture, board.cc:901-915
    const uint32_t group = flood(bit, own, lane);
    if (!__any_sync(kAll, (nbrs(group, lane) & empty) != 0)) return false;
  }
  uint64_t h = hash ^ zobrist(point, black ? 0 : 1);
  if (any_capture) h ^= hash_of(captured, black ? 1 : 0, lane);
  if (check && seen_contains(seen, h, lane)) return false;                 // kRepeatedPosition, board.cc:636-640
  b.bk = black ? own : opp;
  b.wh = black ? opp : own;
  hash = h;
  return true;
}

// ---- kernel 1: replay, atari groups, quick reject -> search tasks ------------------------------------------------------
struct LadderTask {
  int pos, root, liberty;
};

__global__ void __launch_bounds__(128) replay_kernel(const int16_t* __restrict__ moves, const int32_t* __restrict__ num_moves,
                                                     int max_moves, const int8_t* __restrict__ forbidden, int n,
                                                     uint32_t* __restrict__ rows,        // [n][3][32] black, white, forbidden
                                                     uint64_t* __restrict__ hist,        // [n][max_moves + 1]
                                                     int32_t* __restrict__ n_hist, int8_t* __restrict__ boards,
                                                     int8_t* __restrict__ laddered, int32_t* __restrict__ status,
                                                     LadderTask* __restrict__ tasks, int* __restrict__ n_tasks) {
  const int lane = threadIdx.x & 31;
  const int pos = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (pos >= n) return;
  Board b{0, 0};
  uint64_t hash = 0;
  uint64_t* my_hist = hist + static_cast<size_t>(pos) * (max_moves + 1);
  int nh = 0;
  if (lane == 0) my_hist[0] = 0;  // the empty board (Board::Board inserts its hash, board.cc:505-513)
  nh = 1;
  int st = 0;
  const int nm = min(max(num_moves[pos], 0), max_moves);
  const int16_t* mv = moves + static_cast<size_t>(pos) * max_moves;
  const SeenSet none{nullptr, 0, nullptr, 0};
  for (int m = 0; m < nm; ++m) {
    const int code = mv[m];
    const int point = code & (kWhiteBit - 1);
    if (code < 0 || point >= P3_PASS_ENCODING) continue;  // pass (or padding): Board::Pass leaves seen_states_ alone
    const int color = (code & kWhiteBit) ? P3_WHITE : P3_BLACK;
    if (__any_sync(kAll, (point_bit(point, lane) & (b.bk | b.wh)) != 0)) {
      st = 1;  // not a legal game record
      break;
    }
    play(b, hash, point, color, false, 0, none, lane);
    if (lane == 0) my_hist[nh] = hash;
    ++nh;
  }
  // forbidden grid -> row masks
  uint32_t fb = 0;
  if (forbidden && lane < P3_BOARD_LEN) {
    const int8_t* f = forbidden + static_cast<size_t>(pos) * P3_NUM_BOARD_LOCS + lane * P3_BOARD_LEN;
    for (int c = 0; c < P3_BOARD_LEN; ++c) fb |= f[c] ? 1u << c : 0u;
  }
  uint32_t* r = rows + static_cast<size_t>(pos) * 96;
  r[lane] = b.bk;
  r[32 + lane] = b.wh;
  r[64 + lane] = fb;
  if (lane == 0) {
    n_hist[pos] = nh;
    status[pos] = st;
  }
  if (lane < P3_BOARD_LEN) {
    const size_t at = static_cast<size_t>(pos) * P3_NUM_BOARD_LOCS + lane * P3_BOARD_LEN;
    for (int c = 0; c < P3_BOARD_LEN; ++c) {
      if (boards) boards[at + c] = (b.bk >> c) & 1 ? P3_BLACK : ((b.wh >> c) & 1 ? P3_WHITE : P3_EMPTY);
      if (laddered) laddered[at + c] = 0;
    }
  }
  if (!tasks) return;
  // groups in atari (board.cc:874-882) that survive the quick reject (board.cc:857-860)
  const uint32_t empty = ~(b.bk | b.wh) & row_mask(lane);
  // only groups next to an empty point with at most 2 empty neighbours can be laddered candidates; enumerate the
  // groups touching any empty point's neighbourhood instead of all groups: every group has a liberty, so that is all of
  // them - but visit each once
  uint32_t todo = b.bk | b.wh;
  while (true) {
    const int s = first_point(todo);
    if (s < 0) break;
    const uint32_t sbit = point_bit(s, lane);
    const bool is_black = __any_sync(kAll, (sbit & b.bk) != 0);
    const uint32_t group = flood(sbit, is_black ? b.bk : b.wh, lane);
    todo &= ~group;
    const uint32_t libs = nbrs(group, lane) & empty;
    if (warp_count(libs) != 1) continue;
    const int liberty = first_point(libs);
    if (warp_count(nbrs(point_bit(liberty, lane), lane) & empty) >= 3) continue;  // GroupTracker::LibertiesAt(liberty) >= 3
    if (lane == 0) {
      const int t = atomicAdd(n_tasks, 1);
      tasks[t] = LadderTask{pos, s, liberty};
    }
  }
}

// ---- kernel 2: the reader ---------------------------------------------------------------------------------------------
// Per resident warp, in shared memory: the game's position hashes (superko look-ups), the hashes on the search path, and
// the first kSmemFrames frames of the explicit stack; deeper frames live in a global scratch area.  A frame is three
// 32-word rows: [0] black rows in lanes 0..18, candidate moves packed two per word in lanes 19..31; [1] white rows, lane
// 31 = n_cand | next << 8; [2] the rows of the hunted group at this node (it only grows along a path, so the flood
// after a move starts from it and converges in a step or two).  A node's kind needs no storage: the defender moves at
// even call depths, so frame d is an AND node iff d is odd.
constexpr int kSmemFrames = 40;
constexpr int kHistSmem = 608;
constexpr int kReaderWarps = 4;
constexpr int kFrameWords = 96;
constexpr int kMaxCandPacked = 26;
constexpr size_t kWarpSmemBytes = kHistSmem * 8 + kMaxDepth * 8 + kSmemFrames * kFrameWords * 4;
static_assert(kMaxCandPacked <= kMaxCand, "candidate capacity");

struct DeepFrames {  // per resident warp: frames kSmemFrames .. kMaxDepth-1
  uint32_t w[kMaxDepth - kSmemFrames][kFrameWords];
};

// like play() for the searched moves, with the cheap exits that make most nodes flood-free: a neighbouring opposing stone
// with an empty neighbour of its own cannot be captured, and a stone with an empty neighbour (or next to a friendly stone
// that has one) is not self-capture.
__device__ __forceinline__ bool play_checked(Board& b, uint64_t& hash, int point, int color, uint32_t forbidden_row,
                                             const uint64_t* hist_s, int n_hist_s, const uint64_t* hist_g, int n_hist,
                                             const uint64_t* path_s, int n_path, int lane) {
  const uint32_t bit = point_bit(point, lane);
  if (__any_sync(kAll, (bit & (b.bk | b.wh | forbidden_row)) != 0)) return false;  // kLocNotEmpty / kPassAliveRegion
  const bool black = color == P3_BLACK;
  const uint32_t own_prev = black ? b.bk : b.wh;
  const uint32_t own = own_prev | bit;
  uint32_t opp = black ? b.wh : b.bk;
  const uint32_t empty = ~(own | opp) & row_mask(lane);
  const uint32_t near_empty = nbrs(empty, lane);     // points with an empty neighbour
  const uint32_t around = nbrs(bit, lane);
  uint32_t captured = 0;
  bool any_capture = false;
  const uint32_t seeds = around & opp & ~near_empty;  // adjacent opposing stones without a liberty of their own
  if (__any_sync(kAll, seeds != 0)) {
    const uint32_t touched = flood(seeds, opp, lane);
    const uint32_t alive = flood(near_empty & touched, touched, lane);
    captured = touched & ~alive;
    any_capture = __any_sync(kAll, captured != 0);
  }
  if (any_capture) {
    opp &= ~captured;
  } else if (!__any_sync(kAll, ((around & empty) | (around & own_prev & near_empty)) != 0)) {  // IsSelfCapture, board.cc:901-915
    const uint32_t group = flood(bit, own, lane);
    if (!__any_sync(kAll, (nbrs(group, lane) & empty) != 0)) return false;
  }
  uint64_t h = hash ^ zobrist(point, black ? 0 : 1);
  if (any_capture) h ^= hash_of(captured, black ? 1 : 0, lane);
  bool hit = false;                                   // kRepeatedPosition, board.cc:636-640
  for (int i = lane; i < n_hist_s; i += 32) hit |= hist_s[i] == h;
  for (int i = n_hist_s + lane; i < n_hist; i += 32) hit |= hist_g[i] == h;
  for (int i = lane; i < n_path; i += 32) hit |= path_s[i] == h;
  if (__any_sync(kAll, hit)) return false;
  b.bk = black ? own : opp;
  b.wh = black ? opp : own;
  hash = h;
  return true;
}

__global__ void __launch_bounds__(kReaderWarps * 32) ladder_kernel(const LadderTask* __restrict__ tasks, const int* __restrict__ n_tasks,
                                                                   int* __restrict__ next_task, const uint32_t* __restrict__ rows,
                                                                   const uint64_t* __restrict__ hist, const int32_t* __restrict__ n_hist,
                                                                   int max_moves, DeepFrames* __restrict__ deep,
                                                                   int8_t* __restrict__ laddered, int32_t* __restrict__ status,
                                                                   long long* __restrict__ stats) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int warp = blockIdx.x * kReaderWarps + wib;
  unsigned char* base = smem_raw + wib * kWarpSmemBytes;
  uint64_t* hist_s = reinterpret_cast<uint64_t*>(base);
  uint64_t* path_s = hist_s + kHistSmem;
  uint32_t* frames_s = reinterpret_cast<uint32_t*>(path_s + kMaxDepth);
  DeepFrames& D = deep[warp];
  auto frame = [&](int d) -> uint32_t* { return d < kSmemFrames ? frames_s + d * kFrameWords : D.w[d - kSmemFrames]; };
  const int total = *n_tasks;
  while (true) {
    int t = 0;
    if (lane == 0) t = atomicAdd(next_task, 1);
    t = __shfl_sync(kAll, t, 0);
    if (t >= total) break;
    const LadderTask task = tasks[t];
    const uint32_t* r = rows + static_cast<size_t>(task.pos) * 96;
    const Board root_board{r[lane], r[32 + lane]};
    const uint32_t fb = r[64 + lane];
    const uint64_t* my_hist = hist + static_cast<size_t>(task.pos) * (max_moves + 1);
    const int nh = n_hist[task.pos];
    const int nh_s = min(nh, kHistSmem);
    __syncwarp();
    for (int i = lane; i < nh_s; i += 32) hist_s[i] = my_hist[i];
    __syncwarp();
    const uint32_t rootbit = point_bit(task.root, lane);
    const int g_color = __any_sync(kAll, (rootbit & root_board.bk) != 0) ? P3_BLACK : P3_WHITE;
    const uint32_t root_group = flood(rootbit, g_color == P3_BLACK ? root_board.bk : root_board.wh, lane);

    // Solve(board_copy, gid, g_color, OppositeColor(g_color), root, liberty, 0): the defender extends first (board.cc:863-866)
    Board b = root_board;
    uint32_t grp = root_group;
    uint64_t hash = my_hist[nh - 1];
    int top = -1;            // index of the frame whose children are being tried; the call being entered has call_depth top + 1
    int move = task.liberty;
    int mover = g_color;
    bool value = false;
    bool overflow = false;
    long long nodes = 0;
    const long long c0 = stats ? clock64() : 0;
    while (true) {
      ++nodes;
      // ---- enter Solve(move by `mover`) from the position in (b, hash)
      bool returned = true;
      const int depth = top + 1;
      if (depth > 300) {
        value = false;
      } else if (!play_checked(b, hash, move, mover, fb, hist_s, nh_s, my_hist, nh, path_s, depth, lane)) {
        value = mover == g_color;  // board.cc:782-786: an illegal move loses for its mover
      } else {
        const int to_move = -mover;
        const uint32_t own = g_color == P3_BLACK ? b.bk : b.wh;
        const uint32_t opp = g_color == P3_BLACK ? b.wh : b.bk;
        const uint32_t empty = ~(b.bk | b.wh) & row_mask(lane);
        if (mover == g_color) {  // the group of group_root after the defender's stone (board.cc:788-792)
          const uint32_t mbit = point_bit(move, lane);
          if (__any_sync(kAll, (nbrs(mbit, lane) & grp) != 0)) {        // the stone joins the group ...
            grp |= mbit;
            if (__any_sync(kAll, (nbrs(mbit, lane) & own & ~grp) != 0)) grp = flood(grp, own, lane);  // ... and brings others along
          }
        }
        uint32_t libs = nbrs(grp, lane) & empty;
        const int n_libs = warp_count(libs);
        if (to_move != g_color) {          // attacker to move (board.cc:800-812)
          if (n_libs > 2) value = false;
          else if (n_libs <= 1) value = true;
          else returned = false;
        } else {                           // defender to move (board.cc:813-839)
          if (n_libs > 1) value = false;
          else if (n_libs == 0) value = true;  // unreachable after a legal attacker move (the reference CHECK-fails)
          else returned = false;
        }
        if (!returned) {
          int nc = 0;
          uint32_t cand_reg = 0;           // lanes 19..31: two packed candidate points each
          int first = -1;
          while (true) {                   // the group's liberties: two (attacker) or one (defender)
            const int p = first_point(libs);
            if (p < 0) break;
            libs &= ~point_bit(p, lane);
            if (nc == 0) first = p;
            if (lane == 19 + (nc >> 1)) cand_reg |= static_cast<uint32_t>(p) << ((nc & 1) * 16);
            ++nc;
          }
          if (to_move == g_color) {        // FindSurroundingStonesInAtari + FindLiberty (board.cc:744-770, 827-835)
            // a stone with two empty neighbours of its own is not in atari: skip the flood for it
            uint32_t eu = __shfl_up_sync(kAll, empty, 1), ed = __shfl_down_sync(kAll, empty, 1);
            if (lane == 0) eu = 0;
            if (lane == 31) ed = 0;
            const uint32_t el = empty << 1, er = empty >> 1;
            const uint32_t two = (el & er) | (el & eu) | (el & ed) | (er & eu) | (er & ed) | (eu & ed);
            uint32_t around = nbrs(grp, lane) & opp & ~two;
            uint32_t tried = point_bit(first, lane);
            while (true) {
              const int s = first_point(around);
              if (s < 0) break;
              const uint32_t g2 = flood(point_bit(s, lane), opp, lane);
              around &= ~g2;
              const uint32_t l2 = nbrs(g2, lane) & empty;
              if (warp_count(l2) != 1) continue;
              if (__any_sync(kAll, (l2 & tried) != 0)) continue;  // same point, same position: same answer
              tried |= l2;
              if (nc >= kMaxCandPacked) {
                overflow = true;
                break;
              }
              const int p = first_point(l2);
              if (lane == 19 + (nc >> 1)) cand_reg |= static_cast<uint32_t>(p) << ((nc & 1) * 16);
              ++nc;
            }
          }
          uint32_t* f = frame(depth);
          f[lane] = lane < P3_BOARD_LEN ? b.bk : cand_reg;
          f[32 + lane] = lane == 31 ? static_cast<uint32_t>(nc) : b.wh;   // next = 0
          f[64 + lane] = grp;
          if (lane == 0) path_s[depth] = hash;
          __syncwarp();
          top = depth;
          mover = to_move;
          move = first;
          continue;  // enter the first child from this position
        }
      }
      // ---- a call returned `value`: unwind
      bool done = false;
      while (true) {
        if (top < 0) {
          done = true;
          break;
        }
        uint32_t* f = frame(top);
        const uint32_t meta = f[32 + 31];
        const int n_cand = meta & 0xFF, next = static_cast<int>(meta >> 8) + 1;
        const bool is_and = top & 1;                     // frame d was pushed after the move at call depth d; defender moves at even depths
        const bool decided = is_and ? !value : value;    // AND stops at the first false, OR at the first true
        if (decided || next >= n_cand) {                  // exhausted: AND -> true (all true), OR -> false (all false) == value
          --top;
          continue;
        }
        const uint32_t packed = f[19 + (next >> 1)];
        move = (packed >> ((next & 1) * 16)) & 0xFFFF;
        __syncwarp();
        if (lane == 31) f[32 + 31] = static_cast<uint32_t>(n_cand) | (static_cast<uint32_t>(next) << 8);
        b.bk = lane < P3_BOARD_LEN ? f[lane] : 0u;
        b.wh = lane < P3_BOARD_LEN ? f[32 + lane] : 0u;
        grp = f[64 + lane];
        hash = path_s[top];
        mover = is_and ? g_color : -g_color;
        __syncwarp();
        break;
      }
      if (done) break;
    }
    if (stats && lane == 0) {
      stats[2 * t] = nodes;
      stats[2 * t + 1] = clock64() - c0;
    }
    if (overflow && lane == 0) atomicOr(&status[task.pos], 2);
    if (value && lane < P3_BOARD_LEN) {
      int8_t* l = laddered + static_cast<size_t>(task.pos) * P3_NUM_BOARD_LOCS + lane * P3_BOARD_LEN;
      for (int c = 0; c < P3_BOARD_LEN; ++c)
        if ((root_group >> c) & 1) l[c] = static_cast<int8_t>(g_color);
    }
  }
}

// ---- kernel 3: exact legal mask (Game::IsValidMove over all 362 encodings, cc/game/game.cc:45-51) ----------------------
constexpr int kLegalSplit = 8;  // warps per position

__global__ void __launch_bounds__(256) legal_exact_kernel(const uint32_t* __restrict__ rows, const uint64_t* __restrict__ hist,
                                                          const int32_t* __restrict__ n_hist, int max_moves,
                                                          const int8_t* __restrict__ colors, int n, uint8_t* __restrict__ out) {
  const int lane = threadIdx.x & 31, part = threadIdx.x >> 5;
  const int pos = blockIdx.x;
  const uint32_t* r = rows + static_cast<size_t>(pos) * 96;
  const Board root{r[lane], r[32 + lane]};
  const uint32_t fb = r[64 + lane];
  const uint64_t* my_hist = hist + static_cast<size_t>(pos) * (max_moves + 1);
  const int nh = n_hist[pos];
  const SeenSet seen{my_hist, nh, nullptr, 0};
  const uint64_t root_hash = my_hist[nh - 1];
  const int color = colors[pos];
  uint8_t* o = out + static_cast<size_t>(pos) * P3_MAX_MOVES;
  for (int p = part; p < P3_NUM_BOARD_LOCS; p += kLegalSplit) {
    Board b = root;
    uint64_t h = root_hash;
    const bool ok = play(b, h, p, color, true, fb, seen, lane);
    if (lane == 0) o[p] = ok ? 1 : 0;
  }
  if (threadIdx.x == 0) o[P3_NUM_BOARD_LOCS] = 1;  // pass (board.cc:516-518)
}

}  // namespace

// Host entry: all buffers are device pointers except where noted; scratch is allocated per call.
int ladder_run(const int16_t* d_moves, const int32_t* d_num_moves, int max_moves, const int8_t* d_forbidden, const int8_t* d_colors,
               int n, int8_t* d_boards, int8_t* d_laddered, uint8_t* d_legal, int32_t* d_status, cudaStream_t stream) {
  if (n <= 0) return P3_OK;
  uint32_t* rows = nullptr;
  uint64_t* hist = nullptr;
  int32_t* n_hist = nullptr;
  LadderTask* tasks = nullptr;
  int* counters = nullptr;
  DeepFrames* scratch = nullptr;
  int sms = 148;
  int dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int warps_per_block = kReaderWarps;
  const int blocks = sms * 2;                      // 8 resident warps per SM, each with its own frame stack in shared memory
  const int n_warps = blocks * warps_per_block;
  const size_t reader_smem = kReaderWarps * kWarpSmemBytes;
  int rc = P3_OK;
  auto cleanup = [&]() {
    cudaFree(rows), cudaFree(hist), cudaFree(n_hist), cudaFree(tasks), cudaFree(counters), cudaFree(scratch);
  };
#define P3_TRY(call)                                                                                     \
  do {                                                                                                   \
    cudaError_t _e = (call);                                                                             \
    if (_e != cudaSuccess) {                                                                             \
      cleanup();                                                                                         \
      return fail(P3_ERR_CUDA, std::string(#call) + " -> " + cudaGetErrorString(_e) + " (ladder.cu)");   \
    }                                                                                                    \
  } while (0)
  P3_TRY(cudaMalloc(&rows, static_cast<size_t>(n) * 96 * sizeof(uint32_t)));
  P3_TRY(cudaMalloc(&hist, static_cast<size_t>(n) * (max_moves + 1) * sizeof(uint64_t)));
  P3_TRY(cudaMalloc(&n_hist, static_cast<size_t>(n) * sizeof(int32_t)));
  P3_TRY(cudaMalloc(&tasks, static_cast<size_t>(n) * P3_NUM_BOARD_LOCS / 2 * sizeof(LadderTask)));
  P3_TRY(cudaMalloc(&counters, 2 * sizeof(int)));
  P3_TRY(cudaMemsetAsync(counters, 0, 2 * sizeof(int), stream));
  const bool want_ladder = d_laddered != nullptr;
  const bool trace = std::getenv("P3_LADDER_TRACE") != nullptr;
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
  long long* d_stats = nullptr;
  if (trace) {
    cudaMalloc(&d_stats, static_cast<size_t>(n) * P3_NUM_BOARD_LOCS * sizeof(long long));
    cudaMemset(d_stats, 0, static_cast<size_t>(n) * P3_NUM_BOARD_LOCS * sizeof(long long));
    for (auto& e : ev) cudaEventCreate(&e);
    if (want_ladder) cudaMalloc(&scratch, static_cast<size_t>(n_warps) * sizeof(DeepFrames));
    cudaEventRecord(ev[0], stream);
  }
  replay_kernel<<<(n + 3) / 4, 128, 0, stream>>>(d_moves, d_num_moves, max_moves, d_forbidden, n, rows, hist, n_hist, d_boards,
                                                 d_laddered, d_status, want_ladder ? tasks : nullptr, counters);
  P3_TRY(cudaGetLastError());
  if (trace) cudaEventRecord(ev[1], stream);
  if (want_ladder) {
    if (!scratch) P3_TRY(cudaMalloc(&scratch, static_cast<size_t>(n_warps) * sizeof(DeepFrames)));
    P3_TRY(cudaFuncSetAttribute(ladder_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(reader_smem)));
    ladder_kernel<<<blocks, warps_per_block * 32, reader_smem, stream>>>(tasks, counters, counters + 1, rows, hist, n_hist, max_moves, scratch,
                                                               d_laddered, d_status, d_stats);
    P3_TRY(cudaGetLastError());
  }
  if (trace) cudaEventRecord(ev[2], stream);
  if (d_legal && d_colors) {
    legal_exact_kernel<<<n, kLegalSplit * 32, 0, stream>>>(rows, hist, n_hist, max_moves, d_colors, n, d_legal);
    P3_TRY(cudaGetLastError());
  }
  if (trace) cudaEventRecord(ev[3], stream);
  P3_TRY(cudaStreamSynchronize(stream));
  if (trace) {
    float a = 0, b = 0, c = 0;
    int h_counters[2] = {0, 0};
    cudaMemcpy(h_counters, counters, sizeof(h_counters), cudaMemcpyDeviceToHost);
    cudaEventElapsedTime(&a, ev[0], ev[1]), cudaEventElapsedTime(&b, ev[1], ev[2]), cudaEventElapsedTime(&c, ev[2], ev[3]);
    std::fprintf(stderr, "[p3 ladder] n %d  replay+tasks %.3f ms  reader %.3f ms (%d searches)  exact legal %.3f ms\n", n, a, b,
                 h_counters[0], c);
    for (auto& e : ev) cudaEventDestroy(e);
    if (want_ladder && h_counters[0] > 0) {
      std::vector<long long> st(2 * static_cast<size_t>(h_counters[0]));
      cudaMemcpy(st.data(), d_stats, st.size() * sizeof(long long), cudaMemcpyDeviceToHost);
      long long tot = 0, mx = 0, cyc = 0, mxc = 0;
      for (int i = 0; i < h_counters[0]; ++i) {
        tot += st[2 * i], cyc += st[2 * i + 1];
        if (st[2 * i] > mx) mx = st[2 * i];
        if (st[2 * i + 1] > mxc) mxc = st[2 * i + 1];
      }
      std::fprintf(stderr, "[p3 ladder] nodes total %lld  max per search %lld  cycles per node %.0f  longest search %lld cycles\n", tot, mx,
                   tot ? static_cast<double>(cyc) / tot : 0.0, mxc);
    }
    cudaFree(d_stats);
  }
#undef P3_TRY
  cleanup();
  return rc;
}

}  // namespace p3
