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
#include <algorithm>
#include <cstdlib>
#include <mutex>
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
  bool any_capture = false;
  const uint32_t near_empty = nbrs(empty, lane);                            // points with an empty neighbour
  const uint32_t seeds = nbrs(bit, lane) & opp & ~near_empty;               // adjacent opposing stones without a liberty of their own
  if (__any_sync(kAll, seeds != 0)) {
    const uint32_t touched = flood(seeds, opp, lane);                       // their groups
    const uint32_t alive = flood(near_empty & touched, touched, lane);      // ... that still reach an empty point
    captured = touched & ~alive;
    any_capture = __any_sync(kAll, captured != 0);
  }
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

// Pass-alive regions of one colour (GroupTracker::BensonSolver, cc/game/board.cc:246-462) on the board (bk, wh): the stones of
// the surviving groups and every point (empty or opposing stone) of the surviving small regions.
//   region  = connected set of non-`color` points (the reference's visitor walks empties and opposing stones, :318-340);
//   small   = every EMPTY point of it touches a `color` stone (:325-343);
//   vital to a group = every empty point of the region is a liberty of that group (:376-418);
//   a group with fewer than two vital regions goes, together with every small region it touches (:420-462), until nothing
//   changes.  Removal order does not matter (greatest fixed point), so groups are dropped as they are found.
__device__ __forceinline__ uint32_t benson_color(uint32_t bk, uint32_t wh, bool black, int lane) {
  const uint32_t mine = black ? bk : wh;
  const uint32_t empty = ~(bk | wh) & row_mask(lane);
  const uint32_t other = ~mine & row_mask(lane);                       // empty or opposing
  const uint32_t lonely = empty & ~nbrs(mine, lane);                   // empty points that touch no stone of `color`
  uint32_t regions = other & ~flood(lonely, other, lane);              // union of the small regions
  uint32_t groups = mine;                                              // union of the surviving groups
  while (true) {
    bool changed = false;
    uint32_t todo = groups;
    while (true) {
      const int s = first_point(todo);
      if (s < 0) break;
      const uint32_t g = flood(point_bit(s, lane), mine, lane);
      todo &= ~g;
      const uint32_t reach = nbrs(g, lane);
      uint32_t touching = regions & reach;                             // seeds of the regions next to g
      int vital = 0;
      while (vital < 2) {
        const int t = first_point(touching);
        if (t < 0) break;
        const uint32_t r = flood(point_bit(t, lane), regions, lane);
        touching &= ~r;
        if (!__any_sync(kAll, (r & empty & ~reach) != 0)) ++vital;
      }
      if (vital < 2) {
        groups &= ~g;
        const uint32_t gone = flood(regions & reach, regions, lane);
        if (__any_sync(kAll, gone != 0)) regions &= ~gone;
        changed = true;
      }
    }
    if (!changed) break;
  }
  return groups | regions;
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
  // Board::Pass (board.cc:582-593): from the third pass of the game on (and unless the pass ends the game) the pass-alive regions
  // are recomputed AT THAT MOMENT and stay as they are until the next such pass - keep the board of the last one
  int passes = 0, consecutive = 0;
  bool have_snapshot = false;
  Board snap{0, 0};
  int chunk = 0;
  for (int m = 0; m < nm; ++m) {
    if ((m & 31) == 0) chunk = m + lane < nm ? mv[m + lane] : -1;   // 32 moves per global load instead of one dependent load per move
    const int code = __shfl_sync(kAll, chunk, m & 31);
    const int point = code & (kWhiteBit - 1);
    if (code < 0) continue;                                // padding
    if (point >= P3_PASS_ENCODING) {                       // Board::Pass leaves seen_states_ alone
      ++passes;
      ++consecutive;
      if (consecutive != 2 && passes >= 3) {               // kNumPassesBeforeBensons, cc/constants/constants.h:75
        snap = b;
        have_snapshot = true;
      }
      continue;
    }
    consecutive = 0;
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
  if (have_snapshot)  // GroupTracker::CalculatePassAliveRegions (board.cc:223-233); PlayMoveDry refuses both colours there (:607)
    fb |= benson_color(snap.bk, snap.wh, true, lane) | benson_color(snap.bk, snap.wh, false, lane);
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
  // a group holding a stone with two empty neighbours of its own is not in atari: drop those groups wholesale (one multi-seed
  // flood per colour) and look at the few that remain one by one
  uint32_t eu = __shfl_up_sync(kAll, empty, 1), ed = __shfl_down_sync(kAll, empty, 1);
  if (lane == 0) eu = 0;
  if (lane == 31) ed = 0;
  const uint32_t el = empty << 1, er = empty >> 1;
  const uint32_t two = (el & er) | (el & eu) | (el & ed) | (er & eu) | (er & ed) | (eu & ed);
  uint32_t todo = (b.bk & ~flood(two & b.bk, b.bk, lane)) | (b.wh & ~flood(two & b.wh, b.wh, lane));
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
constexpr int kSmemFrames = 16;
constexpr int kHistSmem = 608;
constexpr int kReaderWarps = 4;
constexpr int kFrameWords = 96;
constexpr int kMaxCandPacked = 26;
constexpr size_t kWarpSmemBytes = kHistSmem * 8 + (kMaxDepth + 1) * 8 + kSmemFrames * kFrameWords * 4;
static_assert(kWarpSmemBytes % 8 == 0, "per-warp shared memory slices stay 8-byte aligned");
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

// A long search is split among warps.  Work items are sub-searches: "the value of Solve() entered with this move from this
// position" (the position - rows, hunted group, hashes of the path to it - travels in a state pool; replaying the path instead
// cost more than the searching).  A warp runs an item as a plain depth-first search; when the item has cost kSplitNodes
// nodes it stops, turns its live stack into nodes of an AND/OR tree (one per frame; the untried sibling moves of every
// frame become new items for whichever warp picks them up) and goes back to the queue.  A finished
// item reports its value to its parent node: a deciding value (false under AND, true under OR) settles the parent at once,
// otherwise the parent's pending count drops and the last child settles it with the neutral value; settled nodes report
// upwards, a settled root writes the task's answer.  Items whose ancestors are already settled are dropped unrun.  Every
// sub-search is a pure function of (position, path), so the answer is the sequential one; only the order differs.
constexpr int kSplitNodesDefault = 6;   // measured (1024 random-playout positions): unsplit 10.6 ms; items replaying their path: 0.98 ms at 16;
                                       // items carrying their position: 0.98 ms at 32, 0.53 at 16, 0.40 at 8, 0.37 at 6, 0.41 at 4, 0.61 at 2
constexpr unsigned kDecided = 0x80000000u, kValTrue = 0x40000000u, kIsAnd = 0x20000000u, kPendMask = 0xFFFFu;

struct TreeNode {
  int parent;        // -1: the root of a task's search
  unsigned state;    // kDecided | kValTrue | kIsAnd | pending children
};

struct Item {
  int task, node;
  int base_move;     // call depth of the move to enter | move << 16
  int state_at;      // offset (in 8-byte words) of the position to enter it from in the state pool, -1 = the task's root position:
                     //   [base] hashes of the path's positions, then 48 words = the frame's 96 row words (black, white, hunted group)
};
constexpr int kStateWords = 48;

struct Queue {
  int n_tasks;       // written by replay_kernel (atomic counter)
  int tail;          // items reserved
  int head;          // items claimed
  int outstanding;   // items queued or running
  int n_nodes;       // tree nodes allocated
  int dropped;       // items skipped because an ancestor was settled
  int splits;        // successful splits
  int full;          // splits refused for lack of room (the item then just keeps searching)
  int pool_used;     // 8-byte words of the state pool handed out
  int pad;
  long long nodes;   // Solve() activations over all items
};

__global__ void seed_items_kernel(const LadderTask* __restrict__ tasks, Queue* q, TreeNode* nodes, Item* items, int* ready) {
  const int n = q->n_tasks;   // <= the capacities by construction (ladder_run sizes them from the batch)
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) {
    nodes[t] = TreeNode{-1, 0u};
    items[t] = Item{t, t, tasks[t].liberty << 16, -1};
    __threadfence();
    ready[t] = 1;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    q->tail = n;
    q->outstanding = n;
    q->n_nodes = n;
    q->pool_used = 0;
  }
}

// lane 0 only: child of `parent` finished with `v`; returns true when that settled the task's root (value in *root_value)
__device__ __forceinline__ bool report_up(TreeNode* nodes, int parent, bool v, bool* root_value) {
  while (true) {
    if (parent < 0) {
      *root_value = v;
      return true;
    }
    unsigned* st = &nodes[parent].state;
    bool settled = false;
    while (true) {
      const unsigned s = *reinterpret_cast<volatile unsigned*>(st);
      if (s & kDecided) return false;            // somebody else settled it: this report is moot
      const bool is_and = s & kIsAnd;
      const bool deciding = is_and ? !v : v;
      unsigned ns;
      if (deciding || (s & kPendMask) == 1) {
        ns = (s & ~kPendMask) | kDecided | (v ? kValTrue : 0u);   // the last non-deciding child carries the neutral value
        settled = true;
      } else {
        ns = s - 1;
        settled = false;
      }
      if (atomicCAS(st, s, ns) == s) break;
    }
    if (!settled) return false;
    parent = nodes[parent].parent;               // the parent's value is v: carry it upwards
  }
}

__global__ void __launch_bounds__(kReaderWarps * 32) ladder_kernel(const LadderTask* __restrict__ tasks, Queue* q, TreeNode* nodes,
                                                                   int node_cap, Item* items, uint64_t* pool, int pool_cap,
                                                                   int* ready, int item_cap,
                                                                   const uint32_t* __restrict__ rows, const uint64_t* __restrict__ hist,
                                                                   const int32_t* __restrict__ n_hist, int max_moves,
                                                                   DeepFrames* __restrict__ deep, int8_t* __restrict__ laddered,
                                                                   int32_t* __restrict__ status, int split_nodes) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int warp = blockIdx.x * kReaderWarps + wib;
  unsigned char* base_ptr = smem_raw + wib * kWarpSmemBytes;
  uint64_t* hist_s = reinterpret_cast<uint64_t*>(base_ptr);
  uint64_t* path_s = hist_s + kHistSmem;
  uint32_t* frames_s = reinterpret_cast<uint32_t*>(path_s + kMaxDepth + 1);
  DeepFrames& D = deep[warp];
  auto frame = [&](int d) -> uint32_t* { return d < kSmemFrames ? frames_s + d * kFrameWords : D.w[d - kSmemFrames]; };
  volatile int* v_ready = ready;
  volatile int* v_out = &q->outstanding;
  long long my_nodes = 0;
  while (true) {
    // ---- claim the next item; wait until it is published or nothing is left anywhere
    int idx = 0;
    if (lane == 0) idx = atomicAdd(&q->head, 1);
    idx = __shfl_sync(kAll, idx, 0);
    bool have = false;
    for (long long polls = 0;; ++polls) {
      if (polls > (1ll << 25)) {  // ~10 s of polling: something is wrong; leave instead of hanging the device
        if (lane == 0) {
          atomicOr(&q->full, 1 << 30);
          atomicOr(&status[0], 8);  // reported on the batch's first position: the whole batch's reader output is incomplete
        }
        break;
      }
      int r = 0, out = 1;
      if (lane == 0) {
        r = idx < item_cap ? v_ready[idx] : 0;
        out = *v_out;
      }
      r = __shfl_sync(kAll, r, 0);
      out = __shfl_sync(kAll, out, 0);
      if (r) {
        have = true;
        break;
      }
      if (out == 0) break;
      __nanosleep(200);
    }
    if (!have) break;
    __threadfence();
    const Item it = items[idx];
    const int task_id = it.task, node_id = it.node, base = it.base_move & 0xFFFF, first_move = it.base_move >> 16;
    if (task_id < 0) {  // a reserved slot that was given back
      if (lane == 0) atomicSub(&q->outstanding, 1);
      continue;
    }
    // dropped if an ancestor is already settled
    int cancelled = 0;
    if (lane == 0) {
      int nd = nodes[node_id].parent;
      while (nd >= 0) {
        if (*reinterpret_cast<volatile unsigned*>(&nodes[nd].state) & kDecided) {
          cancelled = 1;
          break;
        }
        nd = nodes[nd].parent;
      }
    }
    cancelled = __shfl_sync(kAll, cancelled, 0);
    if (cancelled) {
      if (lane == 0) {
        atomicAdd(&q->dropped, 1);
        atomicSub(&q->outstanding, 1);
      }
      continue;
    }
    const LadderTask task = tasks[task_id];
    const uint32_t* r = rows + static_cast<size_t>(task.pos) * 96;
    const Board root_board{r[lane], r[32 + lane]};
    const uint32_t fb = r[64 + lane];
    const uint64_t* my_hist = hist + static_cast<size_t>(task.pos) * (max_moves + 1);
    const int nh = n_hist[task.pos];
    const int nh_s = min(nh, kHistSmem);
    __syncwarp();
    for (int i = lane; i < nh_s; i += 32) hist_s[i] = my_hist[i];
    if (it.state_at >= 0)
      for (int i = lane; i < base; i += 32) path_s[i] = pool[it.state_at + i];
    __syncwarp();
    const uint32_t rootbit = point_bit(task.root, lane);
    const int g_color = __any_sync(kAll, (rootbit & root_board.bk) != 0) ? P3_BLACK : P3_WHITE;
    const uint32_t root_group = flood(rootbit, g_color == P3_BLACK ? root_board.bk : root_board.wh, lane);

    Board b = root_board;
    uint32_t grp = root_group;
    uint64_t hash = my_hist[nh - 1];
    bool bad = false;
    // the defender's group after one of his stones (board.cc:788-792): it only grows
    auto grow = [&](int mv) {
      const uint32_t own = g_color == P3_BLACK ? b.bk : b.wh;
      const uint32_t mbit = point_bit(mv, lane);
      if (__any_sync(kAll, (nbrs(mbit, lane) & grp) != 0)) {        // the stone joins the group ...
        grp |= mbit;
        if (__any_sync(kAll, (nbrs(mbit, lane) & own & ~grp) != 0)) grp = flood(grp, own, lane);  // ... and brings others along
      }
    };
    // ---- the position this item starts from: the task's root, or the frame state its parent item left in the pool
    if (it.state_at >= 0) {
      const uint32_t* st = reinterpret_cast<const uint32_t*>(pool + it.state_at + base);
      b.bk = lane < P3_BOARD_LEN ? st[lane] : 0u;
      b.wh = lane < P3_BOARD_LEN ? st[32 + lane] : 0u;
      grp = st[64 + lane];
      hash = path_s[base - 1];
    }
    int top = base - 1;      // index of the frame whose children are being tried; the call being entered has call_depth top + 1
    int move = first_move;
    int mover = (base & 1) ? -g_color : g_color;   // Solve(..., liberty, 0) is the defender's move (board.cc:863-866)
    bool value = false;
    bool overflow = false, split_done = false, may_split = true;
    int nodes_here = 0;
    while (true) {
      // ---- over budget: hand the untried parts of the stack to other warps
      if (nodes_here >= split_nodes && top >= base && may_split) {
        int n_items = 0;
        for (int j = base; j <= top; ++j) {
          const uint32_t meta = frame(j)[32 + 31];
          const int n_cand = meta & 0xFF, next = static_cast<int>(meta >> 8);
          n_items += j < top ? n_cand - next - 1 : n_cand - next;
        }
        const int n_inner = top - base;
        int n_pool = 0;
        for (int j = base; j <= top; ++j) {
          const uint32_t meta = frame(j)[32 + 31];
          const int n_cand = meta & 0xFF, next = static_cast<int>(meta >> 8);
          if ((j < top ? n_cand - next - 1 : n_cand - next) > 0) n_pool += j + 1 + kStateWords;   // one state per frame, shared by its siblings
        }
        int node0 = 0, item0 = 0, pool0 = 0, ok = 0;
        if (lane == 0) {
          node0 = atomicAdd(&q->n_nodes, n_inner + n_items);
          item0 = atomicAdd(&q->tail, n_items);
          pool0 = atomicAdd(&q->pool_used, n_pool);
          ok = node0 + n_inner + n_items <= node_cap && item0 + n_items <= item_cap && pool0 + n_pool <= pool_cap;
          atomicAdd(&q->outstanding, n_items);   // before anything is published
          if (!ok) atomicAdd(&q->full, 1);
        }
        node0 = __shfl_sync(kAll, node0, 0);
        item0 = __shfl_sync(kAll, item0, 0);
        pool0 = __shfl_sync(kAll, pool0, 0);
        ok = __shfl_sync(kAll, ok, 0);
        if (!ok) {
          // no room: give the reserved queue slots back as no-ops so that nobody waits on them, and keep searching alone
          if (lane == 0) {
            int given = 0;
            for (int k = 0; k < n_items; ++k)
              if (item0 + k < item_cap) {
                items[item0 + k].task = -1;
                ++given;
              }
            __threadfence();
            for (int k = 0; k < n_items; ++k)
              if (item0 + k < item_cap) v_ready[item0 + k] = 1;
            atomicSub(&q->outstanding, n_items - given);
          }
          may_split = false;
        } else {
          if (lane == 0) atomicAdd(&q->splits, 1);
          int leaf = node0 + n_inner, slot = item0, at = pool0;
          for (int j = base; j <= top; ++j) {
            const uint32_t* f = frame(j);
            const uint32_t meta = f[32 + 31];
            const int n_cand = meta & 0xFF, next = static_cast<int>(meta >> 8);
            const int first = j < top ? next + 1 : next;                      // frame top's next move has not been entered yet
            const int my_node = j == base ? node_id : node0 + (j - base - 1); // the item's own node turns from a leaf into an inner node
            const int pending = (n_cand - first) + (j < top ? 1 : 0);         // untried siblings + the child in progress (frame j + 1)
            if (lane == 0) {
              if (j > base) nodes[my_node].parent = j - 1 == base ? node_id : my_node - 1;
              atomicExch(&nodes[my_node].state, ((j & 1) ? kIsAnd : 0u) | static_cast<unsigned>(pending));
            }
            if (first < n_cand) {  // the position of frame j (after the move at call depth j) + the hashes of the path to it
              for (int i = lane; i <= j; i += 32) pool[at + i] = path_s[i];
              uint32_t* st = reinterpret_cast<uint32_t*>(pool + at + j + 1);
              st[lane] = f[lane];
              st[32 + lane] = f[32 + lane];
              st[64 + lane] = f[64 + lane];
            }
            for (int c = first; c < n_cand; ++c) {
              const uint32_t packed = f[19 + (c >> 1)];
              const int mv = (packed >> ((c & 1) * 16)) & 0xFFFF;
              if (lane == 0) {
                items[slot] = Item{task_id, leaf, (j + 1) | (mv << 16), at};
                nodes[leaf] = TreeNode{my_node, 0u};
              }
              ++leaf;
              ++slot;
            }
            if (first < n_cand) at += j + 1 + kStateWords;
          }
          __threadfence();
          __syncwarp();
          if (lane == 0)
            for (int k = 0; k < n_items; ++k) v_ready[item0 + k] = 1;
          split_done = true;
          break;
        }
      }
      ++nodes_here;
      // ---- enter Solve(move by `mover`) from the position in (b, hash)
      bool returned = true;
      const int depth = top + 1;
      if (depth > 300) {
        value = false;
      } else if (!play_checked(b, hash, move, mover, fb, hist_s, nh_s, my_hist, nh, path_s, depth, lane)) {
        value = mover == g_color;  // board.cc:782-786: an illegal move loses for its mover
      } else {
        const int to_move = -mover;
        const uint32_t opp = g_color == P3_BLACK ? b.wh : b.bk;
        const uint32_t empty = ~(b.bk | b.wh) & row_mask(lane);
        if (mover == g_color) grow(move);
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
        if (top < base) {
          done = true;
          break;
        }
        uint32_t* f = frame(top);
        const uint32_t meta = f[32 + 31];
        const int n_cand = meta & 0xFF, next = static_cast<int>(meta >> 8) + 1;
        const bool is_and = top & 1;                     // frame d was pushed after the move at call depth d; the defender moves at even depths
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
    my_nodes += nodes_here;
    if ((overflow || bad) && lane == 0) atomicOr(&status[task.pos], overflow ? 2 : 4);
    if (!split_done) {
      int root_done = 0, root_val = 0;
      if (lane == 0) {
        bool rv = false;
        root_done = report_up(nodes, nodes[node_id].parent, value, &rv) ? 1 : 0;
        root_val = rv ? 1 : 0;
      }
      root_done = __shfl_sync(kAll, root_done, 0);
      root_val = __shfl_sync(kAll, root_val, 0);
      if (root_done && root_val && lane < P3_BOARD_LEN) {
        int8_t* l = laddered + static_cast<size_t>(task.pos) * P3_NUM_BOARD_LOCS + lane * P3_BOARD_LEN;
        for (int c = 0; c < P3_BOARD_LEN; ++c)
          if ((root_group >> c) & 1) l[c] = static_cast<int8_t>(g_color);
      }
    }
    __syncwarp();
    if (lane == 0) {
      __threadfence();
      atomicSub(&q->outstanding, 1);
    }
  }
  if (lane == 0 && my_nodes) atomicAdd(reinterpret_cast<unsigned long long*>(&q->nodes), static_cast<unsigned long long>(my_nodes));
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
struct LadderWorkspace {
  int n = 0, max_moves = 0, blocks = 0, item_cap = 0, node_cap = 0, pool_cap = 0, split_nodes = kSplitNodesDefault, sms = 148;
  int split_saved = -1;  // >= 0 while the reader runs unsplit (ladder_workspace_set_unsplit)
  uint32_t* rows = nullptr;
  uint64_t* hist = nullptr;
  int32_t* n_hist = nullptr;
  LadderTask* tasks = nullptr;
  Queue* queue = nullptr;
  TreeNode* nodes = nullptr;
  Item* items = nullptr;
  uint64_t* pool = nullptr;
  int* ready = nullptr;
  DeepFrames* scratch = nullptr;
};

// The retry path of the reader's watchdog (status bit 3): with splitting off every search runs on the warp that claimed it, no
// warp ever waits for another one's item, and the reader cannot stall - at the round-1 price (one 4 598-node search = 10 ms).
void ladder_workspace_set_unsplit(LadderWorkspace* w, bool on) {
  if (!w) return;
  if (on && w->split_saved < 0) {
    w->split_saved = w->split_nodes;
    w->split_nodes = 0x7fffffff;
  } else if (!on && w->split_saved >= 0) {
    w->split_nodes = w->split_saved;
    w->split_saved = -1;
  }
}

namespace {
__global__ void force_watchdog_kernel(int32_t* status) { atomicOr(&status[0], 8); }
}  // namespace

void ladder_workspace_destroy(LadderWorkspace* w) {
  if (!w) return;
  cudaFree(w->rows), cudaFree(w->hist), cudaFree(w->n_hist), cudaFree(w->tasks), cudaFree(w->queue), cudaFree(w->nodes), cudaFree(w->items),
      cudaFree(w->pool), cudaFree(w->ready), cudaFree(w->scratch);
  delete w;
}

int ladder_workspace_create(int n, int max_moves, LadderWorkspace** out) {
  LadderWorkspace* w = new LadderWorkspace();
  w->n = n, w->max_moves = max_moves;
  int dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&w->sms, cudaDevAttrMultiProcessorCount, dev);
  w->blocks = w->sms * 2;                                  // 8 resident warps per SM (measured: 2, 3, 4 blocks per SM are equal - the split chain's latency bounds the reader)
  if (const char* bl = std::getenv("P3_LADDER_BLOCKS_PER_SM")) w->blocks = w->sms * std::max(1, std::atoi(bl));
  const int max_tasks = n * (P3_NUM_BOARD_LOCS / 2);       // groups in atari per position
  w->item_cap = max_tasks + (1 << 17);
  w->node_cap = max_tasks + (1 << 19);
  w->pool_cap = 1 << 23;                                   // 64 MB of sub-search states
  if (const char* sn = std::getenv("P3_LADDER_SPLIT")) w->split_nodes = std::max(1, std::atoi(sn));
#define P3_TRY(call)                                                                                     \
  do {                                                                                                   \
    cudaError_t _e = (call);                                                                             \
    if (_e != cudaSuccess) {                                                                             \
      ladder_workspace_destroy(w);                                                                       \
      return fail(P3_ERR_CUDA, std::string(#call) + " -> " + cudaGetErrorString(_e) + " (ladder.cu)");   \
    }                                                                                                    \
  } while (0)
  P3_TRY(cudaMalloc(&w->rows, static_cast<size_t>(n) * 96 * sizeof(uint32_t)));
  P3_TRY(cudaMalloc(&w->hist, static_cast<size_t>(n) * (max_moves + 1) * sizeof(uint64_t)));
  P3_TRY(cudaMalloc(&w->n_hist, static_cast<size_t>(n) * sizeof(int32_t)));
  P3_TRY(cudaMalloc(&w->queue, sizeof(Queue)));
  P3_TRY(cudaMalloc(&w->tasks, static_cast<size_t>(max_tasks) * sizeof(LadderTask)));
  P3_TRY(cudaMalloc(&w->nodes, static_cast<size_t>(w->node_cap) * sizeof(TreeNode)));
  P3_TRY(cudaMalloc(&w->items, static_cast<size_t>(w->item_cap) * sizeof(Item)));
  P3_TRY(cudaMalloc(&w->pool, static_cast<size_t>(w->pool_cap) * sizeof(uint64_t)));
  P3_TRY(cudaMalloc(&w->ready, static_cast<size_t>(w->item_cap) * sizeof(int)));
  P3_TRY(cudaMalloc(&w->scratch, static_cast<size_t>(w->blocks) * kReaderWarps * sizeof(DeepFrames)));
  P3_TRY(cudaFuncSetAttribute(ladder_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kReaderWarps * kWarpSmemBytes)));
#undef P3_TRY
  *out = w;
  return P3_OK;
}

// Asynchronous: replay -> (reader) -> (exact legal mask) on `stream`; device pointers; n <= the workspace's capacity.
int ladder_enqueue(LadderWorkspace* w, const int16_t* d_moves, const int32_t* d_num_moves, const int8_t* d_forbidden,
                   const int8_t* d_colors, int n, int8_t* d_boards, int8_t* d_laddered, uint8_t* d_legal, int32_t* d_status,
                   cudaStream_t stream, cudaEvent_t* ev) {
  if (n <= 0) return P3_OK;
  if (n > w->n) return fail(P3_ERR_INVALID_ARG, "ladder_enqueue: batch exceeds the workspace");
  const bool want_ladder = d_laddered != nullptr;
  P3_CUDA(cudaMemsetAsync(w->queue, 0, sizeof(Queue), stream));
  if (want_ladder) P3_CUDA(cudaMemsetAsync(w->ready, 0, static_cast<size_t>(w->item_cap) * sizeof(int), stream));
  if (ev) cudaEventRecord(ev[0], stream);
  replay_kernel<<<(n + 3) / 4, 128, 0, stream>>>(d_moves, d_num_moves, w->max_moves, d_forbidden, n, w->rows, w->hist, w->n_hist, d_boards,
                                                 d_laddered, d_status, want_ladder ? w->tasks : nullptr, &w->queue->n_tasks);
  P3_CUDA(cudaGetLastError());
  if (ev) cudaEventRecord(ev[1], stream);
  if (want_ladder) {
    seed_items_kernel<<<w->sms, 256, 0, stream>>>(w->tasks, w->queue, w->nodes, w->items, w->ready);
    P3_CUDA(cudaGetLastError());
    ladder_kernel<<<w->blocks, kReaderWarps * 32, kReaderWarps * kWarpSmemBytes, stream>>>(
        w->tasks, w->queue, w->nodes, w->node_cap, w->items, w->pool, w->pool_cap, w->ready, w->item_cap, w->rows, w->hist, w->n_hist,
        w->max_moves, w->scratch, d_laddered, d_status, w->split_nodes);
    P3_CUDA(cudaGetLastError());
    // test hook: pretend the watchdog fired on every split run, so that the unsplit retry is what produces the results
    const bool force = std::getenv("P3_LADDER_FORCE_WATCHDOG") != nullptr;
    if (force && w->split_saved < 0) force_watchdog_kernel<<<1, 1, 0, stream>>>(d_status);
  }
  if (ev) cudaEventRecord(ev[2], stream);
  if (d_legal && d_colors) {
    legal_exact_kernel<<<n, kLegalSplit * 32, 0, stream>>>(w->rows, w->hist, w->n_hist, w->max_moves, d_colors, n, d_legal);
    P3_CUDA(cudaGetLastError());
  }
  if (ev) cudaEventRecord(ev[3], stream);
  return P3_OK;
}

int ladder_run(const int16_t* d_moves, const int32_t* d_num_moves, int max_moves, const int8_t* d_forbidden, const int8_t* d_colors,
               int n, int8_t* d_boards, int8_t* d_laddered, uint8_t* d_legal, int32_t* d_status, cudaStream_t stream) {
  if (n <= 0) return P3_OK;
  // the stand-alone entry keeps one workspace per device between calls (its ~170 MB of scratch cost more to allocate than the
  // kernels take to run); calls are serialised on it
  static std::mutex mu;
  static LadderWorkspace* cached[64] = {};
  std::lock_guard<std::mutex> lock(mu);
  int dev = 0;
  cudaGetDevice(&dev);
  LadderWorkspace*& w = cached[dev & 63];
  int rc = P3_OK;
  if (!w || w->n < n || w->max_moves != max_moves) {
    ladder_workspace_destroy(w);
    w = nullptr;
    if ((rc = ladder_workspace_create(n, max_moves, &w))) return rc;
  }
  const bool trace = std::getenv("P3_LADDER_TRACE") != nullptr;
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
  if (trace)
    for (auto& e : ev) cudaEventCreate(&e);
  rc = ladder_enqueue(w, d_moves, d_num_moves, d_forbidden, d_colors, n, d_boards, d_laddered, d_legal, d_status, stream,
                      trace ? ev : nullptr);
  cudaError_t se = cudaStreamSynchronize(stream);
  if (!rc && se != cudaSuccess) rc = fail(P3_ERR_CUDA, std::string("ladder_run: ") + cudaGetErrorString(se));
  if (!rc && d_laddered && d_status) {  // the reader's watchdog fired: run the batch again without splitting (cannot stall)
    int32_t st0 = 0;
    if (cudaMemcpy(&st0, d_status, sizeof(st0), cudaMemcpyDeviceToHost) == cudaSuccess && (st0 & 8)) {
      std::fprintf(stderr, "[p3 ladder] reader watchdog fired; re-running %d records unsplit\n", n);
      ladder_workspace_set_unsplit(w, true);
      rc = ladder_enqueue(w, d_moves, d_num_moves, d_forbidden, d_colors, n, d_boards, d_laddered, d_legal, d_status, stream, nullptr);
      se = cudaStreamSynchronize(stream);
      ladder_workspace_set_unsplit(w, false);
      if (!rc && se != cudaSuccess) rc = fail(P3_ERR_CUDA, std::string("ladder_run (unsplit retry): ") + cudaGetErrorString(se));
    }
  }
  if (trace && !rc) {
    float a = 0, b = 0, c = 0;
    Queue h{};
    cudaMemcpy(&h, w->queue, sizeof(h), cudaMemcpyDeviceToHost);
    cudaEventElapsedTime(&a, ev[0], ev[1]), cudaEventElapsedTime(&b, ev[1], ev[2]), cudaEventElapsedTime(&c, ev[2], ev[3]);
    std::fprintf(stderr,
                 "[p3 ladder] n %d  replay+tasks %.3f ms  reader %.3f ms (%d searches, %lld nodes, %d items, %d splits, %d dropped, "
                 "%d refused)  exact legal %.3f ms\n",
                 n, a, b, h.n_tasks, h.nodes, h.tail, h.splits, h.dropped, h.full, c);
  }
  if (trace)
    for (auto& e : ev) cudaEventDestroy(e);
  return rc;
}

// ---- GoFeatures from the derived grids (NNInterface::LoadBatch, cc/nn/nn_interface.cc:245-277, in identity orientation) --------
__global__ void __launch_bounds__(128) assemble_features_kernel(const int16_t* __restrict__ moves, const int32_t* __restrict__ num_moves,
                                                                int max_moves, const int8_t* __restrict__ boards,
                                                                const int8_t* __restrict__ libs, const int8_t* __restrict__ laddered, int n,
                                                                p3_go_features* __restrict__ feats) {
  const int b = blockIdx.x;
  const int nm = num_moves[b];
  if (nm < 0) return;  // this slot was loaded as GoFeatures
  p3_go_features& f = feats[b];
  const size_t at = static_cast<size_t>(b) * P3_NUM_BOARD_LOCS;
  for (int p = threadIdx.x; p < P3_NUM_BOARD_LOCS; p += blockDim.x) {
    f.board[p] = boards[at + p];
    f.stones_atari[p] = libs[at * 3 + p];
    f.stones_two_liberties[p] = libs[at * 3 + P3_NUM_BOARD_LOCS + p];
    f.stones_three_liberties[p] = libs[at * 3 + 2 * P3_NUM_BOARD_LOCS + p];
    f.stones_laddered[p] = laddered[at + p];
  }
  if (threadIdx.x < P3_NUM_LAST_MOVES) {  // the last five moves, oldest first; kNoopLoc pad, kPassLoc for passes (nn_interface.cc:249-257)
    const int k = threadIdx.x;
    const int count = min(nm, max_moves);
    const int off = count - P3_NUM_LAST_MOVES + k;
    p3_loc loc{-1, -1};
    if (off >= 0) {
      const int code = moves[static_cast<size_t>(b) * max_moves + off];
      const int point = code & (kWhiteBit - 1);
      if (code >= 0) loc = point >= P3_PASS_ENCODING ? p3_loc{19, 0} : p3_loc{point / P3_BOARD_LEN, point % P3_BOARD_LEN};
    }
    f.last_moves[k] = loc;
  }
  if (threadIdx.x == 0) f.bsize = P3_BOARD_LEN;
}

int assemble_features_launch(const int16_t* d_moves, const int32_t* d_num_moves, int max_moves, const int8_t* d_boards, const int8_t* d_libs,
                             const int8_t* d_laddered, int n, p3_go_features* d_feats, cudaStream_t stream) {
  if (n <= 0) return P3_OK;
  assemble_features_kernel<<<n, 128, 0, stream>>>(d_moves, d_num_moves, max_moves, d_boards, d_libs, d_laddered, n, d_feats);
  P3_CUDA(cudaGetLastError());
  return P3_OK;
}

}  // namespace p3
