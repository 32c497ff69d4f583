// Board-state -> feature-plane encoding on the GPU (HBM-bound byte work, no tensor cores).
//
//   encode_kernel     nn::LoadGoFeatures after the zero fill  (cc/nn/engine/go_features.cc:10-68,
//                     cc/nn/engine/buf_utils.h:57-87, cc/nn/engine/trt_engine.cc:230-233)
//   liberties_kernel  Board::GetStonesWithLiberties(1|2|3) from the raw position (cc/game/board.cc:670-690)
//   legal_kernel      Board::PlayMoveDry minus history (cc/game/board.cc:595-644, :901-915)
//
// One CTA per position: the 1860-byte game state is staged into shared memory with coalesced
// 32-bit loads, each point's 15 plane bits are formed once, and the NHWC fp32 planes are written
// with fully coalesced stores (consecutive threads -> consecutive floats).
#include "common.cuh"

namespace p3 {
namespace {

constexpr int kFeatBytes = 1860;
constexpr int kFeatWords = kFeatBytes / 4;
static_assert(sizeof(p3_go_features) == kFeatBytes, "p3_go_features must mirror nn::GoFeatures (1860 B)");
static_assert(sizeof(p3_infer_result) == 7568, "p3_infer_result must mirror nn::NNInferResult (7568 B)");

__global__ void __launch_bounds__(256) encode_kernel(const p3_go_features* __restrict__ feats, int n, int version,
                                                     float* __restrict__ planes, float* __restrict__ scalars,
                                                     uint16_t* __restrict__ masks, EncodeExtra ex) {
  __shared__ uint32_t s_raw[kFeatWords];
  __shared__ uint16_t s_mask[P3_NUM_BOARD_LOCS];
  __shared__ float s_scal[P3_NUM_SCALARS_V1];
  const int b = blockIdx.x;
  if (b >= n) return;
  const int np = version == 0 ? P3_NUM_PLANES_V0 : P3_NUM_PLANES_V1;
  const int ns = version == 0 ? P3_NUM_SCALARS_V0 : P3_NUM_SCALARS_V1;

  const uint32_t* src = reinterpret_cast<const uint32_t*>(feats) + static_cast<size_t>(b) * kFeatWords;
  for (int i = threadIdx.x; i < kFeatWords; i += blockDim.x) s_raw[i] = src[i];
  __syncthreads();
  const p3_go_features& f = *reinterpret_cast<const p3_go_features*>(s_raw);
  const int bsize = min(max(f.bsize, 0), P3_BOARD_LEN);  // a torn / garbage record must not index outside the 361-entry grids
  const int8_t color = f.color;
  // ApplySymmetry (cc/game/symmetry.h:42-51): sym_grid[T(i)] = grid[i], i.e. output point p reads source point Tinv(p);
  // the caller hands over identity-orientation features when it gives a symmetry (NNInterface::LoadBatch, nn_interface.cc:245-277)
  const int sym = ex.sym ? ex.sym[b] : 0;

  // plane bits per point: FillPlanePair (buf_utils.h:57-76) reads grid[i*bsize + j] for i, j < bsize
  for (int p = threadIdx.x; p < P3_NUM_BOARD_LOCS; p += blockDim.x) {
    const int i = p / P3_BOARD_LEN, j = p % P3_BOARD_LEN;
    const int sp = sym_transform_inv(sym, p), si = sp / P3_BOARD_LEN, sj = sp % P3_BOARD_LEN;
    uint32_t m = 0;
    if (si < bsize && sj < bsize) {
      const int src_idx = si * bsize + sj;
      auto pair = [&](const int8_t* grid, int ours, int theirs) {
        const int8_t c = grid[src_idx];
        if (c == color) m |= 1u << ours;
        else if (c == static_cast<int8_t>(-color)) m |= 1u << theirs;
      };
      pair(f.board, 0, 1);                    // go_features.cc:12-13
      pair(f.stones_atari, 7, 8);             // :14-15
      pair(f.stones_two_liberties, 9, 10);    // :16-18
      pair(f.stones_three_liberties, 11, 12); // :19-21
      if (version >= 1) pair(f.stones_laddered, 13, 14);  // :22-26
    }
#pragma unroll
    for (int k = 0; k < P3_NUM_LAST_MOVES; ++k) {  // :27-36, skips noop {-1,-1} and pass {19,0}
      const p3_loc lm = f.last_moves[k];
      if (lm.i >= 0 && lm.i < P3_BOARD_LEN && lm.j >= 0 && lm.j < P3_BOARD_LEN && lm.i * P3_BOARD_LEN + lm.j == sp)
        m |= 1u << (k + 2);
    }
    s_mask[p] = static_cast<uint16_t>(m);
    if (masks) masks[static_cast<size_t>(b) * P3_NUM_BOARD_LOCS + p] = static_cast<uint16_t>(m);
    if (ex.masks_padded)  // zero-bordered 23 x 24 grid (the borders are zeroed once by the owner of the buffer)
      ex.masks_padded[static_cast<size_t>(b) * kMaskPadElems + (i + 2) * kMaskPadW + (j + 2)] = static_cast<uint16_t>(m);
  }
  __syncthreads();

  // expand to the reference's NHWC fp32 planes: element e = point * np + channel, coalesced
  float* dst = planes + static_cast<size_t>(b) * P3_NUM_BOARD_LOCS * np;
  const int total = P3_NUM_BOARD_LOCS * np;
  for (int e = threadIdx.x; e < total; e += blockDim.x) {
    const int p = e / np, ch = e - p * np;
    dst[e] = ((s_mask[p] >> ch) & 1u) ? 1.0f : 0.0f;
  }

  if (threadIdx.x < ns) {  // LoadFeatures, go_features.cc:39-60
    const int s = threadIdx.x;
    float v = 0.0f;
    if (s == 0) v = (color == P3_BLACK) ? 1.0f : 0.0f;
    else if (s == 1) v = (color == P3_BLACK) ? 0.0f : 1.0f;
    else if (s < 7) {
      const p3_loc lm = f.last_moves[s - 2];
      v = (lm.i == 19 && lm.j == 0) ? 1.0f : 0.0f;
    } else {  // s == 7, version >= 1: (BLACK ? -1 : 1) * komi / 15, evaluated left to right in fp32
      v = __fdiv_rn(__fmul_rn(color == P3_BLACK ? -1.0f : 1.0f, f.komi), 15.0f);
    }
    scalars[static_cast<size_t>(b) * ns + s] = v;
    s_scal[s] = v;
  }
  if (ex.gs_out) {  // game-state dense of the tower's first layer (python/model.py:1234-1237): [C] bias of this position
    __syncthreads();
    for (int c = threadIdx.x; c < ex.C; c += blockDim.x) {
      float acc = ex.gs_b[c];
      for (int s = 0; s < ns; ++s) acc = fmaf(s_scal[s], ex.gs_w[s * ex.C + c], acc);
      ex.gs_out[static_cast<size_t>(b) * ex.C + c] = acc;
    }
  }
}

// ---- groups & liberties from the raw position -----------------------------------------------------
// Min-label propagation over same-coloured neighbours with pointer jumping, then every empty point
// adds one liberty to each DISTINCT adjacent group.  384 threads, one per point.
__device__ __forceinline__ void group_labels(const int8_t* s_board, int* s_label, int p, bool live) {
  __shared__ int s_changed;
  if (live) s_label[p] = s_board[p] != 0 ? p : -1;
  __syncthreads();
  const int i = p / P3_BOARD_LEN, j = p % P3_BOARD_LEN;
  for (int iter = 0; iter < P3_NUM_BOARD_LOCS; ++iter) {
    if (threadIdx.x == 0) s_changed = 0;
    __syncthreads();
    if (live && s_board[p] != 0) {
      int best = s_label[p];
      const int8_t c = s_board[p];
      if (i > 0 && s_board[p - 19] == c) best = min(best, s_label[p - 19]);
      if (i < 18 && s_board[p + 19] == c) best = min(best, s_label[p + 19]);
      if (j > 0 && s_board[p - 1] == c) best = min(best, s_label[p - 1]);
      if (j < 18 && s_board[p + 1] == c) best = min(best, s_label[p + 1]);
      best = min(best, s_label[best]);  // pointer jump
      if (best < s_label[p]) {
        atomicMin(&s_label[p], best);
        s_changed = 1;
      }
    }
    __syncthreads();
    const int ch = s_changed;
    __syncthreads();
    if (!ch) break;
  }
}

__device__ __forceinline__ void group_liberties(const int8_t* s_board, const int* s_label, int* s_libs, int p,
                                                bool live) {
  if (live) s_libs[p] = 0;
  __syncthreads();
  if (live && s_board[p] == 0) {
    const int i = p / P3_BOARD_LEN, j = p % P3_BOARD_LEN;
    int roots[4];
    int nr = 0;
    auto add = [&](int q) {
      if (s_board[q] == 0) return;
      const int r = s_label[q];
      for (int k = 0; k < nr; ++k)
        if (roots[k] == r) return;
      roots[nr++] = r;
    };
    if (i > 0) add(p - 19);
    if (i < 18) add(p + 19);
    if (j > 0) add(p - 1);
    if (j < 18) add(p + 1);
    for (int k = 0; k < nr; ++k) atomicAdd(&s_libs[roots[k]], 1);
  }
  __syncthreads();
}

__global__ void __launch_bounds__(384) liberties_kernel(const int8_t* __restrict__ boards, int n,
                                                        int8_t* __restrict__ out) {
  __shared__ int8_t s_board[P3_NUM_BOARD_LOCS];
  __shared__ int s_label[P3_NUM_BOARD_LOCS];
  __shared__ int s_libs[P3_NUM_BOARD_LOCS];
  const int b = blockIdx.x, p = threadIdx.x;
  const bool live = p < P3_NUM_BOARD_LOCS;
  if (live) s_board[p] = boards[static_cast<size_t>(b) * P3_NUM_BOARD_LOCS + p];
  __syncthreads();
  group_labels(s_board, s_label, p, live);
  group_liberties(s_board, s_label, s_libs, p, live);
  if (live) {
    const int8_t c = s_board[p];
    const int libs = c != 0 ? s_libs[s_label[p]] : -1;
    int8_t* o = out + static_cast<size_t>(b) * 3 * P3_NUM_BOARD_LOCS;
    o[p] = libs == 1 ? c : 0;
    o[P3_NUM_BOARD_LOCS + p] = libs == 2 ? c : 0;
    o[2 * P3_NUM_BOARD_LOCS + p] = libs == 3 ? c : 0;
  }
}

__global__ void __launch_bounds__(384) legal_kernel(const int8_t* __restrict__ boards,
                                                    const int8_t* __restrict__ colors,
                                                    const int8_t* __restrict__ forbidden, int n,
                                                    uint8_t* __restrict__ out) {
  __shared__ int8_t s_board[P3_NUM_BOARD_LOCS];
  __shared__ int s_label[P3_NUM_BOARD_LOCS];
  __shared__ int s_libs[P3_NUM_BOARD_LOCS];
  const int b = blockIdx.x, p = threadIdx.x;
  const bool live = p < P3_NUM_BOARD_LOCS;
  if (live) s_board[p] = boards[static_cast<size_t>(b) * P3_NUM_BOARD_LOCS + p];
  __syncthreads();
  group_labels(s_board, s_label, p, live);
  group_liberties(s_board, s_label, s_libs, p, live);
  const int8_t color = colors[b];
  uint8_t* o = out + static_cast<size_t>(b) * P3_MAX_MOVES;
  if (live) {
    bool ok = false;
    if (s_board[p] == 0 && !(forbidden && forbidden[static_cast<size_t>(b) * P3_NUM_BOARD_LOCS + p])) {
      const int i = p / P3_BOARD_LEN, j = p % P3_BOARD_LEN;
      auto nb = [&](int q) {
        const int8_t c = s_board[q];
        if (c == 0) ok = true;                                                         // empty neighbour
        else if (c == static_cast<int8_t>(-color) && s_libs[s_label[q]] == 1) ok = true;  // capture (:611-613)
        else if (c == color && s_libs[s_label[q]] > 1) ok = true;                        // not self-capture (:901-915)
      };
      if (i > 0) nb(p - 19);
      if (i < 18) nb(p + 19);
      if (j > 0) nb(p - 1);
      if (j < 18) nb(p + 1);
    }
    o[p] = ok ? 1 : 0;
  }
  if (p == P3_NUM_BOARD_LOCS) o[P3_NUM_BOARD_LOCS] = 1;  // pass is always legal (:596-599)
}

}  // namespace

int encode_launch(const p3_go_features* feats, int n, int version, float* planes, float* scalars, uint16_t* masks,
                  cudaStream_t stream, const EncodeExtra* extra) {
  if (n <= 0) return P3_OK;
  encode_kernel<<<n, 256, 0, stream>>>(feats, n, version, planes, scalars, masks, extra ? *extra : EncodeExtra());
  P3_CUDA(cudaGetLastError());
  return P3_OK;
}

int liberties_launch(const int8_t* boards, int n, int8_t* out, cudaStream_t stream) {
  if (n <= 0) return P3_OK;
  liberties_kernel<<<n, 384, 0, stream>>>(boards, n, out);
  P3_CUDA(cudaGetLastError());
  return P3_OK;
}

int legal_mask_launch(const int8_t* boards, const int8_t* colors, const int8_t* forbidden, int n, uint8_t* out,
                      cudaStream_t stream) {
  if (n <= 0) return P3_OK;
  legal_kernel<<<n, 384, 0, stream>>>(boards, colors, forbidden, n, out);
  P3_CUDA(cudaGetLastError());
  return P3_OK;
}

}  // namespace p3
