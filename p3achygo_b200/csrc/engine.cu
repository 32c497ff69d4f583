// libp3b200 engine: the C ABI of include/p3_b200.h over the kernels in this directory.
//
// One p3_engine = one evaluator on one GPU (the reference runs one engine per process per GPU,
// python/rl_loop/selfplay.py:51-64): weights resident in HBM, a pinned host staging area the
// LoadBatch threads write (1860 B of game state per slot instead of the reference's 21 692 B of fp32
// planes, cc/nn/engine/trt_engine.cc:222-236), and one stream that runs
//     H2D(game state) -> encode -> init conv -> residual tower -> head conv -> heads -> D2H(results)
// The device-side sequence is captured once into a CUDA graph and replayed (as the reference does with
// its TensorRT graph, trt_engine.cc:168-220).
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <atomic>
#include <cstring>
#include <memory>
#include <new>
#include <stdexcept>
#include <mutex>
#include <string>
#include <vector>

#include "common.cuh"
#include "math.cuh"
#include "weights_file.h"

namespace p3 {
int init_tc_pack_weights(const float* wt, int nplanes, int C, std::vector<__nv_bfloat16>& out, bool op_f16);  // init_tc.cu
int init_tc2_pack_weights(const float* wt, int nplanes, int C, std::vector<__nv_bfloat16>& out, bool op_f16);  // init_tc2.cu

static thread_local std::string g_last_error;
void set_error(const std::string& msg) { g_last_error = msg; }
int fail(int code, const std::string& msg) {
  g_last_error = msg;
  return code;
}

namespace {

int check_device(int device) {
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) {
    cudaGetLastError();
    return fail(P3_ERR_NO_DEVICE, std::string("no CUDA device available (") +
                                      (e != cudaSuccess ? cudaGetErrorString(e) : "device count 0") +
                                      "); libp3b200 has no CPU fallback");
  }
  if (device < 0 || device >= count) return fail(P3_ERR_INVALID_ARG, "device index out of range");
  P3_CUDA(cudaSetDevice(device));
  return P3_OK;
}

struct DevBuf {
  void* p = nullptr;
  size_t bytes = 0;
  ~DevBuf() { if (p) cudaFree(p); }
  int alloc(size_t n) {
    bytes = n;
    P3_CUDA(cudaMalloc(&p, n ? n : 1));
    return P3_OK;
  }
  template <typename T> T* as() const { return reinterpret_cast<T*>(p); }
};

int upload(DevBuf& buf, const void* host, size_t bytes) {
  int rc = buf.alloc(bytes);
  if (rc) return rc;
  P3_CUDA(cudaMemcpy(buf.p, host, bytes, cudaMemcpyHostToDevice));
  return P3_OK;
}
int upload_f32(DevBuf& buf, const std::vector<float>& v) { return upload(buf, v.data(), v.size() * sizeof(float)); }
int upload_bf16(DevBuf& buf, const std::vector<float>& v) {
  std::vector<__nv_bfloat16> h(v.size());
  for (size_t i = 0; i < v.size(); ++i) h[i] = __float2bfloat16(v[i]);
  return upload(buf, h.data(), h.size() * sizeof(__nv_bfloat16));
}
// 16-bit tensor-core operands: bf16, or IEEE fp16 (P3_PRECISION_FP16; values beyond +-65504 saturate like the activations do)
int upload_op16(DevBuf& buf, const std::vector<float>& v, bool f16) {
  if (!f16) return upload_bf16(buf, v);
  std::vector<__half> h(v.size());
  for (size_t i = 0; i < v.size(); ++i) h[i] = __float2half_rn(std::min(std::max(v[i], -65504.0f), 65504.0f));
  return upload(buf, h.data(), h.size() * sizeof(__half));
}

// conv kernel OIHW [cout][cin][k][k] -> tap-major copies; tap = i*k + j <-> (dy, dx) = (i - k/2, j - k/2)
// (cross-correlation with "same" padding, Keras Conv2D / python/model.py:108-117).
void conv_repack(const WeightTensor& w, std::vector<float>& tkn /*[taps][cin][cout]*/,
                 std::vector<float>& tnk /*[taps][cout][cin]*/) {
  const int cout = w.dims[0], cin = w.dims[1], k = w.dims[2], taps = k * k;
  tkn.assign(static_cast<size_t>(taps) * cin * cout, 0.0f);
  tnk.assign(static_cast<size_t>(taps) * cin * cout, 0.0f);
  for (int o = 0; o < cout; ++o)
    for (int c = 0; c < cin; ++c)
      for (int t = 0; t < taps; ++t) {
        const float v = w.data[(static_cast<size_t>(o) * cin + c) * taps + t];
        tkn[(static_cast<size_t>(t) * cin + c) * cout + o] = v;
        tnk[(static_cast<size_t>(t) * cout + o) * cin + c] = v;
      }
}
std::vector<int> tap_offsets(int k) {
  std::vector<int> off;
  for (int i = 0; i < k; ++i)
    for (int j = 0; j < k; ++j) off.push_back((i - k / 2) * kRowPitch + (j - k / 2));
  return off;
}

struct ConvLayer {
  int cin = 0, cout = 0, ksize = 1, taps = 1;
  std::vector<int> tap_off;
  DevBuf w_f32, w_bf16;
  DevBuf in_scale, in_shift;  // folded BN applied to this layer's INPUT (consumed by the producer's epilogue)
  bool has_bn = false;
  TcConvPlan* plan = nullptr;
  ~ConvLayer() { if (plan) tc_conv_plan_destroy(plan); }
};

enum StepKind { kStepConv, kStepBroadcast, kStepChain };
struct Step {
  StepKind kind = kStepConv;
  ConvLayer* layer = nullptr;
  const void* in = nullptr;
  ConvEpilogue ep;
  // broadcast
  const float* bw = nullptr;
  const float* bb = nullptr;
  void* b_out = nullptr;
  const float* b_scale = nullptr;
  const float* b_shift = nullptr;
  TcBcastPlan* bplan = nullptr;  // tcgen05 broadcast mix (bf16 mode)
  // fused block boundary (bf16 mode): `layer` = the block's expand conv, `layer2` = the next block's reduce conv
  ConvLayer* layer2 = nullptr;
  TcChainPlan* cplan = nullptr;
  bool tail = false;  // cplan is the tail form: last expand + the heads' 1x1 conv
  TcPwPlan* pplan = nullptr;  // stand-alone 1x1 layer on the CTA-pair kernel (pw_tc.cu) instead of layer->plan
};

}  // namespace
}  // namespace p3

using namespace p3;

struct p3_engine {
  std::string path;
  int device = 0, batch = 0, version = 1, precision = P3_PRECISION_FP32;
  bool bf16 = false;  // tensor-core path (16-bit operands: bf16, or fp16 when `f16`)
  bool f16 = false;   // P3_PRECISION_FP16: the operands are IEEE fp16
  int C = 0, Cb = 0, Ch = 0, Cv = 0, blocks = 0, nplanes = 15, nscalars = 8;
  double flops_per_pos = 0.0;
  int rows = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};

  // host staging (pinned)
  p3_go_features* h_feats = nullptr;
  p3_infer_result* h_results = nullptr;
  int8_t* h_sym = nullptr;  // per-slot game::Symmetry of p3_engine_load_batch_sym (0 = features already oriented by the caller)
  // device IO
  DevBuf d_feats, d_planes, d_scalars, d_masks, d_results, d_aux, d_leaf;
  p3_leaf_result* h_leaf = nullptr;  // pinned: the serial path's compact results (P3_RESULT_LEAF)
  int result_mode = P3_RESULT_FULL;
  // activations
  DevBuf xraw, actA, actB, actS0, actS1, rawB, pgv, head_w_pad;
  bool head_fused = false;  // the tower's last expand and the heads' 1x1 conv run as one launch (chain_tc.cu, tail form)
  // weights
  DevBuf init_wt, init_wt_bf16, gs_w, gs_b, ident_scale, ident_shift;
  bool init_smem = false;  // init conv with the bf16 weight table resident in shared memory
  bool init_tc = false;    // init conv as a tcgen05 implicit GEMM over the plane masks (init_tc.cu / init_tc2.cu)
  bool init_form2 = false; // ... in its shifted-view form (init_tc2.cu)
  DevBuf init_wt_tc, d_masks_pad, d_gs, d_sym;
  InitTcPlan* init_plan = nullptr;
  InitTc2Plan* init_plan2 = nullptr;  // the shifted-view form of the first layer (init_tc2.cu); P3_INIT_TC=1 keeps the first form
  EncodeExtra enc_extra;
  std::vector<std::unique_ptr<ConvLayer>> layers;
  std::vector<DevBuf*> owned;
  std::vector<std::unique_ptr<DevBuf>> misc;
  HeadWeights hw{};
  ConvLayer* head_conv = nullptr;
  const float* first_scale = nullptr;
  const float* first_shift = nullptr;
  std::vector<Step> program;
  Step head_step;

  bool use_graph = true;
  cudaGraphExec_t graph_exec = nullptr;
  cudaGraphExec_t graph_exec_host = nullptr;  // same step with the heads kernel writing NNInferResult straight to pinned host memory
  bool results_to_host = true;                // RunInference: no separate D2H copy (P3_RESULTS_TO_HOST=0 restores it)
  float stage_ms[3] = {0, 0, 0};
  int launches = 0;
  bool aux_host_valid = false;

  // Slot banks of the pipelined form (p3_engine_submit / p3_engine_wait): bank 0 shares the pinned staging of the
  // serial calls, bank 1 has its own.  Copies run on two copy streams so that the H2D of one bank and the D2H of the
  // other overlap the tower of the bank in flight; the kernels of both banks share `stream` and the activation buffers.
  struct Bank {
    p3_go_features* h_feats = nullptr;
    p3_infer_result* h_results = nullptr;
    int8_t* h_sym = nullptr;
    bool owns_host = false;
    DevBuf d_feats, d_sym, d_results, d_aux, d_leaf;   // d_aux / d_leaf: this bank's copy of the step's aux + leaf records
    p3_leaf_result* h_leaf = nullptr;                    // pinned (P3_RESULT_LEAF)
    int mode = P3_RESULT_FULL;                           // result mode of the run in flight / last completed
    bool on_device = false;                              // d_results / d_aux hold the last completed run (a submit, not a serial run)
    // slots loaded as game records (p3_engine_load_game_bank): move lists, move counts (-1 = the slot holds GoFeatures), pass-alive grids
    int16_t* h_moves = nullptr;
    int32_t* h_nmoves = nullptr;
    int8_t* h_forbidden = nullptr;
    int32_t* h_gstatus = nullptr;   // per-slot status of the last derivation (0 = ok), read back with the run
    bool derived = false;
    // A slot may be (re)loaded while a run is copying the bank (the LoadBatch contract: no lock, may overlap RunInference).  The
    // three record arrays travel in separate DMAs, so such a slot can reach the GPU torn and be rejected by the replay; its result
    // is garbage by contract and nobody reads it.  Loads bracket their writes with two increments of gen[slot] (odd = being
    // written); a run snapshots gen and num_moves when it is enqueued and only reports a rejected record for slots whose
    // generation is even and unchanged when the run has completed.
    std::unique_ptr<std::atomic<uint32_t>[]> gen;
    std::vector<uint32_t> sub_gen;
    std::vector<int32_t> sub_nmoves;
    DevBuf d_moves, d_nmoves, d_forbidden;
    cudaEvent_t ev_h2d = nullptr, ev_done = nullptr, ev_d2h = nullptr;
    std::atomic<int> in_flight{0};
  };
  Bank banks[P3_NUM_BANKS];
  cudaStream_t h2d_stream = nullptr, d2h_stream = nullptr;
  // rules from game records (ladder.cu), allocated at the first run that has a game-record slot
  LadderWorkspace* ladder_ws = nullptr;
  DevBuf g_boards, g_laddered, g_libs, g_status;
  std::mutex submit_mu;  // one bank's enqueue sequence at a time (two infer threads may submit concurrently)

  ~p3_engine() {
    for (Step& s : program) {
      if (s.bplan) tc_broadcast_plan_destroy(s.bplan);
      if (s.cplan) tc_chain_plan_destroy(s.cplan);
      if (s.pplan) tc_pw_plan_destroy(s.pplan);
    }
    if (init_plan) init_tc_plan_destroy(init_plan);
    if (init_plan2) init_tc2_plan_destroy(init_plan2);
    if (graph_exec) cudaGraphExecDestroy(graph_exec);
    if (graph_exec_host) cudaGraphExecDestroy(graph_exec_host);
    for (auto& e : ev) if (e) cudaEventDestroy(e);
    for (Bank& b : banks) {
      if (b.ev_h2d) cudaEventDestroy(b.ev_h2d);
      if (b.ev_done) cudaEventDestroy(b.ev_done);
      if (b.ev_d2h) cudaEventDestroy(b.ev_d2h);
      if (b.h_moves) cudaFreeHost(b.h_moves);
      if (b.h_nmoves) cudaFreeHost(b.h_nmoves);
      if (b.h_forbidden) cudaFreeHost(b.h_forbidden);
      if (b.h_gstatus) cudaFreeHost(b.h_gstatus);
      if (b.h_leaf) cudaFreeHost(b.h_leaf);
      if (b.owns_host) {
        if (b.h_feats) cudaFreeHost(b.h_feats);
        if (b.h_results) cudaFreeHost(b.h_results);
        if (b.h_sym) cudaFreeHost(b.h_sym);
      }
    }
    if (ladder_ws) ladder_workspace_destroy(ladder_ws);
    if (h2d_stream) cudaStreamDestroy(h2d_stream);
    if (d2h_stream) cudaStreamDestroy(d2h_stream);
    if (stream) cudaStreamDestroy(stream);
    if (h_feats) cudaFreeHost(h_feats);
    if (h_results) cudaFreeHost(h_results);
    if (h_sym) cudaFreeHost(h_sym);
    if (h_leaf) cudaFreeHost(h_leaf);
  }

  const float* dev_vec(const std::vector<float>& v, int* rc) {
    misc.emplace_back(new DevBuf());
    int r = upload_f32(*misc.back(), v);
    if (r && rc) *rc = r;
    return misc.back()->as<float>();
  }

  int run_conv(const Step& s) {
    ConvLayer& L = *s.layer;
    if (s.pplan) return tc_pw_launch(s.pplan, stream);
    if (bf16)
      return tc_conv_launch(L.plan, stream);
    return conv_fp32_launch(reinterpret_cast<const float*>(s.in), L.w_f32.as<float>(), rows, L.cin, L.cout, L.taps,
                            L.tap_off.data(), s.ep, stream);
  }

  int run_init() {
    if (init_plan2) return init_tc2_launch(init_plan2, stream);
    if (init_tc) return init_tc_launch(init_plan, stream);
    if (init_smem)
      return init_conv_smem_launch(d_masks.as<uint16_t>(), d_scalars.as<float>(), batch, nplanes, nscalars, C,
                                   init_wt_bf16.as<__nv_bfloat16>(), gs_w.as<float>(), gs_b.as<float>(), xraw.as<__half>(),
                                   actA.as<__nv_bfloat16>(), first_scale, first_shift, stream);
    return init_conv_launch(d_masks.as<uint16_t>(), d_scalars.as<float>(), batch, nplanes, nscalars, C, init_wt.as<float>(),
                            gs_w.as<float>(), gs_b.as<float>(), xraw.p, actA.p, bf16, first_scale, first_shift, stream);
  }

  int run_broadcast(const Step& s) {
    if (s.bplan) return tc_broadcast_launch(s.bplan, stream);
    return broadcast_launch(s.in, s.bw, s.bb, batch, C, s.b_out, bf16, s.b_scale, s.b_shift, stream);
  }

  // encode -> init conv -> tower -> head conv -> heads, all on `stream`; optional stage events
  // to_host: the heads kernel stores the results through the mapped pinned buffer (posted PCIe writes that overlap the
  // kernel) instead of HBM + a D2H copy after the step
  int enqueue_device(bool with_events, bool to_host = false) {
    int rc;
    if (with_events) P3_CUDA(cudaEventRecord(ev[0], stream));
    rc = encode_launch(d_feats.as<p3_go_features>(), batch, version, d_planes.as<float>(), d_scalars.as<float>(),
                       d_masks.as<uint16_t>(), stream, &enc_extra);
    if (rc) return rc;
    if (with_events) P3_CUDA(cudaEventRecord(ev[1], stream));
    rc = run_init();
    if (rc) return rc;
    for (const Step& s : program) {
      if (s.kind == kStepConv) rc = run_conv(s);
      else if (s.kind == kStepChain) rc = tc_chain_launch(s.cplan, stream);
      else rc = run_broadcast(s);
      if (rc) return rc;
    }
    if (with_events) P3_CUDA(cudaEventRecord(ev[2], stream));
    if (!head_fused && (rc = run_conv(head_step))) return rc;
    rc = heads_launch(pgv.as<float>(), batch, hw, to_host ? h_results : d_results.as<p3_infer_result>(), d_aux.as<p3_aux_result>(),
                      stream, !bf16, d_sym.as<int8_t>(), d_leaf.as<p3_leaf_result>());
    if (rc) return rc;
    if (with_events) P3_CUDA(cudaEventRecord(ev[3], stream));
    return P3_OK;
  }

  // Slots of `bk` that were loaded as game records: move lists -> board, liberty grids, laddered stones, last moves, written
  // into the step's GoFeatures buffer (d_feats) on `stream`; `copy_stream` carries the H2D of the lists.  No-op without such slots.
  int enqueue_game_records(Bank& bk, cudaStream_t copy_stream, cudaEvent_t copied) {
    bool any = false;
    for (int b = 0; b < batch; ++b) {  // snapshot BEFORE the copies are enqueued (see Bank::gen)
      bk.sub_gen[b] = bk.gen[b].load(std::memory_order_acquire);
      bk.sub_nmoves[b] = bk.h_nmoves[b];
      any = any || bk.sub_nmoves[b] >= 0;
    }
    bk.derived = any;
    if (!any) return P3_OK;
    int rc;
    if (!ladder_ws) {
      if ((rc = ladder_workspace_create(batch, P3_MAX_GAME_MOVES, &ladder_ws))) return rc;
      if ((rc = g_boards.alloc(static_cast<size_t>(361) * batch)) || (rc = g_laddered.alloc(static_cast<size_t>(361) * batch)) ||
          (rc = g_libs.alloc(static_cast<size_t>(3 * 361) * batch)) || (rc = g_status.alloc(sizeof(int32_t) * batch)))
        return rc;
    }
    if (!bk.d_moves.p) {
      if ((rc = bk.d_moves.alloc(sizeof(int16_t) * P3_MAX_GAME_MOVES * batch)) || (rc = bk.d_nmoves.alloc(sizeof(int32_t) * batch)) ||
          (rc = bk.d_forbidden.alloc(static_cast<size_t>(361) * batch)))
        return rc;
    }
    P3_CUDA(cudaMemcpyAsync(bk.d_moves.p, bk.h_moves, bk.d_moves.bytes, cudaMemcpyHostToDevice, copy_stream));
    P3_CUDA(cudaMemcpyAsync(bk.d_nmoves.p, bk.h_nmoves, bk.d_nmoves.bytes, cudaMemcpyHostToDevice, copy_stream));
    P3_CUDA(cudaMemcpyAsync(bk.d_forbidden.p, bk.h_forbidden, bk.d_forbidden.bytes, cudaMemcpyHostToDevice, copy_stream));
    if (copy_stream != stream) {
      P3_CUDA(cudaEventRecord(copied, copy_stream));
      P3_CUDA(cudaStreamWaitEvent(stream, copied, 0));
    }
    rc = ladder_enqueue(ladder_ws, bk.d_moves.as<int16_t>(), bk.d_nmoves.as<int32_t>(), bk.d_forbidden.as<int8_t>(), nullptr, batch,
                        g_boards.as<int8_t>(), g_laddered.as<int8_t>(), nullptr, g_status.as<int32_t>(), stream, nullptr);
    if (rc) return rc;
    if ((rc = liberties_launch(g_boards.as<int8_t>(), batch, g_libs.as<int8_t>(), stream))) return rc;
    P3_CUDA(cudaMemcpyAsync(bk.h_gstatus, g_status.p, sizeof(int32_t) * batch, cudaMemcpyDeviceToHost, stream));
    return assemble_features_launch(bk.d_moves.as<int16_t>(), bk.d_nmoves.as<int32_t>(), P3_MAX_GAME_MOVES, g_boards.as<int8_t>(),
                                    g_libs.as<int8_t>(), g_laddered.as<int8_t>(), batch, d_feats.as<p3_go_features>(), stream);
  }

  // after the run has completed: a slot whose move list was not a legal game record is an error (the reference CHECK-fails on
  // impossible states); the other slots' results are valid
  // `retry` (optional): set instead of failing when the reader's watchdog fired, so that the caller can run the batch again unsplit
  int check_game_records(Bank& bk, bool* retry = nullptr) {
    if (!bk.derived) return P3_OK;
    const bool watchdog = (bk.h_gstatus[0] & 8) != 0;  // the reader gave up (ladder.cu): the batch's laddered grids are incomplete
    if (watchdog && retry) {
      *retry = true;
      return P3_OK;
    }
    for (int b = 0; b < batch; ++b) {
      const int st = bk.h_gstatus[b] & ~8;
      if (bk.sub_nmoves[b] < 0 || st == 0) continue;
      // a slot that was being (re)loaded while the bank was copied reached the GPU torn: not an error, its result is unread
      if ((bk.sub_gen[b] & 1u) || bk.gen[b].load(std::memory_order_acquire) != bk.sub_gen[b]) continue;
      return fail(P3_ERR_INVALID_ARG, "slot " + std::to_string(b) + ": game record rejected by the replay (status " +
                                          std::to_string(st) + ": 1 = move onto an occupied point, 2 = reader overflow)");
    }
    if (watchdog) return fail(P3_ERR_CUDA, "ladder reader watchdog fired: laddered stones of this batch are incomplete");
    return P3_OK;
  }

  int ensure_graph(bool to_host = false) {
    cudaGraphExec_t& graph_exec = to_host ? this->graph_exec_host : this->graph_exec;
    if (graph_exec || !use_graph) return P3_OK;
    cudaGraph_t graph = nullptr;
    P3_CUDA(cudaStreamBeginCapture(stream, cudaStreamCaptureModeThreadLocal));
    int rc = enqueue_device(false, to_host);
    cudaError_t e = cudaStreamEndCapture(stream, &graph);
    if (rc) {
      if (graph) cudaGraphDestroy(graph);
      return rc;
    }
    if (e != cudaSuccess) return fail(P3_ERR_CUDA, std::string("graph capture: ") + cudaGetErrorString(e));
    e = cudaGraphInstantiate(&graph_exec, graph, 0);
    cudaGraphDestroy(graph);
    if (e != cudaSuccess) return fail(P3_ERR_CUDA, std::string("graph instantiate: ") + cudaGetErrorString(e));
    return P3_OK;
  }

  int enqueue_device_maybe_graph(bool to_host = false) {
    if (use_graph) {
      int rc = ensure_graph(to_host);
      if (rc) return rc;
      P3_CUDA(cudaGraphLaunch(to_host ? graph_exec_host : graph_exec, stream));
      return P3_OK;
    }
    return enqueue_device(false, to_host);
  }
};

namespace p3 {
namespace {

// max |x| and the number of saturated values (|x| == 65504, what cvt.rn.satfinite leaves of anything larger) of an fp16 buffer
__global__ void range_scan_f16_kernel(const __half* __restrict__ x, size_t n, unsigned* __restrict__ max_bits,
                                      unsigned long long* __restrict__ n_sat) {
  float m = 0.0f;
  unsigned long long sat = 0;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const float v = fabsf(__half2float(x[i]));
    m = fmaxf(m, v);
    sat += (v >= 65504.0f || v != v) ? 1u : 0u;
  }
  m = warp_max(m);
  for (int o = 16; o > 0; o >>= 1) sat += __shfl_xor_sync(0xffffffffu, sat, o);
  if ((threadIdx.x & 31) == 0) {
    atomicMax(max_bits, __float_as_uint(m));  // non-negative floats order like their bit patterns
    if (sat) atomicAdd(n_sat, sat);
  }
}
__global__ void range_scan_f32_kernel(const float* __restrict__ x, size_t n, unsigned* __restrict__ max_bits,
                                      unsigned long long* __restrict__ n_bad) {
  float m = 0.0f;
  unsigned long long bad = 0;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const float v = fabsf(x[i]);
    m = fmaxf(m, v == v ? v : 0.0f);
    bad += (v != v || v > 3.0e38f) ? 1u : 0u;
  }
  m = warp_max(m);
  for (int o = 16; o > 0; o >>= 1) bad += __shfl_xor_sync(0xffffffffu, bad, o);
  if ((threadIdx.x & 31) == 0) {
    atomicMax(max_bits, __float_as_uint(m));
    if (bad) atomicAdd(n_bad, bad);
  }
}

struct Builder {
  p3_engine& e;
  const WeightFile& wf;
  std::string err;

  const WeightTensor* need(const std::string& name) {
    const WeightTensor* t = wf.find(name);
    if (!t && err.empty()) err = "weight file lacks tensor " + name;
    return t;
  }

  // folded BN of tag/batch_norm -> (scale, shift)
  bool fold_bn(const std::string& tag, std::vector<float>& scale, std::vector<float>& shift) {
    const WeightTensor *g = need(tag + "/batch_norm/gamma"), *b = need(tag + "/batch_norm/beta"),
                       *m = need(tag + "/batch_norm/moving_mean"), *v = need(tag + "/batch_norm/moving_variance"),
                       *eps = need(tag + "/batch_norm/epsilon");
    if (!g || !b || !m || !v || !eps) return false;
    const size_t n = g->data.size();
    scale.resize(n);
    shift.resize(n);
    for (size_t i = 0; i < n; ++i) {
      // gamma * (x - mean) / sqrt(var + eps) + beta  (python/model.py:231, inference form)
      const float s = g->data[i] / std::sqrt(v->data[i] + eps->data[0]);
      scale[i] = s;
      shift[i] = b->data[i] - m->data[i] * s;
    }
    return true;
  }

  ConvLayer* make_conv(const WeightTensor& w, const std::string& bn_tag, int* rc) {
    e.layers.emplace_back(new ConvLayer());
    ConvLayer* L = e.layers.back().get();
    L->cout = w.dims[0];
    L->cin = w.dims[1];
    L->ksize = w.dims[2];
    L->taps = L->ksize * L->ksize;
    L->tap_off = tap_offsets(L->ksize);
    std::vector<float> tkn, tnk;
    conv_repack(w, tkn, tnk);
    int r = e.bf16 ? upload_op16(L->w_bf16, tnk, e.f16) : upload_f32(L->w_f32, tkn);
    if (r) { *rc = r; return L; }
    if (!bn_tag.empty()) {
      std::vector<float> sc, sh;
      if (!fold_bn(bn_tag, sc, sh)) { *rc = P3_ERR_IO; return L; }
      if ((r = upload_f32(L->in_scale, sc)) || (r = upload_f32(L->in_shift, sh))) { *rc = r; return L; }
      L->has_bn = true;
    }
    return L;
  }
};

std::string block_tag(int i, bool bcast, bool btl, bool nbt = false) {
  char buf[64];
  std::snprintf(buf, sizeof buf, "model/trunk/%02d:%s", i, bcast ? "broadcast_res" : (btl ? "bottleneck_res" : (nbt ? "nbt_res" : "classic_res")));
  return buf;
}
std::string sub_tag(const std::string& base, int j, const char* kind) {
  char buf[32];
  std::snprintf(buf, sizeof buf, "/%02d:%s", j, kind);
  return base + buf;
}

int build_engine(p3_engine& e, const WeightFile& wf) {
  Builder bd{e, wf, ""};
  int rc = P3_OK;
  const int P = e.nplanes = wf.meta_or("ninput_planes", 15);
  e.nscalars = wf.meta_or("ninput_features", 8);
  e.blocks = wf.meta_or("nlayers", 0);
  const int C = e.C = wf.meta_or("nchannels", 0);
  const int Cb = e.Cb = wf.meta_or("nbtl_channels", 0);
  const int Ch = e.Ch = wf.meta_or("nhead_channels", 0);
  const int Cv = e.Cv = wf.meta_or("nval_channels", 0);
  const int nbtl = wf.meta_or("nbtl", 0);
  const int ksz = wf.meta_or("conv_size", 3);
  const int bint = wf.meta_or("broadcast_interval", 1 << 30);
  const bool btl = wf.meta_or("trunk_block_type", 0) == 0;
  const bool nbt = wf.meta_or("trunk_block_type", 0) == 2;  // NbtResidualBlock, python/model.py:431-470
  if (wf.meta_or("board_len", 19) != 19) return fail(P3_ERR_UNSUPPORTED, "only 19x19 boards are supported");
  if (C <= 0 || e.blocks <= 0 || Ch <= 0 || Cv <= 0) return fail(P3_ERR_IO, "weight file metadata incomplete");
  if (e.version == 1 && (P != 15 || e.nscalars != 8)) return fail(P3_ERR_UNSUPPORTED, "feature version 1 needs 15 planes / 8 scalars");
  if (e.version == 0 && (P != 13 || e.nscalars != 7)) return fail(P3_ERR_UNSUPPORTED, "feature version 0 needs 13 planes / 7 scalars");
  if (e.bf16) {
    bool ok = tc_conv_supported(C, C) && tc_conv_supported(C, 3 * Ch);
    if (btl || nbt) ok = ok && tc_conv_supported(C, Cb) && tc_conv_supported(Cb, Cb) && tc_conv_supported(Cb, C);
    if (!ok) return fail(P3_ERR_UNSUPPORTED, "P3_PRECISION_BF16 needs channel counts that are multiples of 64 "
                                             "(and 3*head_channels a multiple of 32); use P3_PRECISION_FP32 for this net");
  }
  const int B = e.batch;
  const size_t R = e.rows = B * kRowsPerPos;
  const size_t esz = e.bf16 ? 2 : 4;

  // ---- IO + activation buffers
  P3_CUDA(cudaMallocHost(reinterpret_cast<void**>(&e.h_feats), sizeof(p3_go_features) * B));
  P3_CUDA(cudaMallocHost(reinterpret_cast<void**>(&e.h_results), sizeof(p3_infer_result) * B));
  std::memset(e.h_feats, 0, sizeof(p3_go_features) * B);
  for (int b = 0; b < B; ++b) { e.h_feats[b].bsize = 19; e.h_feats[b].color = P3_BLACK; }
  std::memset(e.h_results, 0, sizeof(p3_infer_result) * B);
  P3_CUDA(cudaMallocHost(reinterpret_cast<void**>(&e.h_sym), B));
  std::memset(e.h_sym, 0, B);
  if ((rc = e.d_sym.alloc(B))) return rc;
  P3_CUDA(cudaMemset(e.d_sym.p, 0, B));
  e.enc_extra.sym = e.d_sym.as<int8_t>();
  if ((rc = e.d_feats.alloc(sizeof(p3_go_features) * B))) return rc;
  P3_CUDA(cudaMemcpy(e.d_feats.p, e.h_feats, sizeof(p3_go_features) * B, cudaMemcpyHostToDevice));
  if ((rc = e.d_planes.alloc(sizeof(float) * B * 361 * P))) return rc;
  if ((rc = e.d_scalars.alloc(sizeof(float) * B * e.nscalars))) return rc;
  if ((rc = e.d_masks.alloc(sizeof(uint16_t) * B * 361))) return rc;
  if ((rc = e.d_results.alloc(sizeof(p3_infer_result) * B))) return rc;
  P3_CUDA(cudaMemset(e.d_results.p, 0, e.d_results.bytes));  // struct padding travels with the D2H copies
  if ((rc = e.d_aux.alloc(sizeof(p3_aux_result) * B))) return rc;
  if ((rc = e.d_leaf.alloc(sizeof(p3_leaf_result) * B))) return rc;
  P3_CUDA(cudaMallocHost(reinterpret_cast<void**>(&e.h_leaf), sizeof(p3_leaf_result) * B));
  std::memset(e.h_leaf, 0, sizeof(p3_leaf_result) * B);
  for (int k = 0; k < P3_NUM_BANKS; ++k) {
    p3_engine::Bank& bk = e.banks[k];
    if (k == 0) {
      bk.h_feats = e.h_feats, bk.h_results = e.h_results, bk.h_sym = e.h_sym;
    } else {
      bk.owns_host = true;
      P3_CUDA(cudaMallocHost(reinterpret_cast<void**>(&bk.h_feats), sizeof(p3_go_features) * B));
      P3_CUDA(cudaMallocHost(reinterpret_cast<void**>(&bk.h_results), sizeof(p3_infer_result) * B));
      P3_CUDA(cudaMallocHost(reinterpret_cast<void**>(&bk.h_sym), B));
      std::memcpy(bk.h_feats, e.h_feats, sizeof(p3_go_features) * B);
      std::memset(bk.h_results, 0, sizeof(p3_infer_result) * B);
      std::memset(bk.h_sym, 0, B);
    }
    P3_CUDA(cudaMallocHost(reinterpret_cast<void**>(&bk.h_moves), sizeof(int16_t) * P3_MAX_GAME_MOVES * B));
    P3_CUDA(cudaMallocHost(reinterpret_cast<void**>(&bk.h_nmoves), sizeof(int32_t) * B));
    P3_CUDA(cudaMallocHost(reinterpret_cast<void**>(&bk.h_forbidden), static_cast<size_t>(361) * B));
    P3_CUDA(cudaMallocHost(reinterpret_cast<void**>(&bk.h_gstatus), sizeof(int32_t) * B));
    std::memset(bk.h_gstatus, 0, sizeof(int32_t) * B);
    for (int b = 0; b < B; ++b) bk.h_nmoves[b] = -1;
    std::memset(bk.h_forbidden, 0, static_cast<size_t>(361) * B);
    if ((rc = bk.d_feats.alloc(sizeof(p3_go_features) * B))) return rc;
    if ((rc = bk.d_sym.alloc(B))) return rc;
    if ((rc = bk.d_results.alloc(sizeof(p3_infer_result) * B))) return rc;
    if ((rc = bk.d_aux.alloc(sizeof(p3_aux_result) * B)) || (rc = bk.d_leaf.alloc(sizeof(p3_leaf_result) * B))) return rc;
    P3_CUDA(cudaMallocHost(reinterpret_cast<void**>(&bk.h_leaf), sizeof(p3_leaf_result) * B));
    std::memset(bk.h_leaf, 0, sizeof(p3_leaf_result) * B);
    bk.gen.reset(new std::atomic<uint32_t>[B]);
    for (int b = 0; b < B; ++b) bk.gen[b].store(0, std::memory_order_relaxed);
    bk.sub_gen.assign(B, 0);
    bk.sub_nmoves.assign(B, -1);
    P3_CUDA(cudaEventCreateWithFlags(&bk.ev_h2d, cudaEventDisableTiming));
    P3_CUDA(cudaEventCreateWithFlags(&bk.ev_done, cudaEventDisableTiming));
    P3_CUDA(cudaEventCreateWithFlags(&bk.ev_d2h, cudaEventDisableTiming));
  }
  P3_CUDA(cudaStreamCreateWithFlags(&e.h2d_stream, cudaStreamNonBlocking));
  P3_CUDA(cudaStreamCreateWithFlags(&e.d2h_stream, cudaStreamNonBlocking));
  // residual stream: fp32, or IEEE fp16 in the bf16 engine (fp32 accumulation inside every block)
  if ((rc = e.xraw.alloc((e.bf16 ? sizeof(__half) : sizeof(float)) * R * C))) return rc;
  if ((rc = e.actA.alloc(esz * R * C))) return rc;
  if ((rc = e.actB.alloc(esz * R * C))) return rc;
  if (nbt) {  // Cb-wide raw stream of the nested classic blocks (fp16 in the bf16 engine, like the trunk stream)
    if ((rc = e.rawB.alloc((e.bf16 ? sizeof(__half) : sizeof(float)) * R * Cb))) return rc;
    P3_CUDA(cudaMemset(e.rawB.p, 0, e.rawB.bytes));
  }
  if (btl || nbt) {
    if ((rc = e.actS0.alloc(esz * R * Cb))) return rc;
    if ((rc = e.actS1.alloc(esz * R * Cb))) return rc;
    P3_CUDA(cudaMemset(e.actS0.p, 0, e.actS0.bytes));
    P3_CUDA(cudaMemset(e.actS1.p, 0, e.actS1.bytes));
  }
  if ((rc = e.pgv.alloc(sizeof(float) * R * 3 * Ch))) return rc;
  P3_CUDA(cudaMemset(e.xraw.p, 0, e.xraw.bytes));
  P3_CUDA(cudaMemset(e.actA.p, 0, e.actA.bytes));
  P3_CUDA(cudaMemset(e.actB.p, 0, e.actB.bytes));
  P3_CUDA(cudaMemset(e.pgv.p, 0, e.pgv.bytes));

  // ---- init conv (model.py:1152-1161, 1230-1237): OIHW [C][P][5][5] -> [25][P][C]
  {
    const WeightTensor* w = bd.need("model/init_conv/conv/kernel");
    const WeightTensor* gw = bd.need("model/init_game_state/dense/kernel");
    const WeightTensor* gb = bd.need("model/init_game_state/dense/bias");
    if (!w || !gw || !gb) return fail(P3_ERR_IO, bd.err);
    if (w->dims.size() != 4 || static_cast<int>(w->dims[0]) != C || static_cast<int>(w->dims[1]) != P || w->dims[2] != 5)
      return fail(P3_ERR_UNSUPPORTED, "init conv must be 5x5, planes -> channels");
    std::vector<float> wt(static_cast<size_t>(25) * P * C);
    for (int o = 0; o < C; ++o)
      for (int c = 0; c < P; ++c)
        for (int t = 0; t < 25; ++t) wt[(static_cast<size_t>(t) * P + c) * C + o] = w->data[(static_cast<size_t>(o) * P + c) * 25 + t];
    if ((rc = upload_f32(e.init_wt, wt)) || (rc = upload_f32(e.gs_w, gw->data)) || (rc = upload_f32(e.gs_b, gb->data))) return rc;
    const char* env_ic = std::getenv("P3_INIT_SMEM");
    e.init_smem = e.bf16 && !e.f16 && init_conv_smem_supported(P, C) && !(env_ic && std::atoi(env_ic) == 0);
    if (e.init_smem && (rc = upload_bf16(e.init_wt_bf16, wt))) return rc;  // (CUDA-core fallback of the first layer: bf16 table)
    const char* env_it = std::getenv("P3_INIT_TC");
    e.init_tc = e.bf16 && init_tc_supported(P, e.nscalars, C) && !(env_it && std::atoi(env_it) == 0);
    if (e.init_tc) {
      std::vector<__nv_bfloat16> packed;
      const bool form2 = init_tc2_supported(P, e.nscalars, C) && !(env_it && std::atoi(env_it) == 1);
      if (form2) init_tc2_pack_weights(wt.data(), P, C, packed, e.f16);
      else init_tc_pack_weights(wt.data(), P, C, packed, e.f16);
      e.init_form2 = form2;
      if ((rc = upload(e.init_wt_tc, packed.data(), packed.size() * sizeof(__nv_bfloat16)))) return rc;
      if ((rc = e.d_masks_pad.alloc(sizeof(uint16_t) * B * kMaskPadElems)) || (rc = e.d_gs.alloc(sizeof(float) * B * C))) return rc;
      P3_CUDA(cudaMemset(e.d_masks_pad.p, 0, e.d_masks_pad.bytes));  // the grid borders stay zero for the engine's lifetime
      e.enc_extra.masks_padded = e.d_masks_pad.as<uint16_t>();
      e.enc_extra.gs_w = e.gs_w.as<float>();
      e.enc_extra.gs_b = e.gs_b.as<float>();
      e.enc_extra.C = C;
      e.enc_extra.gs_out = e.d_gs.as<float>();
    }
    if (e.f16 && !e.init_tc)
      return fail(P3_ERR_UNSUPPORTED, "P3_PRECISION_FP16 needs the tensor-core first layer (15/13 input planes, channels % 64 == 0, P3_INIT_TC not 0)");
  }
  {
    std::vector<float> ones(std::max(C, 3 * Ch), 1.0f), zeros(std::max(C, 3 * Ch), 0.0f);
    if ((rc = upload_f32(e.ident_scale, ones)) || (rc = upload_f32(e.ident_shift, zeros))) return rc;
  }

  // ---- trunk (model.py:1000-1047): build layers, then wire the program
  struct BlockDesc {
    bool bcast;
    std::vector<ConvLayer*> convs;
    const float* bw = nullptr;
    const float* bb = nullptr;
    const WeightTensor* bw_host = nullptr;
    const WeightTensor* bb_host = nullptr;
  };
  std::vector<BlockDesc> blocks(e.blocks);
  for (int i = 0; i < e.blocks; ++i) {
    const bool bcast = (i % bint) == bint - 1;  // model.py:1003
    blocks[i].bcast = bcast;
    const std::string bt = block_tag(i, bcast, btl, nbt);
    std::vector<int> idx;
    if (bcast) idx = {0, 2};
    else if (btl) for (int j = 0; j < nbtl + 2; ++j) idx.push_back(j);
    else if (nbt) idx = {0, 1, 2, 3, 4, 5};
    else idx = {0, 1};
    for (int j : idx) {
      const std::string tag = sub_tag(bt, j, "conv_block");
      const WeightTensor* w = bd.need(tag + "/conv/kernel");
      if (!w) return fail(P3_ERR_IO, bd.err);
      ConvLayer* L = bd.make_conv(*w, tag, &rc);
      if (rc) return rc == P3_ERR_IO ? fail(rc, bd.err) : rc;
      blocks[i].convs.push_back(L);
    }
    if (bcast) {
      const std::string tag = sub_tag(bt, 1, "broadcast");
      const WeightTensor *w = bd.need(tag + "/dense/kernel"), *b = bd.need(tag + "/dense/bias");
      if (!w || !b) return fail(P3_ERR_IO, bd.err);
      blocks[i].bw = e.dev_vec(w->data, &rc);
      blocks[i].bb = e.dev_vec(b->data, &rc);
      blocks[i].bw_host = w;
      blocks[i].bb_host = b;
      if (rc) return rc;
    }
    (void)ksz;
  }
  // head conv: conv_p | conv_g | conv_v as one [3Ch][C][1][1] kernel (model.py:784-785, 889)
  {
    const WeightTensor *wp = bd.need("model/policy_head/conv_policy/conv/kernel"),
                       *wg = bd.need("model/policy_head/conv_global/conv/kernel"),
                       *wv = bd.need("model/value_head/conv_value/conv/kernel");
    if (!wp || !wg || !wv) return fail(P3_ERR_IO, bd.err);
    WeightTensor cat;
    cat.dims = {static_cast<uint32_t>(3 * Ch), static_cast<uint32_t>(C), 1, 1};
    cat.data.reserve(static_cast<size_t>(3) * Ch * C);
    for (const WeightTensor* t : {wp, wg, wv}) cat.data.insert(cat.data.end(), t->data.begin(), t->data.end());
    e.head_conv = bd.make_conv(cat, "", &rc);
    if (rc) return rc;
  }

  // wire: `cur` holds the activated input of the next block
  void* cur = e.actA.p;
  void* other = e.actB.p;
  e.first_scale = blocks[0].convs[0]->in_scale.as<float>();
  e.first_shift = blocks[0].convs[0]->in_shift.as<float>();
  if (e.init_tc && e.init_form2) {
    if ((rc = init_tc2_plan_create(e.d_masks_pad.as<uint16_t>(), e.d_gs.as<float>(), B, C, e.init_wt_tc.as<__nv_bfloat16>(),
                                   e.xraw.as<__half>(), e.actA.as<__nv_bfloat16>(), e.first_scale, e.first_shift, &e.init_plan2, e.f16)))
      return rc;
  } else if (e.init_tc && (rc = init_tc_plan_create(e.d_masks_pad.as<uint16_t>(), e.d_gs.as<float>(), B, C,
                                                    e.init_wt_tc.as<__nv_bfloat16>(), e.xraw.as<__half>(), e.actA.as<__nv_bfloat16>(),
                                                    e.first_scale, e.first_shift, &e.init_plan, e.f16)))
    return rc;
  const char* env_pw = std::getenv("P3_TC_PW");
  const bool pw_enabled = !(env_pw && std::atoi(env_pw) == 0);
  auto add_conv = [&](ConvLayer* L, const void* in, const void* residual, void* raw, void* act, int mode,
                      const ConvLayer* next) -> int {
    Step s;
    s.kind = kStepConv;
    s.layer = L;
    s.in = in;
    s.ep.residual = residual;
    s.ep.raw_out = raw;
    s.ep.raw_f16 = e.bf16 && (residual != nullptr || raw != nullptr);
    s.ep.act_out = act;
    s.ep.act_mode = mode;
    s.ep.op_f16 = e.f16;
    if (mode == kActMishBN) {
      s.ep.scale = next->in_scale.as<float>();
      s.ep.shift = next->in_shift.as<float>();
    }
    if (e.bf16 && pw_enabled && L->taps == 1 && tc_pw_supported(L->cin, L->cout) && (act != nullptr || raw != nullptr)) {
      int r = tc_pw_plan_create(reinterpret_cast<const __nv_bfloat16*>(in), L->w_bf16.as<__nv_bfloat16>(), e.rows, L->cin,
                                L->cout, s.ep, &s.pplan);
      if (r) return r;
    } else if (e.bf16) {
      int r = tc_conv_plan_create(reinterpret_cast<const __nv_bfloat16*>(in), L->w_bf16.as<__nv_bfloat16>(), e.rows,
                                  L->cin, L->cout, L->taps, L->tap_off.data(), s.ep, &L->plan);
      if (r) return r;
    }
    e.program.push_back(s);
    return P3_OK;
  };
  void* chain_out = nullptr;  // set when the previous block ended with a fused boundary
  bool bcast_conv0_done = false;  // the previous btl block's fused boundary ran the broadcast block's first conv
  const char* env_chain = std::getenv("P3_TC_CHAIN");
  const bool chain_enabled = !(env_chain && std::atoi(env_chain) == 0);
  const char* env_cb = std::getenv("P3_TC_CHAIN_BCAST");
  const bool chain_bcast = !(env_cb && std::atoi(env_cb) == 0);
  const char* env_tail = std::getenv("P3_TC_TAIL");
  const bool tail_enabled = !(env_tail && std::atoi(env_tail) == 0);
  for (int i = 0; i < e.blocks; ++i) {
    BlockDesc& bk = blocks[i];
    const bool last_block = i == e.blocks - 1;
    // bf16 engine: nothing reads the raw stream after the last block (the heads take the identity-activated copy)
    void* raw_dst = (last_block && e.bf16) ? nullptr : e.xraw.p;
    const ConvLayer* next_first = last_block ? nullptr : blocks[i + 1].convs[0];
    // what the block's final conv writes besides the raw residual stream
    const int end_mode = last_block ? (e.bf16 ? kActIdentity : kActNone) : kActMishBN;
    void* end_act = (last_block && !e.bf16) ? nullptr : other;
    if (bk.bcast) {  // BroadcastResidualBlock, model.py:583-607
      // buffers: t0 = output of the first conv (input of the mix), t1 = output of the mix, both C-wide
      void* t0 = other;
      void* t1 = cur;
      if (bcast_conv0_done) {  // the previous block's fused boundary already ran this block's first conv into `cur`
        t0 = cur;
        t1 = other;
        bcast_conv0_done = false;
      } else if ((rc = add_conv(bk.convs[0], cur, nullptr, nullptr, t0, kActMish, nullptr))) {
        return rc;
      }
      Step s;
      s.kind = kStepBroadcast;
      s.in = t0;
      s.bw = bk.bw;
      s.bb = bk.bb;
      s.b_out = t1;
      s.b_scale = bk.convs[1]->in_scale.as<float>();
      s.b_shift = bk.convs[1]->in_shift.as<float>();
      const char* env_tb = std::getenv("P3_TC_BROADCAST");
      if (e.f16 && (!tc_broadcast_supported(C) || (env_tb && std::atoi(env_tb) == 0)))
        return fail(P3_ERR_UNSUPPORTED, "P3_PRECISION_FP16 needs the tensor-core broadcast mix");
      if (e.bf16 && tc_broadcast_supported(C) && !(env_tb && std::atoi(env_tb) == 0)) {
        if ((rc = tc_broadcast_plan_create(bk.bw_host->data.data(), bk.bb_host->data.data(), t0, t1, B, C, s.b_scale,
                                           s.b_shift, &s.bplan, e.f16)))
          return rc;
      }
      e.program.push_back(s);
      // block boundary broadcast -> btl: second conv + residual + next block's reduce as ONE launch (chain_tc.cu)
      ConvLayer* c2 = bk.convs[1];
      ConvLayer* nr = (!last_block && !blocks[i + 1].bcast && btl) ? blocks[i + 1].convs[0] : nullptr;
      if (e.bf16 && chain_enabled && chain_bcast && nr && blocks[i + 1].convs.size() >= 2 && nr->taps == 1 &&
          tc_chain_supported(c2->cin, c2->cout, nr->cout)) {
        Step cs;
        cs.kind = kStepChain;
        cs.layer = c2;
        cs.layer2 = nr;
        cs.in = t1;
        const ConvLayer* n2 = blocks[i + 1].convs[1];
        if ((rc = tc_chain_plan_create(reinterpret_cast<const __nv_bfloat16*>(t1), c2->w_bf16.as<__nv_bfloat16>(),
                                       nr->w_bf16.as<__nv_bfloat16>(), e.rows, c2->cin, c2->cout, nr->cout, e.xraw.p, e.xraw.p,
                                       nr->in_scale.as<float>(), nr->in_shift.as<float>(), e.actS0.p, n2->in_scale.as<float>(),
                                       n2->in_shift.as<float>(), kActMishBN, &cs.cplan, e.f16)))
          return rc;
        e.program.push_back(cs);
        chain_out = e.actS0.p;
      } else {
        // the block's activated output goes to the buffer that is not the mix output; make it `other` (swapped below)
        if (t1 == other) std::swap(cur, other);
        end_act = (last_block && !e.bf16) ? nullptr : other;
        if ((rc = add_conv(c2, cur, e.xraw.p, raw_dst, end_act, end_mode, next_first))) return rc;
      }
    } else if (btl) {  // BottleneckResidualConvBlock, model.py:372-412
      void* s0 = e.actS0.p;
      void* s1 = e.actS1.p;
      const int nc = static_cast<int>(bk.convs.size());
      if (chain_out) {  // the previous block's fused boundary already ran this block's reduce conv into chain_out
        if (chain_out == s1) std::swap(s0, s1);
        chain_out = nullptr;
      } else if ((rc = add_conv(bk.convs[0], cur, nullptr, nullptr, s0, kActMishBN, bk.convs[1]))) {
        return rc;
      }
      for (int j = 1; j < nc - 1; ++j) {
        if ((rc = add_conv(bk.convs[j], s0, nullptr, nullptr, s1, kActMishBN, bk.convs[j + 1]))) return rc;
        std::swap(s0, s1);
      }
      // block boundary btl -> btl: expand + residual + next block's reduce as ONE launch (chain_tc.cu)
      ConvLayer* ex = bk.convs[nc - 1];
      ConvLayer* nr = (!last_block && !blocks[i + 1].bcast) ? blocks[i + 1].convs[0] : nullptr;
      if (e.bf16 && chain_enabled && nr && blocks[i + 1].convs.size() >= 2 && ex->taps == 1 && nr->taps == 1 &&
          tc_chain_supported(ex->cin, ex->cout, nr->cout)) {
        Step s;
        s.kind = kStepChain;
        s.layer = ex;
        s.layer2 = nr;
        s.in = s0;
        const ConvLayer* n2 = blocks[i + 1].convs[1];
        if ((rc = tc_chain_plan_create(reinterpret_cast<const __nv_bfloat16*>(s0), ex->w_bf16.as<__nv_bfloat16>(),
                                       nr->w_bf16.as<__nv_bfloat16>(), e.rows, ex->cin, ex->cout, nr->cout, e.xraw.p, e.xraw.p,
                                       nr->in_scale.as<float>(), nr->in_shift.as<float>(), s1, n2->in_scale.as<float>(),
                                       n2->in_shift.as<float>(), kActMishBN, &s.cplan, e.f16)))
          return rc;
        e.program.push_back(s);
        chain_out = s1;
      } else if (e.bf16 && chain_enabled && chain_bcast && !last_block && blocks[i + 1].bcast && ex->taps == 1 &&
                 tc_chain_supported(ex->cin, ex->cout, blocks[i + 1].convs[0]->cout)) {
        // block boundary btl -> broadcast: expand + residual + the broadcast block's first conv (act = mish, no BN: model.py:574)
        ConvLayer* c0 = blocks[i + 1].convs[0];
        Step cs;
        cs.kind = kStepChain;
        cs.layer = ex;
        cs.layer2 = c0;
        cs.in = s0;
        if ((rc = tc_chain_plan_create(reinterpret_cast<const __nv_bfloat16*>(s0), ex->w_bf16.as<__nv_bfloat16>(),
                                       c0->w_bf16.as<__nv_bfloat16>(), e.rows, ex->cin, ex->cout, c0->cout, e.xraw.p, e.xraw.p,
                                       c0->in_scale.as<float>(), c0->in_shift.as<float>(), other, nullptr, nullptr, kActMish,
                                       &cs.cplan, e.f16)))
          return rc;
        e.program.push_back(cs);
        bcast_conv0_done = true;  // its output is in `other`, which becomes `cur` below
      } else if (e.bf16 && chain_enabled && tail_enabled && last_block && ex->taps == 1 && 3 * Ch <= 128 && ex->cout == C &&
                 tc_chain_supported(ex->cin, ex->cout, 128)) {
        // end of the tower: expand + residual + the heads' 1x1 conv (policy / global-pool / value channels) as ONE launch.  The
        // heads take the raw trunk output (no BN, no mish: model.py:783-786, 887-889), so u = x'; nothing reads x' afterwards.
        // The head conv's 3 * Ch output channels are padded to a 128-wide N tile with zero weight rows.
        if ((rc = e.head_w_pad.alloc(static_cast<size_t>(128) * C * 2))) return rc;
        P3_CUDA(cudaMemset(e.head_w_pad.p, 0, e.head_w_pad.bytes));
        P3_CUDA(cudaMemcpy(e.head_w_pad.p, e.head_conv->w_bf16.p, static_cast<size_t>(3 * Ch) * C * 2, cudaMemcpyDeviceToDevice));
        Step cs;
        cs.kind = kStepChain;
        cs.layer = ex;
        cs.layer2 = e.head_conv;
        cs.in = s0;
        cs.tail = true;
        if ((rc = tc_chain_plan_create(reinterpret_cast<const __nv_bfloat16*>(s0), ex->w_bf16.as<__nv_bfloat16>(),
                                       e.head_w_pad.as<__nv_bfloat16>(), e.rows, ex->cin, ex->cout, 128, e.xraw.p, e.xraw.p, e.first_scale,
                                       e.first_shift, s1, nullptr, nullptr, kActIdentity, &cs.cplan, e.f16, e.pgv.as<float>(), e.rows,
                                       3 * Ch)))
          return rc;
        e.program.push_back(cs);
        e.head_fused = true;
      } else if ((rc = add_conv(ex, s0, e.xraw.p, raw_dst, end_act, end_mode, next_first))) {
        return rc;
      }
    } else if (nbt) {  // NbtResidualBlock, model.py:431-470: 1x1 reduce, two classic blocks at Cb (inner residuals), 1x1 expand
      void* s0 = e.actS0.p;
      void* s1 = e.actS1.p;
      void* tb = e.rawB.p;
      if ((rc = add_conv(bk.convs[0], cur, nullptr, tb, s0, kActMishBN, bk.convs[1]))) return rc;   // t, mish(BN(t))
      if ((rc = add_conv(bk.convs[1], s0, nullptr, nullptr, s1, kActMishBN, bk.convs[2]))) return rc;
      if ((rc = add_conv(bk.convs[2], s1, tb, tb, s0, kActMishBN, bk.convs[3]))) return rc;          // t += nbt_res0(t)
      if ((rc = add_conv(bk.convs[3], s0, nullptr, nullptr, s1, kActMishBN, bk.convs[4]))) return rc;
      if ((rc = add_conv(bk.convs[4], s1, tb, nullptr, s0, kActMishBN, bk.convs[5]))) return rc;     // t += nbt_res1(t), only its activation is needed
      if ((rc = add_conv(bk.convs[5], s0, e.xraw.p, raw_dst, end_act, end_mode, next_first))) return rc;
    } else {  // ClassicResidualBlock, model.py:330-354
      if ((rc = add_conv(bk.convs[0], cur, nullptr, nullptr, other, kActMishBN, bk.convs[1]))) return rc;
      // second conv reads `other`; `cur` is free again and becomes the block output
      if ((rc = add_conv(bk.convs[1], other, e.xraw.p, raw_dst, last_block && !e.bf16 ? nullptr : cur,
                         end_mode, next_first))) return rc;
      continue;  // output already in `cur`
    }
    std::swap(cur, other);
  }
  // heads read the RAW trunk output (no leading BN: model.py:783-786, 887-889)
  {
    const void* head_in = e.bf16 ? cur : e.xraw.p;
    Step s;
    s.kind = kStepConv;
    s.layer = e.head_conv;
    s.in = head_in;
    s.ep.raw_out = e.pgv.as<float>();
    s.ep.op_f16 = e.f16;
    s.ep.raw_transposed = true;  // channel-major [3Ch, rows]: what the heads kernel reads coalesced
    if (e.bf16 && !e.head_fused) {
      int r = tc_conv_plan_create(reinterpret_cast<const __nv_bfloat16*>(head_in), e.head_conv->w_bf16.as<__nv_bfloat16>(),
                                  e.rows, C, 3 * Ch, 1, e.head_conv->tap_off.data(), s.ep, &e.head_conv->plan);
      if (r) return r;
    }
    e.head_step = s;
  }

  // ---- head weights
  {
    HeadWeights& hw = e.hw;
    hw.Ch = Ch;
    hw.Cv = Cv;
    const std::string ph = "model/policy_head", vh = "model/value_head";
    std::vector<float> sc, sh;
    if (!bd.fold_bn(ph + "/global_pool_bias", sc, sh)) return fail(P3_ERR_IO, bd.err);
    hw.gp_scale = e.dev_vec(sc, &rc);
    hw.gp_shift = e.dev_vec(sh, &rc);
    auto T = [&](const std::string& n) { return bd.need(n); };
    const WeightTensor *gdw = T(ph + "/global_pool_bias/dense/kernel"), *gdb = T(ph + "/global_pool_bias/dense/bias"),
                       *mv = T(ph + "/conv_moves/conv/kernel"), *sm = T(ph + "/conv_soft_moves/conv/kernel"),
                       *om = T(ph + "/conv_optimistic_moves/conv/kernel"), *pw = T(ph + "/dense_pass/dense/kernel"),
                       *pb = T(ph + "/dense_pass/dense/bias"), *spw = T(ph + "/dense_soft_pass/dense/kernel"),
                       *spb = T(ph + "/dense_soft_pass/dense/bias"), *opw = T(ph + "/dense_optimistic_pass/dense/kernel"),
                       *opb = T(ph + "/dense_optimistic_pass/dense/bias");
    const WeightTensor *opre = T(vh + "/dense_outcome_pre/dense/kernel"), *opreb = T(vh + "/dense_outcome_pre/dense/bias"),
                       *ow = T(vh + "/dense_outcome/dense/kernel"), *ob = T(vh + "/dense_outcome/dense/bias"),
                       *mw = T(vh + "/dense_mcts_dist/dense/kernel"), *mb = T(vh + "/dense_mcts_dist/dense/bias"),
                       *own = T(vh + "/ownership/conv/kernel"), *gpw = T(vh + "/dense_gamma_pre/dense/kernel"),
                       *gpb = T(vh + "/dense_gamma_pre/dense/bias"), *gw = T(vh + "/dense_gamma/dense/kernel"),
                       *gb = T(vh + "/dense_gamma/dense/bias"), *spre = T(vh + "/dense_scores_pre/dense/kernel"),
                       *spreb = T(vh + "/dense_scores_pre/dense/bias"), *sw = T(vh + "/dense_scores/dense/kernel"),
                       *sb = T(vh + "/dense_scores/dense/bias"), *scores = T(vh + "/scores");
    if (!bd.err.empty()) return fail(P3_ERR_IO, bd.err);
    hw.gp_dense_w = e.dev_vec(gdw->data, &rc);
    hw.gp_dense_b = e.dev_vec(gdb->data, &rc);
    std::vector<float> moves(4 * Ch);  // conv_moves OIHW [2][Ch] rows 0,1; soft row 2; optimistic row 3
    for (int c = 0; c < Ch; ++c) {
      moves[c] = mv->data[c];
      moves[Ch + c] = mv->data[Ch + c];
      moves[2 * Ch + c] = sm->data[c];
      moves[3 * Ch + c] = om->data[c];
    }
    hw.moves_w = e.dev_vec(moves, &rc);
    std::vector<float> passw(2 * Ch * 4), passb(4);
    for (int i = 0; i < 2 * Ch; ++i) {
      passw[i * 4 + 0] = pw->data[i * 2 + 0];
      passw[i * 4 + 1] = pw->data[i * 2 + 1];
      passw[i * 4 + 2] = spw->data[i];
      passw[i * 4 + 3] = opw->data[i];
    }
    passb[0] = pb->data[0] - 3.0f;  // model.py:795
    passb[1] = pb->data[1] - 3.0f;
    passb[2] = spb->data[0] - 3.0f;  // model.py:803
    passb[3] = opb->data[0] - 3.0f;  // model.py:805
    hw.pass_w = e.dev_vec(passw, &rc);
    hw.pass_b = e.dev_vec(passb, &rc);
    hw.outcome_pre_w = e.dev_vec(opre->data, &rc);
    hw.outcome_pre_b = e.dev_vec(opreb->data, &rc);
    hw.outcome_w = e.dev_vec(ow->data, &rc);
    hw.outcome_b = e.dev_vec(ob->data, &rc);
    hw.mcts_w = e.dev_vec(mw->data, &rc);
    hw.mcts_b = e.dev_vec(mb->data, &rc);
    hw.own_w = e.dev_vec(own->data, &rc);
    hw.gamma_pre_w = e.dev_vec(gpw->data, &rc);
    hw.gamma_pre_b = e.dev_vec(gpb->data, &rc);
    hw.gamma_w = e.dev_vec(gw->data, &rc);
    hw.gamma_b = e.dev_vec(gb->data, &rc);
    hw.score_pre_w = e.dev_vec(spre->data, &rc);
    hw.score_pre_b = e.dev_vec(spreb->data, &rc);
    hw.score_w = e.dev_vec(sw->data, &rc);
    hw.score_b = e.dev_vec(sb->data, &rc);
    hw.scores = e.dev_vec(scores->data, &rc);
    if (rc) return rc;
  }

  // algorithmic FLOPs
  {
    const double Pn = 361.0;
    double mac = 25.0 * P * C * Pn + static_cast<double>(e.nscalars) * C;
    for (int i = 0; i < e.blocks; ++i) {
      if (blocks[i].bcast) mac += 2.0 * C * C * Pn + C * Pn * Pn;
      else if (btl) mac += (2.0 * C * Cb + nbtl * 9.0 * Cb * Cb) * Pn;
      else if (nbt) mac += (2.0 * C * Cb + 4 * 9.0 * Cb * Cb) * Pn;
      else mac += 18.0 * C * C * Pn;
    }
    mac += 2.0 * C * Ch * Pn + 4.0 * Ch * Pn + 2.0 * Ch * (Ch + 4);
    mac += C * Ch * Pn + Ch * Pn + 2.0 * 2 * Ch * Cv + Cv * 66.0 + 2.0 * Ch * Cv + 2.0 * 800 * Cv;
    e.flops_per_pos = 2.0 * mac;
  }
  // Tile order: consecutive pair-kernel launches walk the batch in opposite directions, so each starts on the rows its
  // predecessor wrote last (still in L2).  The first layer walks upwards; kernels with a fixed order count as upwards.  P3_TILE_ALTERNATE=0 keeps every launch upwards.
  {
    const char* env_alt = std::getenv("P3_TILE_ALTERNATE");
    const bool alternate = !(env_alt && std::atoi(env_alt) == 0);
    bool prev_up = true;  // the first layer
    for (Step& s : e.program) {
      const bool want_rev = alternate && prev_up;
      bool is_rev = false;
      if (s.kind == kStepChain) {
        tc_chain_plan_set_reverse(s.cplan, want_rev);
        is_rev = want_rev;
      } else if (s.kind == kStepConv && s.pplan) {
        tc_pw_plan_set_reverse(s.pplan, want_rev);
        is_rev = want_rev;
      } else if (s.kind == kStepConv && e.bf16 && s.layer->plan) {
        is_rev = tc_conv_plan_set_reverse(s.layer->plan, want_rev) && want_rev;
      } else if (s.kind == kStepBroadcast && s.bplan) {
        tc_broadcast_plan_set_reverse(s.bplan, want_rev);
        is_rev = want_rev;
      }
      prev_up = !is_rev;
    }
  }
  e.launches = 2 + static_cast<int>(e.program.size()) + (e.head_fused ? 1 : 2);
  return P3_OK;
}

}  // namespace
}  // namespace p3

// =================================================================================================
// C ABI
// =================================================================================================
extern "C" {

const char* p3_last_error(void) { return p3::g_last_error.c_str(); }
const char* p3_version(void) { return "p3achygo-b200 0.1 (sm_100a)"; }

static int engine_create_impl(const char* weights_path, int device, int batch_size, int feat_version, int precision, p3_engine** out);

// No C++ exception crosses the C boundary (a malformed weight file or an allocation failure is an error code, not a terminate).
int p3_engine_create(const char* weights_path, int device, int batch_size, int feat_version, int precision,
                     p3_engine** out) {
  try {
    return engine_create_impl(weights_path, device, batch_size, feat_version, precision, out);
  } catch (const std::bad_alloc&) {
    if (out) *out = nullptr;
    return fail(P3_ERR_IO, "p3_engine_create: out of host memory (malformed weight file?)");
  } catch (const std::exception& ex) {
    if (out) *out = nullptr;
    return fail(P3_ERR_IO, std::string("p3_engine_create: ") + ex.what());
  }
}

static int engine_create_impl(const char* weights_path, int device, int batch_size, int feat_version, int precision,
                              p3_engine** out) {
  if (!weights_path || !out || batch_size <= 0) return fail(P3_ERR_INVALID_ARG, "p3_engine_create: bad argument");
  if (precision != P3_PRECISION_FP32 && precision != P3_PRECISION_BF16 && precision != P3_PRECISION_FP16)
    return fail(P3_ERR_INVALID_ARG, "p3_engine_create: unknown precision");
  if (feat_version != 0 && feat_version != 1) return fail(P3_ERR_INVALID_ARG, "p3_engine_create: feature version must be 0 or 1");
  *out = nullptr;
  WeightFile wf;
  std::string err = wf.load(weights_path);
  if (!err.empty()) return fail(P3_ERR_IO, err);
  int rc = check_device(device);
  if (rc) return rc;
  std::unique_ptr<p3_engine> e(new p3_engine());
  e->path = weights_path;
  e->device = device;
  e->batch = batch_size;
  e->version = feat_version;
  e->precision = precision;
  e->bf16 = precision != P3_PRECISION_FP32;
  e->f16 = precision == P3_PRECISION_FP16;
  if (const char* g = std::getenv("P3_CUDA_GRAPH")) e->use_graph = std::atoi(g) != 0;
  if (const char* g = std::getenv("P3_RESULTS_TO_HOST")) e->results_to_host = std::atoi(g) != 0;
  P3_CUDA(cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking));
  for (auto& ev : e->ev) P3_CUDA(cudaEventCreate(&ev));
  rc = build_engine(*e, wf);
  if (rc) return rc;
  P3_CUDA(cudaDeviceSynchronize());
  // warm-up run (the reference warms up and captures at construction, trt_engine.cc:162-166)
  rc = e->enqueue_device(false);
  if (rc) return rc;
  P3_CUDA(cudaStreamSynchronize(e->stream));
  rc = e->ensure_graph();
  if (rc) return rc;
  *out = e.release();
  return P3_OK;
}

void p3_engine_destroy(p3_engine* e) {
  if (!e) return;
  cudaSetDevice(e->device);
  cudaStreamSynchronize(e->stream);
  cudaStreamSynchronize(e->h2d_stream);
  cudaStreamSynchronize(e->d2h_stream);
  delete e;
}

int p3_engine_load_batch(p3_engine* e, int batch_id, const p3_go_features* features) {
  if (!e || !features || batch_id < 0 || batch_id >= e->batch) return fail(P3_ERR_INVALID_ARG, "load_batch: bad argument");
  e->banks[0].gen[batch_id].fetch_add(1, std::memory_order_acq_rel);
  std::memcpy(&e->h_feats[batch_id], features, sizeof(p3_go_features));
  e->h_sym[batch_id] = 0;
  e->banks[0].h_nmoves[batch_id] = -1;
  e->banks[0].gen[batch_id].fetch_add(1, std::memory_order_release);
  return P3_OK;
}

int p3_engine_load_batch_sym(p3_engine* e, int batch_id, const p3_go_features* features, int sym) {
  if (!e || !features || batch_id < 0 || batch_id >= e->batch || sym < 0 || sym > 7)
    return fail(P3_ERR_INVALID_ARG, "load_batch_sym: bad argument");
  e->banks[0].gen[batch_id].fetch_add(1, std::memory_order_acq_rel);
  std::memcpy(&e->h_feats[batch_id], features, sizeof(p3_go_features));
  e->h_sym[batch_id] = static_cast<int8_t>(sym);
  e->banks[0].h_nmoves[batch_id] = -1;
  e->banks[0].gen[batch_id].fetch_add(1, std::memory_order_release);
  return P3_OK;
}

static int run_inference_once(p3_engine* e) {
  static const bool trace = std::getenv("P3_TIME_RUN") != nullptr;  // perf experiments: where RunInference's wall time goes
  if (trace) P3_CUDA(cudaEventRecord(e->ev[0], e->stream));
  P3_CUDA(cudaMemcpyAsync(e->d_feats.p, e->h_feats, sizeof(p3_go_features) * e->batch, cudaMemcpyHostToDevice, e->stream));
  P3_CUDA(cudaMemcpyAsync(e->d_sym.p, e->h_sym, e->batch, cudaMemcpyHostToDevice, e->stream));
  int rc = e->enqueue_game_records(e->banks[0], e->stream, nullptr);
  if (rc) return rc;
  if (trace) P3_CUDA(cudaEventRecord(e->ev[1], e->stream));
  const bool leaf = e->result_mode == P3_RESULT_LEAF;
  const bool to_host = e->results_to_host && !leaf;
  e->banks[0].mode = e->result_mode;
  e->banks[0].on_device = false;
  rc = e->enqueue_device_maybe_graph(to_host);
  if (rc) return rc;
  if (trace) P3_CUDA(cudaEventRecord(e->ev[2], e->stream));
  if (leaf)
    P3_CUDA(cudaMemcpyAsync(e->h_leaf, e->d_leaf.p, sizeof(p3_leaf_result) * e->batch, cudaMemcpyDeviceToHost, e->stream));
  else if (!to_host)
    P3_CUDA(cudaMemcpyAsync(e->h_results, e->d_results.p, sizeof(p3_infer_result) * e->batch, cudaMemcpyDeviceToHost, e->stream));
  if (trace) P3_CUDA(cudaEventRecord(e->ev[3], e->stream));
  P3_CUDA(cudaStreamSynchronize(e->stream));
  if (trace) {
    float h2d = 0, dev = 0, d2h = 0;
    cudaEventElapsedTime(&h2d, e->ev[0], e->ev[1]);
    cudaEventElapsedTime(&dev, e->ev[1], e->ev[2]);
    cudaEventElapsedTime(&d2h, e->ev[2], e->ev[3]);
    std::fprintf(stderr, "[p3 run] h2d %.1f us  device %.1f us  d2h %.1f us\n", h2d * 1e3f, dev * 1e3f, d2h * 1e3f);
  }
  return P3_OK;
}

int p3_engine_run_inference(p3_engine* e) {
  if (!e) return fail(P3_ERR_INVALID_ARG, "run_inference: null engine");
  P3_CUDA(cudaSetDevice(e->device));
  int rc = run_inference_once(e);
  if (rc) return rc;
  bool retry = false;
  rc = e->check_game_records(e->banks[0], &retry);
  if (!rc && retry) {  // the ladder reader's watchdog fired: the same batch again with the searches unsplit (cannot stall)
    std::fprintf(stderr, "[p3] ladder reader watchdog fired; re-running the batch unsplit\n");
    p3::ladder_workspace_set_unsplit(e->ladder_ws, true);
    rc = run_inference_once(e);
    p3::ladder_workspace_set_unsplit(e->ladder_ws, false);
    if (!rc) rc = e->check_game_records(e->banks[0]);
  }
  if (rc) return rc;
  const char* rc_env = std::getenv("P3_RANGE_CHECK");
  if (rc_env && std::atoi(rc_env) != 0) {  // validation mode: fail loudly instead of evaluating a net whose fp16 residual stream saturates
    float mx = 0.0f;
    long long sat = 0;
    if ((rc = p3_engine_range_check(e, &mx, &sat))) return rc;
    if (sat > 0)
      return fail(P3_ERR_UNSUPPORTED, "residual stream out of range: " + std::to_string(sat) + " values saturated (max |x| " +
                                          std::to_string(mx) + "); run this net with P3_PRECISION_FP32");
  }
  return P3_OK;
}

int p3_engine_get_batch(p3_engine* e, int batch_id, p3_infer_result* result) {
  if (!e || !result || batch_id < 0 || batch_id >= e->batch) return fail(P3_ERR_INVALID_ARG, "get_batch: bad argument");
  if (e->banks[0].mode != P3_RESULT_FULL) return fail(P3_ERR_INVALID_ARG, "get_batch: the last run was in P3_RESULT_LEAF mode (use p3_engine_get_leaf)");
  std::memcpy(result, &e->h_results[batch_id], sizeof(p3_infer_result));
  return P3_OK;
}

// ---- pipelined form: two slot banks ----------------------------------------------------------------------------------

int p3_engine_load_batch_bank(p3_engine* e, int bank, int batch_id, const p3_go_features* features, int sym) {
  if (!e || !features || bank < 0 || bank >= P3_NUM_BANKS || batch_id < 0 || batch_id >= e->batch || sym < 0 || sym > 7)
    return fail(P3_ERR_INVALID_ARG, "load_batch_bank: bad argument");
  p3_engine::Bank& bk = e->banks[bank];
  bk.gen[batch_id].fetch_add(1, std::memory_order_acq_rel);
  std::memcpy(&bk.h_feats[batch_id], features, sizeof(p3_go_features));
  bk.h_sym[batch_id] = static_cast<int8_t>(sym);
  bk.h_nmoves[batch_id] = -1;
  bk.gen[batch_id].fetch_add(1, std::memory_order_release);
  return P3_OK;
}

int p3_engine_load_game_bank(p3_engine* e, int bank, int batch_id, const int16_t* moves, int num_moves, int color, float komi,
                             const int8_t* forbidden, int sym) {
  if (!e || bank < 0 || bank >= P3_NUM_BANKS || batch_id < 0 || batch_id >= e->batch || sym < 0 || sym > 7 || num_moves < 0 ||
      num_moves > P3_MAX_GAME_MOVES || (num_moves > 0 && !moves) || (color != P3_BLACK && color != P3_WHITE))
    return fail(P3_ERR_INVALID_ARG, "load_game_bank: bad argument");
  p3_engine::Bank& bk = e->banks[bank];
  bk.gen[batch_id].fetch_add(1, std::memory_order_acq_rel);  // odd: the record is being written (see Bank::gen)
  if (num_moves > 0) std::memcpy(bk.h_moves + static_cast<size_t>(batch_id) * P3_MAX_GAME_MOVES, moves, sizeof(int16_t) * num_moves);
  if (forbidden) std::memcpy(bk.h_forbidden + static_cast<size_t>(batch_id) * 361, forbidden, 361);
  else std::memset(bk.h_forbidden + static_cast<size_t>(batch_id) * 361, 0, 361);
  p3_go_features& f = bk.h_feats[batch_id];   // colour, komi, bsize travel in the record; the grids and last moves are derived on the GPU
  f.bsize = P3_BOARD_LEN;
  f.color = static_cast<int8_t>(color);
  f.komi = komi;
  bk.h_sym[batch_id] = static_cast<int8_t>(sym);
  bk.h_nmoves[batch_id] = num_moves;
  bk.gen[batch_id].fetch_add(1, std::memory_order_release);
  return P3_OK;
}

// everything p3_engine_submit enqueues for one bank (caller holds submit_mu)
static int submit_enqueue(p3_engine* e, p3_engine::Bank& bk) {
  const size_t fbytes = sizeof(p3_go_features) * e->batch, rbytes = sizeof(p3_infer_result) * e->batch;
  // game state of this bank -> its device staging copy, on the H2D stream (overlaps the other bank's kernels)
  P3_CUDA(cudaMemcpyAsync(bk.d_feats.p, bk.h_feats, fbytes, cudaMemcpyHostToDevice, e->h2d_stream));
  P3_CUDA(cudaMemcpyAsync(bk.d_sym.p, bk.h_sym, e->batch, cudaMemcpyHostToDevice, e->h2d_stream));
  P3_CUDA(cudaEventRecord(bk.ev_h2d, e->h2d_stream));
  // kernels: staging copy -> the step's fixed input buffers, the captured step, results -> the bank's device copy
  P3_CUDA(cudaStreamWaitEvent(e->stream, bk.ev_h2d, 0));
  P3_CUDA(cudaMemcpyAsync(e->d_feats.p, bk.d_feats.p, fbytes, cudaMemcpyDeviceToDevice, e->stream));
  P3_CUDA(cudaMemcpyAsync(e->d_sym.p, bk.d_sym.p, e->batch, cudaMemcpyDeviceToDevice, e->stream));
  int rc = e->enqueue_game_records(bk, e->h2d_stream, bk.ev_h2d);
  if (rc) return rc;
  rc = e->enqueue_device_maybe_graph(false);
  if (rc) return rc;
  bk.mode = e->result_mode;
  bk.on_device = true;
  const size_t lbytes = sizeof(p3_leaf_result) * e->batch;
  P3_CUDA(cudaMemcpyAsync(bk.d_results.p, e->d_results.p, rbytes, cudaMemcpyDeviceToDevice, e->stream));
  P3_CUDA(cudaMemcpyAsync(bk.d_leaf.p, e->d_leaf.p, lbytes, cudaMemcpyDeviceToDevice, e->stream));
  P3_CUDA(cudaMemcpyAsync(bk.d_aux.p, e->d_aux.p, sizeof(p3_aux_result) * e->batch, cudaMemcpyDeviceToDevice, e->stream));
  P3_CUDA(cudaEventRecord(bk.ev_done, e->stream));
  // results -> pinned host memory on the D2H stream (overlaps the next bank's kernels); the compact records only in leaf mode
  P3_CUDA(cudaStreamWaitEvent(e->d2h_stream, bk.ev_done, 0));
  if (bk.mode == P3_RESULT_LEAF) P3_CUDA(cudaMemcpyAsync(bk.h_leaf, bk.d_leaf.p, lbytes, cudaMemcpyDeviceToHost, e->d2h_stream));
  else P3_CUDA(cudaMemcpyAsync(bk.h_results, bk.d_results.p, rbytes, cudaMemcpyDeviceToHost, e->d2h_stream));
  P3_CUDA(cudaEventRecord(bk.ev_d2h, e->d2h_stream));
  return P3_OK;
}

int p3_engine_submit(p3_engine* e, int bank) {
  if (!e || bank < 0 || bank >= P3_NUM_BANKS) return fail(P3_ERR_INVALID_ARG, "submit: bad argument");
  p3_engine::Bank& bk = e->banks[bank];
  if (bk.in_flight.exchange(1, std::memory_order_acq_rel))
    return fail(P3_ERR_INVALID_ARG, "submit: bank already in flight (p3_engine_wait it first)");
  std::lock_guard<std::mutex> lock(e->submit_mu);
  P3_CUDA(cudaSetDevice(e->device));
  return submit_enqueue(e, bk);
}

int p3_engine_wait(p3_engine* e, int bank) {
  if (!e || bank < 0 || bank >= P3_NUM_BANKS) return fail(P3_ERR_INVALID_ARG, "wait: bad argument");
  p3_engine::Bank& bk = e->banks[bank];
  if (!bk.in_flight.load(std::memory_order_acquire)) return fail(P3_ERR_INVALID_ARG, "wait: bank was not submitted");
  P3_CUDA(cudaSetDevice(e->device));
  P3_CUDA(cudaEventSynchronize(bk.ev_d2h));
  bool retry = false;
  int rc = e->check_game_records(bk, &retry);
  if (!rc && retry) {
    // the ladder reader's watchdog fired: this bank again, unsplit, behind whatever the other bank has queued (the bank's host
    // slots are untouched until p3_engine_wait returns; the other bank waits at submit_mu for the duration of the enqueue only,
    // but a submit of it that slips in between would run split again - the flag is per workspace - so hold the lock to the end)
    std::fprintf(stderr, "[p3] ladder reader watchdog fired; re-running bank %d unsplit\n", bank);
    std::lock_guard<std::mutex> lock(e->submit_mu);
    p3::ladder_workspace_set_unsplit(e->ladder_ws, true);
    rc = submit_enqueue(e, bk);
    cudaError_t se = cudaEventSynchronize(bk.ev_d2h);
    p3::ladder_workspace_set_unsplit(e->ladder_ws, false);
    if (!rc && se != cudaSuccess) rc = fail(P3_ERR_CUDA, std::string("wait (unsplit retry): ") + cudaGetErrorString(se));
    if (!rc) rc = e->check_game_records(bk);
  }
  bk.in_flight.store(0, std::memory_order_release);
  return rc;
}

int p3_engine_get_batch_bank(p3_engine* e, int bank, int batch_id, p3_infer_result* result) {
  if (!e || !result || bank < 0 || bank >= P3_NUM_BANKS || batch_id < 0 || batch_id >= e->batch)
    return fail(P3_ERR_INVALID_ARG, "get_batch_bank: bad argument");
  if (e->banks[bank].mode != P3_RESULT_FULL)
    return fail(P3_ERR_INVALID_ARG, "get_batch_bank: the bank's last run was in P3_RESULT_LEAF mode (use p3_engine_get_leaf_bank)");
  std::memcpy(result, &e->banks[bank].h_results[batch_id], sizeof(p3_infer_result));
  return P3_OK;
}

int p3_engine_set_result_mode(p3_engine* e, int mode) {
  if (!e || (mode != P3_RESULT_FULL && mode != P3_RESULT_LEAF)) return fail(P3_ERR_INVALID_ARG, "set_result_mode: bad argument");
  e->result_mode = mode;
  return P3_OK;
}

int p3_engine_get_leaf(p3_engine* e, int batch_id, p3_leaf_result* leaf) {
  if (!e || !leaf || batch_id < 0 || batch_id >= e->batch) return fail(P3_ERR_INVALID_ARG, "get_leaf: bad argument");
  if (e->banks[0].mode != P3_RESULT_LEAF || e->banks[0].on_device)
    return fail(P3_ERR_INVALID_ARG, "get_leaf: the last serial run was not in P3_RESULT_LEAF mode");
  std::memcpy(leaf, &e->h_leaf[batch_id], sizeof(p3_leaf_result));
  return P3_OK;
}

int p3_engine_get_leaf_bank(p3_engine* e, int bank, int batch_id, p3_leaf_result* leaf) {
  if (!e || !leaf || bank < 0 || bank >= P3_NUM_BANKS || batch_id < 0 || batch_id >= e->batch)
    return fail(P3_ERR_INVALID_ARG, "get_leaf_bank: bad argument");
  p3_engine::Bank& bk = e->banks[bank];
  if (bk.mode != P3_RESULT_LEAF || !bk.on_device) return fail(P3_ERR_INVALID_ARG, "get_leaf_bank: the bank's last run was not a P3_RESULT_LEAF submit");
  if (bk.in_flight.load(std::memory_order_acquire)) return fail(P3_ERR_INVALID_ARG, "get_leaf_bank: bank in flight (p3_engine_wait it first)");
  std::memcpy(leaf, &bk.h_leaf[batch_id], sizeof(p3_leaf_result));
  return P3_OK;
}

int p3_engine_get_aux_bank(p3_engine* e, int bank, int batch_id, p3_aux_result* aux) {
  if (!e || !aux || bank < 0 || bank >= P3_NUM_BANKS || batch_id < 0 || batch_id >= e->batch)
    return fail(P3_ERR_INVALID_ARG, "get_aux_bank: bad argument");
  p3_engine::Bank& bk = e->banks[bank];
  if (!bk.on_device || bk.in_flight.load(std::memory_order_acquire))
    return fail(P3_ERR_INVALID_ARG, "get_aux_bank: needs a bank that was submitted and waited");
  P3_CUDA(cudaSetDevice(e->device));
  P3_CUDA(cudaMemcpy(aux, bk.d_aux.as<p3_aux_result>() + batch_id, sizeof(p3_aux_result), cudaMemcpyDeviceToHost));
  return P3_OK;
}

int p3_engine_get_ownership_bank(p3_engine* e, int bank, int batch_id, float own[P3_NUM_BOARD_LOCS]) {
  if (!e || !own || bank < 0 || bank >= P3_NUM_BANKS || batch_id < 0 || batch_id >= e->batch)
    return fail(P3_ERR_INVALID_ARG, "get_ownership_bank: bad argument");
  p3_engine::Bank& bk = e->banks[bank];
  if (!bk.on_device || bk.in_flight.load(std::memory_order_acquire))
    return fail(P3_ERR_INVALID_ARG, "get_ownership_bank: needs a bank that was submitted and waited");
  P3_CUDA(cudaSetDevice(e->device));
  const char* src = reinterpret_cast<const char*>(bk.d_aux.as<p3_aux_result>() + batch_id) + offsetof(p3_aux_result, ownership);
  P3_CUDA(cudaMemcpy(own, src, sizeof(float) * P3_NUM_BOARD_LOCS, cudaMemcpyDeviceToHost));
  return P3_OK;
}

int p3_engine_gumbel_topk_bank(p3_engine* e, int bank, const int32_t* slots, int n, const uint8_t* legal, uint64_t* prng_state,
                               float noise_scaling, int k, int32_t* out_moves, float* out_scores, int32_t* out_kvalid) {
  if (!e || bank < 0 || bank >= P3_NUM_BANKS || !slots || !legal || !prng_state || !out_moves || !out_scores || !out_kvalid || n < 0)
    return fail(P3_ERR_INVALID_ARG, "gumbel_topk_bank: bad argument");
  if (n == 0) return P3_OK;
  p3_engine::Bank& bk = e->banks[bank];
  if (bk.in_flight.load(std::memory_order_acquire)) return fail(P3_ERR_INVALID_ARG, "gumbel_topk_bank: bank in flight (p3_engine_wait it first)");
  for (int i = 0; i < n; ++i)
    if (slots[i] < 0 || slots[i] >= e->batch) return fail(P3_ERR_INVALID_ARG, "gumbel_topk_bank: slot out of range");
  P3_CUDA(cudaSetDevice(e->device));
  // where the last completed run of this bank left its move_logits: the bank's device copy after a submit; after a serial run
  // the leaf records (always written to HBM) hold the same logits
  const float* logits;
  size_t stride;
  if (bk.on_device) {
    logits = reinterpret_cast<const float*>(bk.d_results.p);
    stride = sizeof(p3_infer_result) / sizeof(float);
  } else if (bank == 0) {
    logits = reinterpret_cast<const float*>(e->d_leaf.p);
    stride = sizeof(p3_leaf_result) / sizeof(float);
  } else {
    return fail(P3_ERR_INVALID_ARG, "gumbel_topk_bank: the bank has no completed run");
  }
  DevBuf dsl, dm, dst, dmv, dsc, dkv;
  int rc;
  if ((rc = upload(dsl, slots, sizeof(int32_t) * n)) || (rc = upload(dm, legal, static_cast<size_t>(n) * P3_MAX_MOVES)) ||
      (rc = upload(dst, prng_state, sizeof(uint64_t) * n)) || (rc = dmv.alloc(sizeof(int32_t) * n * k)) ||
      (rc = dsc.alloc(sizeof(float) * n * k)) || (rc = dkv.alloc(sizeof(int32_t) * n)))
    return rc;
  if ((rc = gumbel_launch(logits, dm.as<uint8_t>(), dst.as<uint64_t>(), n, noise_scaling, k, dmv.as<int32_t>(), dsc.as<float>(),
                          dkv.as<int32_t>(), 0, stride, dsl.as<int32_t>(), false)))
    return rc;
  P3_CUDA(cudaMemcpy(out_moves, dmv.p, dmv.bytes, cudaMemcpyDeviceToHost));
  P3_CUDA(cudaMemcpy(out_scores, dsc.p, dsc.bytes, cudaMemcpyDeviceToHost));
  P3_CUDA(cudaMemcpy(out_kvalid, dkv.p, dkv.bytes, cudaMemcpyDeviceToHost));
  P3_CUDA(cudaMemcpy(prng_state, dst.p, dst.bytes, cudaMemcpyDeviceToHost));
  return P3_OK;
}

int p3_engine_get_aux(p3_engine* e, int batch_id, p3_aux_result* aux) {
  if (!e || !aux || batch_id < 0 || batch_id >= e->batch) return fail(P3_ERR_INVALID_ARG, "get_aux: bad argument");
  P3_CUDA(cudaSetDevice(e->device));
  P3_CUDA(cudaMemcpy(aux, e->d_aux.as<p3_aux_result>() + batch_id, sizeof(p3_aux_result), cudaMemcpyDeviceToHost));
  return P3_OK;
}

int p3_engine_get_ownership(p3_engine* e, int batch_id, float own[P3_NUM_BOARD_LOCS]) {
  if (!e || !own || batch_id < 0 || batch_id >= e->batch) return fail(P3_ERR_INVALID_ARG, "get_ownership: bad argument");
  P3_CUDA(cudaSetDevice(e->device));
  const char* src = reinterpret_cast<const char*>(e->d_aux.as<p3_aux_result>() + batch_id) + offsetof(p3_aux_result, ownership);
  P3_CUDA(cudaMemcpy(own, src, sizeof(float) * P3_NUM_BOARD_LOCS, cudaMemcpyDeviceToHost));
  return P3_OK;
}

const char* p3_engine_path(const p3_engine* e) { return e ? e->path.c_str() : ""; }
int p3_engine_batch_size(const p3_engine* e) { return e ? e->batch : 0; }

int p3_engine_get_planes(p3_engine* e, int batch_id, float* planes, float* scalars) {
  if (!e || !planes || !scalars || batch_id < 0 || batch_id >= e->batch) return fail(P3_ERR_INVALID_ARG, "get_planes: bad argument");
  P3_CUDA(cudaSetDevice(e->device));
  const size_t pn = static_cast<size_t>(361) * e->nplanes;
  P3_CUDA(cudaMemcpy(planes, e->d_planes.as<float>() + batch_id * pn, sizeof(float) * pn, cudaMemcpyDeviceToHost));
  P3_CUDA(cudaMemcpy(scalars, e->d_scalars.as<float>() + static_cast<size_t>(batch_id) * e->nscalars,
                     sizeof(float) * e->nscalars, cudaMemcpyDeviceToHost));
  return P3_OK;
}

int p3_engine_run_device(p3_engine* e, float* ms_total) {
  if (!e) return fail(P3_ERR_INVALID_ARG, "run_device: null engine");
  P3_CUDA(cudaSetDevice(e->device));
  P3_CUDA(cudaEventRecord(e->ev[0], e->stream));
  int rc = e->enqueue_device_maybe_graph();
  if (rc) return rc;
  P3_CUDA(cudaEventRecord(e->ev[3], e->stream));
  P3_CUDA(cudaStreamSynchronize(e->stream));
  float ms = 0.0f;
  P3_CUDA(cudaEventElapsedTime(&ms, e->ev[0], e->ev[3]));
  if (ms_total) *ms_total = ms;
  return P3_OK;
}

int p3_engine_upload(p3_engine* e) {
  if (!e) return fail(P3_ERR_INVALID_ARG, "upload: null engine");
  P3_CUDA(cudaSetDevice(e->device));
  P3_CUDA(cudaMemcpyAsync(e->d_feats.p, e->h_feats, sizeof(p3_go_features) * e->batch, cudaMemcpyHostToDevice, e->stream));
  P3_CUDA(cudaMemcpyAsync(e->d_sym.p, e->h_sym, e->batch, cudaMemcpyHostToDevice, e->stream));
  int rc = e->enqueue_game_records(e->banks[0], e->stream, nullptr);   // slots loaded as game records: derive their features now
  if (rc) return rc;
  P3_CUDA(cudaStreamSynchronize(e->stream));
  return e->check_game_records(e->banks[0]);
}

int p3_engine_profile(p3_engine* e, float ms[P3_NUM_KERNEL_CLASSES], int launches[P3_NUM_KERNEL_CLASSES],
                      double flops[P3_NUM_KERNEL_CLASSES], double bytes[P3_NUM_KERNEL_CLASSES]) {
  if (!e || !ms || !launches || !flops) return fail(P3_ERR_INVALID_ARG, "profile: bad argument");
  P3_CUDA(cudaSetDevice(e->device));
  const int n_launch = e->launches;
  std::vector<cudaEvent_t> evs(n_launch + 1);
  for (auto& v : evs) P3_CUDA(cudaEventCreate(&v));
  std::vector<int> cls;
  std::vector<double> fl, by;
  int idx = 0, rc = P3_OK;
  const double B = e->batch, Pn = 361.0, R = e->rows;
  const double esz = e->bf16 ? 2.0 : 4.0, rsz = e->bf16 ? 2.0 : 4.0;  // operand / residual-stream element sizes
  auto rec = [&]() { return cudaEventRecord(evs[idx++], e->stream); };
  P3_CUDA(rec());
  rc = encode_launch(e->d_feats.as<p3_go_features>(), e->batch, e->version, e->d_planes.as<float>(), e->d_scalars.as<float>(),
                     e->d_masks.as<uint16_t>(), e->stream, &e->enc_extra);
  cls.push_back(0); fl.push_back(0.0);
  by.push_back(B * (sizeof(p3_go_features) + 361.0 * e->nplanes * 4 + e->nscalars * 4 + 722 + (e->init_tc ? kMaskPadElems * 2.0 + e->C * 4.0 : 0.0)));
  P3_CUDA(rec());
  if (!rc) rc = e->run_init();
  cls.push_back(1); fl.push_back(2.0 * (25.0 * e->nplanes * e->C * Pn + double(e->nscalars) * e->C) * B);
  by.push_back(B * (e->init_tc ? kMaskPadElems * 2.0 + e->C * 4.0 : 722.0) + R * e->C * (rsz + esz));
  P3_CUDA(rec());
  for (const Step& s : e->program) {
    if (rc) break;
    if (s.kind == kStepConv) {
      rc = e->run_conv(s);
      cls.push_back(s.layer->taps == 1 ? 2 : 3);
      fl.push_back(2.0 * s.layer->taps * double(s.layer->cin) * s.layer->cout * Pn * B);
      by.push_back(R * (s.layer->cin * esz + (s.ep.act_out ? s.layer->cout * esz : 0.0) + (s.ep.residual ? s.layer->cout * rsz : 0.0) +
                        (s.ep.raw_out ? s.layer->cout * rsz : 0.0)));
    } else if (s.kind == kStepChain) {
      rc = tc_chain_launch(s.cplan, e->stream);
      cls.push_back(s.tail ? 5 : 7);  // the tail form is the head conv's launch
      fl.push_back(2.0 * (double(s.layer->cin) * s.layer->cout + double(s.layer2->cin) * s.layer2->cout) * Pn * B);
      if (s.tail) by.push_back(R * (s.layer->cin * esz + s.layer->cout * rsz + s.layer2->cout * 4.0));  // t, x in, fp32 head channels out
      else by.push_back(R * (s.layer->cin * esz + 2.0 * s.layer->cout * rsz + s.layer2->cout * esz));  // t, x in, x' out, next reduce out
    } else {
      rc = e->run_broadcast(s);
      cls.push_back(4); fl.push_back(2.0 * e->C * Pn * Pn * B);
      by.push_back(R * e->C * 2.0 * esz);
    }
    P3_CUDA(rec());
  }
  if (!e->head_fused) {
    if (!rc) rc = e->run_conv(e->head_step);
    cls.push_back(5); fl.push_back(2.0 * e->C * 3.0 * e->Ch * Pn * B);
    by.push_back(R * (e->C * (e->bf16 ? esz : 4.0) + 3.0 * e->Ch * 4.0));
    P3_CUDA(rec());
  }
  if (!rc) rc = heads_launch(e->pgv.as<float>(), e->batch, e->hw, e->d_results.as<p3_infer_result>(), e->d_aux.as<p3_aux_result>(), e->stream, !e->bf16, e->d_sym.as<int8_t>(), e->d_leaf.as<p3_leaf_result>());
  cls.push_back(6); fl.push_back(0.0);
  by.push_back(R * 3.0 * e->Ch * 4.0 + B * (sizeof(p3_infer_result) + sizeof(p3_aux_result) + sizeof(p3_leaf_result)));
  P3_CUDA(rec());
  P3_CUDA(cudaStreamSynchronize(e->stream));
  for (int c = 0; c < P3_NUM_KERNEL_CLASSES; ++c) {
    ms[c] = 0.0f; launches[c] = 0; flops[c] = 0.0;
    if (bytes) bytes[c] = 0.0;
  }
  for (size_t i = 0; i < cls.size() && !rc; ++i) {
    float t = 0.0f;
    cudaEventElapsedTime(&t, evs[i], evs[i + 1]);
    ms[cls[i]] += t;
    launches[cls[i]] += 1;
    flops[cls[i]] += fl[i];
    if (bytes) bytes[cls[i]] += by[i];
    if (std::getenv("P3_PROFILE_VERBOSE")) {  // per-launch listing (perf work)
      const Step* st = (i >= 2 && i - 2 < e->program.size()) ? &e->program[i - 2] : nullptr;
      if (st && st->kind == kStepChain)
        std::fprintf(stderr, "[p3 profile] #%zu class %d chain %d -> %d (+res) -> %d: %.1f us, %.0f MB\n", i, cls[i], st->layer->cin,
                     st->layer->cout, st->layer2->cout, t * 1e3f, by[i] / 1e6);
      else if (st && st->kind == kStepConv)
        std::fprintf(stderr, "[p3 profile] #%zu class %d conv taps=%d cin=%d cout=%d res=%d raw=%d act=%d: %.1f us, %.0f MB\n", i, cls[i],
                     st->layer->taps, st->layer->cin, st->layer->cout, st->ep.residual != nullptr, st->ep.raw_out != nullptr,
                     st->ep.act_out != nullptr, t * 1e3f, by[i] / 1e6);
      else
        std::fprintf(stderr, "[p3 profile] #%zu class %d: %.1f us, %.0f MB\n", i, cls[i], t * 1e3f, by[i] / 1e6);
    }
  }
  for (auto& v : evs) cudaEventDestroy(v);
  return rc;
}

int p3_engine_range_check(p3_engine* e, float* max_abs, long long* n_saturated) {
  if (!e || !max_abs || !n_saturated) return fail(P3_ERR_INVALID_ARG, "range_check: bad argument");
  P3_CUDA(cudaSetDevice(e->device));
  DevBuf acc;
  int rc = acc.alloc(16);
  if (rc) return rc;
  P3_CUDA(cudaMemsetAsync(acc.p, 0, 16, e->stream));
  unsigned* d_max = acc.as<unsigned>();
  unsigned long long* d_sat = reinterpret_cast<unsigned long long*>(acc.as<char>() + 8);
  const size_t n = static_cast<size_t>(e->rows) * e->C;
  auto scan = [&]() {
    if (e->bf16) range_scan_f16_kernel<<<592, 256, 0, e->stream>>>(e->xraw.as<__half>(), n, d_max, d_sat);
    else range_scan_f32_kernel<<<592, 256, 0, e->stream>>>(e->xraw.as<float>(), n, d_max, d_sat);
    return cudaGetLastError();
  };
  rc = encode_launch(e->d_feats.as<p3_go_features>(), e->batch, e->version, e->d_planes.as<float>(), e->d_scalars.as<float>(),
                     e->d_masks.as<uint16_t>(), e->stream, &e->enc_extra);
  if (!rc) rc = e->run_init();
  if (rc) return rc;
  P3_CUDA(scan());
  for (const Step& s : e->program) {  // the residual stream is rewritten in place block by block: look at it after every launch
    if (s.kind == kStepConv) rc = e->run_conv(s);
    else if (s.kind == kStepChain) rc = tc_chain_launch(s.cplan, e->stream);
    else rc = e->run_broadcast(s);
    if (rc) return rc;
    P3_CUDA(scan());
  }
  P3_CUDA(cudaStreamSynchronize(e->stream));
  unsigned char h[16];
  P3_CUDA(cudaMemcpy(h, acc.p, 16, cudaMemcpyDeviceToHost));
  unsigned mb;
  unsigned long long ns;
  std::memcpy(&mb, h, 4);
  std::memcpy(&ns, h + 8, 8);
  std::memcpy(max_abs, &mb, 4);
  *n_saturated = static_cast<long long>(ns);
  return P3_OK;
}

int p3_engine_first_layer(p3_engine* e, void* stream_out, void* act_out, size_t bytes) {
  if (!e || !stream_out || !act_out) return fail(P3_ERR_INVALID_ARG, "first_layer: bad argument");
  if (!e->bf16) return fail(P3_ERR_UNSUPPORTED, "first_layer: 16-bit engines only");
  const size_t need = static_cast<size_t>(e->batch) * kRowsPerPos * e->C * 2;
  if (bytes < need) return fail(P3_ERR_INVALID_ARG, "first_layer: buffers smaller than batch * 400 * C * 2 bytes");
  P3_CUDA(cudaSetDevice(e->device));
  int rc = encode_launch(e->d_feats.as<p3_go_features>(), e->batch, e->version, e->d_planes.as<float>(), e->d_scalars.as<float>(),
                         e->d_masks.as<uint16_t>(), e->stream, &e->enc_extra);
  if (!rc) rc = e->run_init();
  if (rc) return rc;
  P3_CUDA(cudaMemcpyAsync(stream_out, e->xraw.p, need, cudaMemcpyDeviceToHost, e->stream));
  P3_CUDA(cudaMemcpyAsync(act_out, e->actA.p, need, cudaMemcpyDeviceToHost, e->stream));
  P3_CUDA(cudaStreamSynchronize(e->stream));
  return P3_OK;
}

int p3_engine_stage_ms(p3_engine* e, float ms[3]) {
  if (!e || !ms) return fail(P3_ERR_INVALID_ARG, "stage_ms: bad argument");
  P3_CUDA(cudaSetDevice(e->device));
  int rc = e->enqueue_device(true);
  if (rc) return rc;
  P3_CUDA(cudaStreamSynchronize(e->stream));
  for (int i = 0; i < 3; ++i) P3_CUDA(cudaEventElapsedTime(&ms[i], e->ev[i], e->ev[i + 1]));
  return P3_OK;
}

int p3_engine_launches_per_run(const p3_engine* e) { return e ? e->launches : 0; }
double p3_engine_flops_per_position(const p3_engine* e) { return e ? e->flops_per_pos : 0.0; }

int p3_engine_set_cuda_graph(p3_engine* e, int enabled) {
  if (!e) return fail(P3_ERR_INVALID_ARG, "set_cuda_graph: null engine");
  e->use_graph = enabled != 0;
  return P3_OK;
}

// ---- stand-alone kernels -------------------------------------------------------------------------
int p3_encode_features(int device, const p3_go_features* features, int n, int feat_version, float* planes,
                       float* scalars) {
  if (!features || !planes || !scalars || n < 0) return fail(P3_ERR_INVALID_ARG, "encode_features: bad argument");
  int rc = check_device(device);
  if (rc) return rc;
  if (n == 0) return P3_OK;
  const int np = feat_version == 0 ? P3_NUM_PLANES_V0 : P3_NUM_PLANES_V1;
  const int ns = feat_version == 0 ? P3_NUM_SCALARS_V0 : P3_NUM_SCALARS_V1;
  DevBuf df, dp, ds;
  if ((rc = upload(df, features, sizeof(p3_go_features) * n))) return rc;
  if ((rc = dp.alloc(sizeof(float) * n * 361 * np)) || (rc = ds.alloc(sizeof(float) * n * ns))) return rc;
  if ((rc = encode_launch(df.as<p3_go_features>(), n, feat_version, dp.as<float>(), ds.as<float>(), nullptr, 0))) return rc;
  P3_CUDA(cudaMemcpy(planes, dp.p, dp.bytes, cudaMemcpyDeviceToHost));
  P3_CUDA(cudaMemcpy(scalars, ds.p, ds.bytes, cudaMemcpyDeviceToHost));
  return P3_OK;
}

int p3_board_liberties(int device, const int8_t* boards, int n, int8_t* out) {
  if (!boards || !out || n < 0) return fail(P3_ERR_INVALID_ARG, "board_liberties: bad argument");
  int rc = check_device(device);
  if (rc) return rc;
  if (n == 0) return P3_OK;
  DevBuf db, dout;
  if ((rc = upload(db, boards, static_cast<size_t>(n) * 361)) || (rc = dout.alloc(static_cast<size_t>(n) * 3 * 361))) return rc;
  if ((rc = liberties_launch(db.as<int8_t>(), n, dout.as<int8_t>(), 0))) return rc;
  P3_CUDA(cudaMemcpy(out, dout.p, dout.bytes, cudaMemcpyDeviceToHost));
  return P3_OK;
}

int p3_legal_mask(int device, const int8_t* boards, const int8_t* colors, const int8_t* forbidden, int n, uint8_t* out) {
  if (!boards || !colors || !out || n < 0) return fail(P3_ERR_INVALID_ARG, "legal_mask: bad argument");
  int rc = check_device(device);
  if (rc) return rc;
  if (n == 0) return P3_OK;
  DevBuf db, dc, dfb, dout;
  if ((rc = upload(db, boards, static_cast<size_t>(n) * 361)) || (rc = upload(dc, colors, n)) ||
      (rc = dout.alloc(static_cast<size_t>(n) * 362)))
    return rc;
  if (forbidden && (rc = upload(dfb, forbidden, static_cast<size_t>(n) * 361))) return rc;
  if ((rc = legal_mask_launch(db.as<int8_t>(), dc.as<int8_t>(), forbidden ? dfb.as<int8_t>() : nullptr, n, dout.as<uint8_t>(), 0)))
    return rc;
  P3_CUDA(cudaMemcpy(out, dout.p, dout.bytes, cudaMemcpyDeviceToHost));
  return P3_OK;
}

int p3_game_derive(int device, const int16_t* moves, const int32_t* num_moves, int max_moves, const int8_t* forbidden,
                   const int8_t* colors, int n, int8_t* boards, int8_t* laddered, uint8_t* legal, int32_t* status) {
  if (!moves || !num_moves || max_moves <= 0 || n < 0 || !status || (legal && !colors))
    return fail(P3_ERR_INVALID_ARG, "game_derive: bad argument");
  int rc = check_device(device);
  if (rc) return rc;
  if (n == 0) return P3_OK;
  DevBuf dm, dn, dfb, dc, db, dl, dlegal, dst;
  if ((rc = upload(dm, moves, static_cast<size_t>(n) * max_moves * sizeof(int16_t))) ||
      (rc = upload(dn, num_moves, static_cast<size_t>(n) * sizeof(int32_t))) || (rc = dst.alloc(static_cast<size_t>(n) * sizeof(int32_t))))
    return rc;
  if (forbidden && (rc = upload(dfb, forbidden, static_cast<size_t>(n) * 361))) return rc;
  if (colors && (rc = upload(dc, colors, n))) return rc;
  if (boards && (rc = db.alloc(static_cast<size_t>(n) * 361))) return rc;
  if (laddered && (rc = dl.alloc(static_cast<size_t>(n) * 361))) return rc;
  if (legal && (rc = dlegal.alloc(static_cast<size_t>(n) * 362))) return rc;
  rc = ladder_run(dm.as<int16_t>(), dn.as<int32_t>(), max_moves, forbidden ? dfb.as<int8_t>() : nullptr,
                  colors ? dc.as<int8_t>() : nullptr, n, boards ? db.as<int8_t>() : nullptr, laddered ? dl.as<int8_t>() : nullptr,
                  legal ? dlegal.as<uint8_t>() : nullptr, dst.as<int32_t>(), 0);
  if (rc) return rc;
  if (boards) P3_CUDA(cudaMemcpy(boards, db.p, db.bytes, cudaMemcpyDeviceToHost));
  if (laddered) P3_CUDA(cudaMemcpy(laddered, dl.p, dl.bytes, cudaMemcpyDeviceToHost));
  if (legal) P3_CUDA(cudaMemcpy(legal, dlegal.p, dlegal.bytes, cudaMemcpyDeviceToHost));
  P3_CUDA(cudaMemcpy(status, dst.p, dst.bytes, cudaMemcpyDeviceToHost));
  return P3_OK;
}

int p3_gumbel_topk(int device, const float* logits, const uint8_t* legal, uint64_t* prng_state, int n, float noise_scaling,
                   int k, int32_t* out_moves, float* out_scores, int32_t* out_kvalid) {
  if (!logits || !legal || !prng_state || !out_moves || !out_scores || !out_kvalid || n < 0)
    return fail(P3_ERR_INVALID_ARG, "gumbel_topk: bad argument");
  int rc = check_device(device);
  if (rc) return rc;
  if (n == 0) return P3_OK;
  DevBuf dl, dm, dst, dmv, dsc, dkv;
  if ((rc = upload(dl, logits, sizeof(float) * n * 362)) || (rc = upload(dm, legal, static_cast<size_t>(n) * 362)) ||
      (rc = upload(dst, prng_state, sizeof(uint64_t) * n)) || (rc = dmv.alloc(sizeof(int32_t) * n * k)) ||
      (rc = dsc.alloc(sizeof(float) * n * k)) || (rc = dkv.alloc(sizeof(int32_t) * n)))
    return rc;
  if ((rc = gumbel_launch(dl.as<float>(), dm.as<uint8_t>(), dst.as<uint64_t>(), n, noise_scaling, k, dmv.as<int32_t>(),
                          dsc.as<float>(), dkv.as<int32_t>(), 0)))
    return rc;
  P3_CUDA(cudaMemcpy(out_moves, dmv.p, dmv.bytes, cudaMemcpyDeviceToHost));
  P3_CUDA(cudaMemcpy(out_scores, dsc.p, dsc.bytes, cudaMemcpyDeviceToHost));
  P3_CUDA(cudaMemcpy(out_kvalid, dkv.p, dkv.bytes, cudaMemcpyDeviceToHost));
  P3_CUDA(cudaMemcpy(prng_state, dst.p, dst.bytes, cudaMemcpyDeviceToHost));
  return P3_OK;
}

int p3_broadcast_test(int device, int precision, const float* x, const float* w, const float* bias, int n, int C, float* y) {
  if (!x || !w || !bias || !y || n <= 0) return fail(P3_ERR_INVALID_ARG, "broadcast_test: bad argument");
  int rc = check_device(device);
  if (rc) return rc;
  const bool bf16 = precision == P3_PRECISION_BF16;
  if (bf16 && !tc_broadcast_supported(C)) return fail(P3_ERR_UNSUPPORTED, "broadcast_test: C % 64 != 0");
  const size_t R = static_cast<size_t>(n) * kRowsPerPos;
  std::vector<float> xp(R * C, 0.0f);
  for (int b = 0; b < n; ++b)
    for (int p = 0; p < 361; ++p)
      std::memcpy(&xp[(static_cast<size_t>(b) * kRowsPerPos + board_row(p)) * C], &x[(static_cast<size_t>(b) * 361 + p) * C], sizeof(float) * C);
  std::vector<float> ones(C, 1.0f), zeros(C, 0.0f), wv(w, w + 361 * 361), bv(bias, bias + 361);
  DevBuf dx, dy, dsc, dsh, dw, db;
  if ((rc = upload_f32(dsc, ones)) || (rc = upload_f32(dsh, zeros))) return rc;
  std::vector<float> yp(R * C);
  if (bf16) {
    if ((rc = upload_bf16(dx, xp)) || (rc = dy.alloc(R * C * 2))) return rc;
    P3_CUDA(cudaMemset(dy.p, 0xff, dy.bytes));  // halo rows must be overwritten with zeros
    TcBcastPlan* plan = nullptr;
    if ((rc = tc_broadcast_plan_create(w, bias, dx.p, dy.p, n, C, dsc.as<float>(), dsh.as<float>(), &plan))) return rc;
    rc = tc_broadcast_launch(plan, 0);
    cudaError_t se = cudaDeviceSynchronize();
    tc_broadcast_plan_destroy(plan);
    if (rc) return rc;
    if (se != cudaSuccess) return fail(P3_ERR_CUDA, std::string("tc_broadcast kernel: ") + cudaGetErrorString(se));
    std::vector<__nv_bfloat16> hb(R * C);
    P3_CUDA(cudaMemcpy(hb.data(), dy.p, dy.bytes, cudaMemcpyDeviceToHost));
    for (size_t i = 0; i < yp.size(); ++i) yp[i] = __bfloat162float(hb[i]);
  } else {
    if ((rc = upload_f32(dx, xp)) || (rc = dy.alloc(R * C * 4)) || (rc = upload_f32(dw, wv)) || (rc = upload_f32(db, bv))) return rc;
    if ((rc = broadcast_launch(dx.p, dw.as<float>(), db.as<float>(), n, C, dy.p, false, dsc.as<float>(), dsh.as<float>(), 0))) return rc;
    P3_CUDA(cudaDeviceSynchronize());
    P3_CUDA(cudaMemcpy(yp.data(), dy.p, dy.bytes, cudaMemcpyDeviceToHost));
  }
  for (size_t row = 0; row < R; ++row)  // halo rows are part of the contract: they must come back as zeros
    if (!row_is_live(static_cast<int>(row % kRowsPerPos)))
      for (int c = 0; c < C; ++c)
        if (yp[row * C + c] != 0.0f) return fail(P3_ERR_CUDA, "broadcast_test: halo row not zero");
  for (int b = 0; b < n; ++b)
    for (int p = 0; p < 361; ++p)
      std::memcpy(&y[(static_cast<size_t>(b) * 361 + p) * C], &yp[(static_cast<size_t>(b) * kRowsPerPos + board_row(p)) * C], sizeof(float) * C);
  return P3_OK;
}

int p3_block_boundary_test(int device, int fused, const float* t, const float* x, const float* w1, const float* w2,
                           const float* scale1, const float* shift1, const float* scale2, const float* shift2, int n, int k1,
                           int n1, int n2, float* xprime, float* out) {
  if (!t || !x || !w1 || !w2 || !scale1 || !shift1 || !scale2 || !shift2 || !xprime || !out || n <= 0)
    return fail(P3_ERR_INVALID_ARG, "block_boundary_test: bad argument");
  int rc = check_device(device);
  if (rc) return rc;
  if (fused ? !tc_chain_supported(k1, n1, n2) : !(tc_pw_supported(k1, n1) && tc_pw_supported(n1, n2)))
    return fail(P3_ERR_UNSUPPORTED, "block_boundary_test: shape not supported");
  const size_t R = static_cast<size_t>(n) * kRowsPerPos;
  auto pad = [&](const float* src, int C) {  // [n,361,C] -> padded board-row layout [n*400, C]
    std::vector<float> v(R * C, 0.0f);
    for (int b = 0; b < n; ++b)
      for (int p = 0; p < 361; ++p)
        std::memcpy(&v[(static_cast<size_t>(b) * kRowsPerPos + board_row(p)) * C], &src[(static_cast<size_t>(b) * 361 + p) * C], sizeof(float) * C);
    return v;
  };
  const std::vector<float> tp = pad(t, k1), xp = pad(x, n1);
  std::vector<__half> xh(xp.size());
  for (size_t i = 0; i < xp.size(); ++i) xh[i] = __float2half_rn(xp[i]);
  DevBuf dt, dx, dw1, dw2, ds1, dh1, ds2, dh2, du, dout;
  if ((rc = upload_bf16(dt, tp)) || (rc = upload(dx, xh.data(), xh.size() * sizeof(__half))) ||
      (rc = upload_bf16(dw1, std::vector<float>(w1, w1 + static_cast<size_t>(n1) * k1))) ||
      (rc = upload_bf16(dw2, std::vector<float>(w2, w2 + static_cast<size_t>(n2) * n1))) ||
      (rc = upload_f32(ds1, std::vector<float>(scale1, scale1 + n1))) || (rc = upload_f32(dh1, std::vector<float>(shift1, shift1 + n1))) ||
      (rc = upload_f32(ds2, std::vector<float>(scale2, scale2 + n2))) || (rc = upload_f32(dh2, std::vector<float>(shift2, shift2 + n2))) ||
      (rc = du.alloc(R * n1 * 2)) || (rc = dout.alloc(R * n2 * 2)))
    return rc;
  {  // padding rows inside a position's tiles must come back as zeros; its 20 leading padding rows are never written by the
     // position-aligned pair tiles (common.cuh) and keep the zeros every engine buffer starts with
    std::vector<uint16_t> init(R * n2, 0xffffu);
    for (size_t row = 0; row < R; ++row)
      if (row % kRowsPerPos < static_cast<size_t>(kRowBase)) std::fill(init.begin() + row * n2, init.begin() + (row + 1) * n2, uint16_t(0));
    P3_CUDA(cudaMemcpy(dout.p, init.data(), dout.bytes, cudaMemcpyHostToDevice));
  }
  if (fused) {
    TcChainPlan* plan = nullptr;
    if ((rc = tc_chain_plan_create(dt.as<__nv_bfloat16>(), dw1.as<__nv_bfloat16>(), dw2.as<__nv_bfloat16>(), static_cast<int>(R), k1, n1,
                                   n2, dx.p, dx.p, ds1.as<float>(), dh1.as<float>(), dout.p, ds2.as<float>(), dh2.as<float>(),
                                   kActMishBN, &plan)))
      return rc;
    rc = tc_chain_launch(plan, 0);
    cudaError_t se = cudaDeviceSynchronize();
    tc_chain_plan_destroy(plan);
    if (rc) return rc;
    if (se != cudaSuccess) return fail(P3_ERR_CUDA, std::string("tc_chain kernel: ") + cudaGetErrorString(se));
  } else {
    ConvEpilogue e1, e2;
    e1.residual = dx.p; e1.raw_out = dx.p; e1.raw_f16 = true; e1.act_out = du.p; e1.act_mode = kActMishBN;
    e1.scale = ds1.as<float>(); e1.shift = dh1.as<float>();
    e2.act_out = dout.p; e2.act_mode = kActMishBN; e2.scale = ds2.as<float>(); e2.shift = dh2.as<float>();
    TcPwPlan *p1 = nullptr, *p2 = nullptr;
    if ((rc = tc_pw_plan_create(dt.as<__nv_bfloat16>(), dw1.as<__nv_bfloat16>(), static_cast<int>(R), k1, n1, e1, &p1))) return rc;
    if ((rc = tc_pw_plan_create(du.as<__nv_bfloat16>(), dw2.as<__nv_bfloat16>(), static_cast<int>(R), n1, n2, e2, &p2))) {
      tc_pw_plan_destroy(p1);
      return rc;
    }
    rc = tc_pw_launch(p1, 0);
    if (!rc) rc = tc_pw_launch(p2, 0);
    cudaError_t se = cudaDeviceSynchronize();
    tc_pw_plan_destroy(p1);
    tc_pw_plan_destroy(p2);
    if (rc) return rc;
    if (se != cudaSuccess) return fail(P3_ERR_CUDA, std::string("tc_pw kernel: ") + cudaGetErrorString(se));
  }
  std::vector<__half> hx(R * n1);
  std::vector<__nv_bfloat16> ho(R * n2);
  P3_CUDA(cudaMemcpy(hx.data(), dx.p, dx.bytes, cudaMemcpyDeviceToHost));
  P3_CUDA(cudaMemcpy(ho.data(), dout.p, dout.bytes, cudaMemcpyDeviceToHost));
  for (size_t row = 0; row < R; ++row)
    if (!row_is_live(static_cast<int>(row % kRowsPerPos))) {
      for (int c = 0; c < n2; ++c)
        if (__bfloat162float(ho[row * n2 + c]) != 0.0f) return fail(P3_ERR_CUDA, "block_boundary_test: padding row of out not zero");
      for (int c = 0; c < n1; ++c)
        if (__half2float(hx[row * n1 + c]) != 0.0f) return fail(P3_ERR_CUDA, "block_boundary_test: padding row of x' not zero");
    }
  for (int b = 0; b < n; ++b)
    for (int p = 0; p < 361; ++p) {
      const size_t row = static_cast<size_t>(b) * kRowsPerPos + board_row(p);
      for (int c = 0; c < n1; ++c) xprime[(static_cast<size_t>(b) * 361 + p) * n1 + c] = __half2float(hx[row * n1 + c]);
      for (int c = 0; c < n2; ++c) out[(static_cast<size_t>(b) * 361 + p) * n2 + c] = __bfloat162float(ho[row * n2 + c]);
    }
  return P3_OK;
}

int p3_conv_test(int device, int precision, const float* x, const float* w, int n, int cin, int cout, int ksize, float* y) {
  if (!x || !w || !y || n <= 0 || (ksize != 1 && ksize != 3)) return fail(P3_ERR_INVALID_ARG, "conv_test: bad argument");
  int rc = check_device(device);
  if (rc) return rc;
  const bool bf16 = precision == P3_PRECISION_BF16;
  if (bf16 && !tc_conv_supported(cin, cout)) return fail(P3_ERR_UNSUPPORTED, "conv_test: shape not supported by the tcgen05 path");
  const size_t R = static_cast<size_t>(n) * kRowsPerPos;
  std::vector<float> xp(R * cin, 0.0f);
  for (int b = 0; b < n; ++b)
    for (int p = 0; p < 361; ++p)
      std::memcpy(&xp[(static_cast<size_t>(b) * kRowsPerPos + board_row(p)) * cin], &x[(static_cast<size_t>(b) * 361 + p) * cin],
                  sizeof(float) * cin);
  WeightTensor wt;
  wt.dims = {static_cast<uint32_t>(cout), static_cast<uint32_t>(cin), static_cast<uint32_t>(ksize), static_cast<uint32_t>(ksize)};
  wt.data.assign(w, w + static_cast<size_t>(cout) * cin * ksize * ksize);
  std::vector<float> tkn, tnk;
  conv_repack(wt, tkn, tnk);
  std::vector<int> off = tap_offsets(ksize);
  DevBuf dx, dw, dy;
  if ((rc = dy.alloc(sizeof(float) * R * cout))) return rc;
  ConvEpilogue ep;
  ep.raw_out = dy.as<float>();
  if (bf16) {
    if ((rc = upload_bf16(dx, xp)) || (rc = upload_bf16(dw, tnk))) return rc;
    TcConvPlan* plan = nullptr;
    // a 3x3 test layer with only the activated copy requested exercises the resident-weight kernel (identity act),
    // otherwise the streaming kernel writes the raw fp32 sum
    DevBuf dact;
    const bool want_act = std::getenv("P3_CONV_TEST_ACT") != nullptr;
    if (want_act) {
      if ((rc = dact.alloc(sizeof(__nv_bfloat16) * R * cout))) return rc;
      ep.raw_out = nullptr;
      ep.act_out = dact.p;
      ep.act_mode = kActIdentity;
    }
    if ((rc = tc_conv_plan_create(dx.as<__nv_bfloat16>(), dw.as<__nv_bfloat16>(), static_cast<int>(R), cin, cout, ksize * ksize,
                                  off.data(), ep, &plan)))
      return rc;
    rc = tc_conv_launch(plan, 0);
    if (want_act && !rc) {
      cudaDeviceSynchronize();
      std::vector<__nv_bfloat16> ha(R * cout);
      cudaMemcpy(ha.data(), dact.p, dact.bytes, cudaMemcpyDeviceToHost);
      std::vector<float> hf(R * cout);
      for (size_t i = 0; i < hf.size(); ++i) hf[i] = __bfloat162float(ha[i]);
      cudaMemcpy(dy.p, hf.data(), dy.bytes, cudaMemcpyHostToDevice);
    }
    cudaError_t se = cudaDeviceSynchronize();
    tc_conv_plan_destroy(plan);
    if (rc) return rc;
    if (se != cudaSuccess) return fail(P3_ERR_CUDA, std::string("tc_conv kernel: ") + cudaGetErrorString(se));
  } else {
    if ((rc = upload_f32(dx, xp)) || (rc = upload_f32(dw, tkn))) return rc;
    if ((rc = conv_fp32_launch(dx.as<float>(), dw.as<float>(), static_cast<int>(R), cin, cout, ksize * ksize, off.data(), ep, 0)))
      return rc;
    P3_CUDA(cudaDeviceSynchronize());
  }
  std::vector<float> yp(R * cout);
  P3_CUDA(cudaMemcpy(yp.data(), dy.p, dy.bytes, cudaMemcpyDeviceToHost));
  for (int b = 0; b < n; ++b)
    for (int p = 0; p < 361; ++p)
      std::memcpy(&y[(static_cast<size_t>(b) * 361 + p) * cout], &yp[(static_cast<size_t>(b) * kRowsPerPos + board_row(p)) * cout],
                  sizeof(float) * cout);
  return P3_OK;
}

}  // extern "C"
