// Host-side mirror of the slot synchronisation of the reference's nn::NNInterface (cc/nn/nn_interface.{h,cc}), SignalKind
// kAuto, over any nn::Engine — without the constants::kMaxNumThreads = 256 cap (cc/constants/constants.h:78), so one
// interface can drive a 1024-slot B200 engine.
//
// Same protocol and invariants as the reference (cc/nn/nn_interface.cc:286-371, nn_interface.h:251-312):
//   * a worker fills its slot with Engine::LoadBatch WITHOUT the lock, marks it loaded, and blocks until res_ready;
//   * the infer thread runs Engine::RunInference once every registered thread has loaded, or after `timeout_us` with
//     whatever is loaded; never while a result is still unread, never with nothing loaded; only the slots that were loaded
//     before the run are marked ready (a slot loaded during the run waits for the next cycle);
//   * Engine::GetBatch never overlaps RunInference; res_ready is cleared after GetBatch.
// What is NOT here: building GoFeatures from a game::Game (cc/game stays on the reference side of the boundary) and the
// per-thread NN cache (cc/core/lru_cache.h, keyed on zobrist hashes of game::Board).
#pragma once
#include <atomic>
#include <condition_variable>
#include <cstdint>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

#include "engine_iface.h"

namespace nn {

class B200Engine;

class NNInterfaceB200 {
 public:
  static constexpr int64_t kTimeoutUs = 400;  // nn_interface.h:205
  NNInterfaceB200(int num_threads, int64_t timeout_us, std::unique_ptr<Engine> engine);
  ~NNInterfaceB200();
  NNInterfaceB200(const NNInterfaceB200&) = delete;
  NNInterfaceB200& operator=(const NNInterfaceB200&) = delete;

  void RegisterThread(int thread_id);    // nn_interface.cc:198-209
  void UnregisterThread(int thread_id);  // nn_interface.cc:211-222

  // LoadBatch -> SignalLoadedAndBlockUntilReady -> GetBatch (NNInterface::LoadAndGetInference, nn_interface.cc:120-132,
  // minus the Game -> GoFeatures step and the cache).  `features` already carry the symmetry the caller chose.
  NNInferResult LoadAndGetInference(int thread_id, const GoFeatures& features);
  // Same with the symmetry applied / un-applied on the GPU (engine must be a B200Engine; p3_engine_load_batch_sym).
  NNInferResult LoadAndGetInferenceSym(int thread_id, const GoFeatures& features, int sym);

  uint64_t num_inferences() const { return num_inferences_.load(std::memory_order_relaxed); }
  Engine* engine() { return engine_.get(); }

 private:
  struct ThreadInfo {
    bool registered = true;
    bool loaded_for_inference = false;
    std::atomic<bool> res_ready{false};
  };
  void SignalLoadedAndBlockUntilReady(int thread_id);
  void InferLoop();
  void Infer();
  bool ShouldInfer() const;  // mu_ held

  const int num_threads_;
  const int64_t timeout_us_;
  std::unique_ptr<Engine> engine_;
  B200Engine* b200_ = nullptr;
  mutable std::mutex mu_;
  std::condition_variable infer_cv_;   // workers -> infer thread ("a slot was loaded / a thread left")
  std::condition_variable ready_cv_;   // infer thread -> workers ("results are ready")
  std::vector<ThreadInfo> thread_info_;
  int num_registered_threads_;
  std::atomic<bool> running_{true};
  std::atomic<uint64_t> num_inferences_{0};
  std::thread infer_thread_;
};

}  // namespace nn
