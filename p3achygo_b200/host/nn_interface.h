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
// Double-buffered form: with `num_banks` = 2 over a B200Engine of batch num_threads / 2, the worker threads
// are split into two slot banks (thread t -> bank t / (num_threads / 2)), each with its own lock, infer thread and the
// protocol above; a bank's infer thread calls B200Engine::Submit + Wait instead of RunInference, so bank 1's workers load
// and read their slots (and its copies run) while bank 0's step is on the GPU.  The invariants hold per bank.
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
  NNInterfaceB200(int num_threads, int64_t timeout_us, std::unique_ptr<Engine> engine, int num_banks = 1);
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

  // NNInterface::LoadBatch from the game record itself (nn_interface.cc:245-277): the worker hands over Game::moves() (codes of
  // p3_game_derive), colour to move, komi, optional pass-alive grid and the symmetry; the engine derives every grid on the GPU.
  NNInferResult LoadAndGetInferenceGame(int thread_id, const int16_t* moves, int num_moves, int color, float komi,
                                        const int8_t* forbidden, int sym);

  uint64_t num_inferences() const { return num_inferences_.load(std::memory_order_relaxed); }
  Engine* engine() { return engine_.get(); }

 private:
  struct ThreadInfo {
    bool registered = true;
    bool loaded_for_inference = false;
    std::atomic<bool> res_ready{false};
  };
  // one slot bank: the reference NNInterface's state (nn_interface.h:329-345) for `size` consecutive thread ids
  struct Bank {
    explicit Bank(int index, int first, int size) : index(index), first(first), size(size), thread_info(size), num_registered(size) {}
    const int index, first, size;
    mutable std::mutex mu;
    std::condition_variable infer_cv;  // workers -> infer thread ("a slot was loaded / a thread left")
    std::condition_variable ready_cv;  // infer thread -> workers ("results are ready")
    std::vector<ThreadInfo> thread_info;
    int num_registered;
    std::thread infer_thread;
  };
  Bank& BankOf(int thread_id) { return *banks_[thread_id / slots_per_bank_]; }
  void SignalLoadedAndBlockUntilReady(int thread_id);
  void InferLoop(Bank* bank);
  void Infer(Bank& bank);
  bool ShouldInfer(const Bank& bank) const;  // bank.mu held
  void EngineLoad(int thread_id, const GoFeatures& features, int sym, bool with_sym);
  void EngineGet(int thread_id, NNInferResult& result);

  const int num_threads_;
  const int64_t timeout_us_;
  std::unique_ptr<Engine> engine_;
  B200Engine* b200_ = nullptr;
  const int num_banks_;
  const int slots_per_bank_;
  std::vector<std::unique_ptr<Bank>> banks_;
  std::atomic<bool> running_{true};
  std::atomic<uint64_t> num_inferences_{0};
};

}  // namespace nn
