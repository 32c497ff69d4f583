#include "nn_interface.h"

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <random>

#include "b200_engine.h"

namespace nn {

NNInterfaceB200::NNInterfaceB200(int num_threads, int64_t timeout_us, std::unique_ptr<Engine> engine, int num_banks)
    : num_threads_(num_threads), timeout_us_(timeout_us), engine_(std::move(engine)), num_banks_(num_banks),
      slots_per_bank_(num_threads / (num_banks < 1 ? 1 : num_banks)) {
  b200_ = dynamic_cast<B200Engine*>(engine_.get());
  if (num_banks_ < 1 || num_banks_ > B200Engine::kNumBanks || num_threads_ % num_banks_ != 0 || (num_banks_ > 1 && !b200_) ||
      (num_banks_ > 1 && b200_->batch_size() != slots_per_bank_)) {
    std::fprintf(stderr, "NNInterfaceB200: %d banks need a B200Engine of batch num_threads / banks\n", num_banks);
    std::abort();  // LOG(FATAL) in the reference's idiom
  }
  for (int k = 0; k < num_banks_; ++k) banks_.emplace_back(new Bank(k, k * slots_per_bank_, slots_per_bank_));
  if (num_threads_ > 1)
    for (auto& b : banks_) b->infer_thread = std::thread(&NNInterfaceB200::InferLoop, this, b.get());
}

NNInterfaceB200::~NNInterfaceB200() {  // nn_interface.cc:82-90
  for (auto& b : banks_) {
    std::lock_guard<std::mutex> l(b->mu);
    running_.store(false, std::memory_order_release);
  }
  for (auto& b : banks_) b->infer_cv.notify_all();
  for (auto& b : banks_)
    if (b->infer_thread.joinable()) b->infer_thread.join();
}

void NNInterfaceB200::RegisterThread(int thread_id) {
  Bank& bank = BankOf(thread_id);
  std::lock_guard<std::mutex> l(bank.mu);
  ThreadInfo& t = bank.thread_info[thread_id - bank.first];
  if (t.registered) return;
  t.registered = true;
  t.loaded_for_inference = false;
  t.res_ready.store(false, std::memory_order_relaxed);
  ++bank.num_registered;
}

void NNInterfaceB200::UnregisterThread(int thread_id) {
  Bank& bank = BankOf(thread_id);
  {
    std::lock_guard<std::mutex> l(bank.mu);
    ThreadInfo& t = bank.thread_info[thread_id - bank.first];
    if (!t.registered) return;
    t.registered = false;
    t.loaded_for_inference = false;
    --bank.num_registered;
  }
  bank.infer_cv.notify_all();  // the remaining threads may now satisfy ShouldInfer
}

void NNInterfaceB200::SignalLoadedAndBlockUntilReady(int thread_id) {  // nn_interface.h:293-309
  if (num_threads_ == 1) {
    engine_->RunInference();
    num_inferences_.fetch_add(1, std::memory_order_relaxed);
    return;
  }
  Bank& bank = BankOf(thread_id);
  ThreadInfo& t = bank.thread_info[thread_id - bank.first];
  {
    std::lock_guard<std::mutex> l(bank.mu);
    t.loaded_for_inference = true;
    t.res_ready.store(false, std::memory_order_relaxed);
  }
  bank.infer_cv.notify_all();
  std::unique_lock<std::mutex> l(bank.mu);
  bank.ready_cv.wait(l, [&]() { return t.res_ready.load(std::memory_order_acquire); });
}

void NNInterfaceB200::EngineLoad(int thread_id, const GoFeatures& features, int sym, bool with_sym) {
  if (num_banks_ > 1) {
    b200_->LoadBatchBank(thread_id / slots_per_bank_, thread_id % slots_per_bank_, features, with_sym ? sym : 0);
  } else if (with_sym) {
    b200_->LoadBatchSym(thread_id, features, sym);
  } else {
    engine_->LoadBatch(thread_id, features);
  }
}

void NNInterfaceB200::EngineGet(int thread_id, NNInferResult& result) {
  if (num_banks_ > 1) b200_->GetBatchBank(thread_id / slots_per_bank_, thread_id % slots_per_bank_, result);
  else engine_->GetBatch(thread_id, result);
}

NNInferResult NNInterfaceB200::LoadAndGetInference(int thread_id, const GoFeatures& features) {
  EngineLoad(thread_id, features, 0, false);  // no lock held (nn_interface.cc:276)
  SignalLoadedAndBlockUntilReady(thread_id);
  NNInferResult r;
  EngineGet(thread_id, r);
  Bank& bank = BankOf(thread_id);
  bank.thread_info[thread_id - bank.first].res_ready.store(false, std::memory_order_release);  // nn_interface.h:256-261
  return r;
}

NNInferResult NNInterfaceB200::LoadAndGetInferenceSym(int thread_id, const GoFeatures& features, int sym) {
  if (!b200_) {
    std::fprintf(stderr, "NNInterfaceB200::LoadAndGetInferenceSym needs a B200Engine\n");
    std::abort();
  }
  EngineLoad(thread_id, features, sym, true);
  SignalLoadedAndBlockUntilReady(thread_id);
  NNInferResult r;
  EngineGet(thread_id, r);
  Bank& bank = BankOf(thread_id);
  bank.thread_info[thread_id - bank.first].res_ready.store(false, std::memory_order_release);
  return r;
}

NNInferResult NNInterfaceB200::LoadAndGetInferenceGame(int thread_id, const int16_t* moves, int num_moves, int color, float komi,
                                                       const int8_t* forbidden, int sym) {
  if (!b200_) {
    std::fprintf(stderr, "NNInterfaceB200::LoadAndGetInferenceGame needs a B200Engine\n");
    std::abort();
  }
  b200_->LoadGameBank(thread_id / slots_per_bank_, thread_id % slots_per_bank_, moves, num_moves, color, komi, forbidden, sym);
  SignalLoadedAndBlockUntilReady(thread_id);
  NNInferResult r;
  EngineGet(thread_id, r);
  Bank& bank = BankOf(thread_id);
  bank.thread_info[thread_id - bank.first].res_ready.store(false, std::memory_order_release);
  return r;
}

void NNInterfaceB200::InferLoop(Bank* bank) {
  while (running_.load(std::memory_order_acquire)) Infer(*bank);
}

bool NNInterfaceB200::ShouldInfer(const Bank& bank) const {  // kAuto branch of nn_interface.cc:373-400
  if (!running_.load(std::memory_order_acquire)) return true;
  bool exists_pending = false;
  for (const ThreadInfo& t : bank.thread_info) {
    if (!t.registered) continue;
    if (!t.loaded_for_inference) return false;
    exists_pending = true;
  }
  return exists_pending;
}

void NNInterfaceB200::Infer(Bank& bank) {  // nn_interface.cc:286-371
  std::unique_lock<std::mutex> l(bank.mu);
  if (timeout_us_ > 0)
    bank.infer_cv.wait_for(l, std::chrono::microseconds(timeout_us_), [&]() { return ShouldInfer(bank); });
  else
    bank.infer_cv.wait(l, [&]() { return ShouldInfer(bank); });
  if (!running_.load(std::memory_order_acquire) || bank.num_registered == 0) return;
  // never overwrite an unread result
  for (const ThreadInfo& t : bank.thread_info)
    if (t.res_ready.load(std::memory_order_acquire)) return;
  bool any_loaded = false;
  for (const ThreadInfo& t : bank.thread_info) any_loaded = any_loaded || t.loaded_for_inference;
  if (!any_loaded) return;
  // Only the slots that are loaded NOW get this run's results: a slot whose worker is still inside LoadBatch (or loads
  // while the engine runs) is marked loaded later and waits for the next cycle (nn_interface.cc:355-369).  The reference
  // holds mu_ across RunInference; workers only need mu_ to mark themselves loaded, so releasing it here is equivalent for
  // them and lets them queue up for the next cycle while the GPU works.
  std::vector<char> in_batch(bank.size, 0);
  for (int i = 0; i < bank.size; ++i) in_batch[i] = bank.thread_info[i].registered && bank.thread_info[i].loaded_for_inference;
  l.unlock();
  if (num_banks_ > 1) {
    b200_->Submit(bank.index);  // the other bank's step may be on the GPU: this one queues behind it, copies overlap
    b200_->Wait(bank.index);
  } else {
    engine_->RunInference();
  }
  num_inferences_.fetch_add(1, std::memory_order_relaxed);
  l.lock();
  for (int i = 0; i < bank.size; ++i) {
    if (!in_batch[i]) continue;
    ThreadInfo& t = bank.thread_info[i];
    t.loaded_for_inference = false;
    t.res_ready.store(true, std::memory_order_release);
  }
  l.unlock();
  bank.ready_cv.notify_all();
}

}  // namespace nn

// =====================================================================================================================
// C entry points for the tests (ctypes)
// =====================================================================================================================
namespace {

// Port of cc/nn/__tests__/nn_interface_sync_test.cc: an engine whose RunInference writes a per-slot function to many
// elements (yielding between slots) and whose GetBatch checks (1) no torn read, (2) result generation > load generation,
// (3) own slot's value.
class CountingEngine : public nn::Engine {
 public:
  static constexpr int kSlotElems = 32;
  static constexpr int kPrime = (1 << 19) - 1;
  static int SlotFn(int tid) { return (tid + tid) % kPrime; }
  explicit CountingEngine(int n) : n_(n), buffer_(n * kSlotElems), result_gen_(n), load_gen_(n) {}
  Kind kind() override { return Kind::kUnknown; }
  std::string path() override { return ""; }
  void GetOwnership(int, std::array<float, P3_NUM_BOARD_LOCS>&) override {}
  void LoadBatch(int t, const nn::GoFeatures&) override {
    load_gen_[t].store(generation_.load(std::memory_order_acquire), std::memory_order_release);
  }
  void RunInference() override {
    const int gen = generation_.fetch_add(1, std::memory_order_relaxed) + 1;
    for (int t = 0; t < n_; ++t) {
      const int f = SlotFn(t);
      for (int i = 0; i < kSlotElems; ++i) buffer_[t * kSlotElems + i].store(f, std::memory_order_relaxed);
      result_gen_[t].store(gen, std::memory_order_relaxed);
      std::this_thread::yield();
    }
  }
  void GetBatch(int id, nn::NNInferResult& r) override {
    int vals[kSlotElems];
    for (int i = 0; i < kSlotElems; ++i) {
      vals[i] = buffer_[id * kSlotElems + i].load(std::memory_order_relaxed);
      std::this_thread::yield();
    }
    for (int i = 1; i < kSlotElems; ++i)
      if (vals[i] != vals[0]) race.store(true);
    if (result_gen_[id].load(std::memory_order_relaxed) <= load_gen_[id].load(std::memory_order_acquire)) stale.store(true);
    if (vals[0] != SlotFn(id)) wrong_slot.store(true);
    r.move_logits[0] = static_cast<float>(vals[0]);
  }
  std::atomic<bool> race{false}, stale{false}, wrong_slot{false};

 private:
  const int n_;
  std::atomic<int> generation_{0};
  std::vector<std::atomic<int>> buffer_, result_gen_, load_gen_;
};

}  // namespace

extern "C" {

// The reference's sync stress test on NNInterfaceB200: `threads` workers (every 8th one slow enough to force partial,
// timed-out batches) hammer LoadAndGetInference for `millis` ms.  out[0..3] = race, stale, wrong-slot, wrong-value counts
// (all must be 0), out[4] = inference cycles, out[5] = evaluations served.
int p3_host_iface_sync_test(int threads, int millis, int timeout_us, long long* out) {
  auto* engine = new CountingEngine(threads);
  std::atomic<long long> served{0}, wrong_value{0};
  {
    nn::NNInterfaceB200 iface(threads, timeout_us, std::unique_ptr<nn::Engine>(engine));
    std::atomic<bool> stop{false};
    std::vector<std::thread> pool;
    for (int tid = 0; tid < threads; ++tid)
      pool.emplace_back([&, tid]() {
        std::mt19937 rng(static_cast<uint32_t>(tid) * 2654435761u);
        std::uniform_int_distribution<int> jitter_us(100, 1000), slow_ms(5, 50);
        const bool is_slow = tid % 8 == 0;
        nn::GoFeatures f{};
        while (!stop.load(std::memory_order_relaxed)) {
          if (is_slow) std::this_thread::sleep_for(std::chrono::milliseconds(slow_ms(rng)));
          else std::this_thread::sleep_for(std::chrono::microseconds(jitter_us(rng)));
          const nn::NNInferResult r = iface.LoadAndGetInference(tid, f);
          if (static_cast<int>(r.move_logits[0]) != CountingEngine::SlotFn(tid)) wrong_value.fetch_add(1);
          served.fetch_add(1, std::memory_order_relaxed);
        }
        iface.UnregisterThread(tid);
      });
    std::this_thread::sleep_for(std::chrono::milliseconds(millis));
    stop.store(true);
    for (auto& t : pool) t.join();
    out[4] = static_cast<long long>(iface.num_inferences());
    out[0] = engine->race.load();
    out[1] = engine->stale.load();
    out[2] = engine->wrong_slot.load();
  }
  out[3] = wrong_value.load();
  out[5] = served.load();
  return 0;
}

// `threads` workers each evaluate `per_thread` positions through NNInterfaceB200 over a real B200 engine with `threads`
// slots (1024 is fine: no kMaxNumThreads cap); thread t evaluates positions t, t + threads, ... and stores NNInferResult
// records at results[position].  use_sym != 0: LoadAndGetInferenceSym with symmetry (position % 8).
// `banks` = 2: the double-buffered interface over an engine of batch threads / 2.
int p3_host_iface_run_banks(const char* weights_path, int device, int threads, int version, int precision,
                            const p3_go_features* positions, int n_positions, int use_sym, int timeout_us, int banks,
                            p3_infer_result* results, long long* n_inferences, double* seconds) {
  auto engine = nn::B200Engine::Create(weights_path, threads / banks, version, device, precision);
  nn::NNInterfaceB200 iface(threads, timeout_us, std::move(engine), banks);
  const auto t0 = std::chrono::steady_clock::now();
  std::vector<std::thread> pool;
  for (int tid = 0; tid < threads; ++tid)
    pool.emplace_back([&, tid]() {
      for (int p = tid; p < n_positions; p += threads)
        results[p] = use_sym ? iface.LoadAndGetInferenceSym(tid, positions[p], p % 8) : iface.LoadAndGetInference(tid, positions[p]);
      iface.UnregisterThread(tid);
    });
  for (auto& t : pool) t.join();
  if (seconds) *seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  if (n_inferences) *n_inferences = static_cast<long long>(iface.num_inferences());
  return 0;
}

// As p3_host_iface_run_banks, with every position given as a game record (n_positions move lists of max_moves codes).
int p3_host_iface_run_games(const char* weights_path, int device, int threads, int version, int precision, const int16_t* games,
                            const int32_t* num_moves, const int8_t* colors, const int8_t* forbidden, int max_moves, int n_positions,
                            int timeout_us, int banks, p3_infer_result* results, long long* n_inferences) {
  auto engine = nn::B200Engine::Create(weights_path, threads / banks, version, device, precision);
  nn::NNInterfaceB200 iface(threads, timeout_us, std::move(engine), banks);
  std::vector<std::thread> pool;
  for (int tid = 0; tid < threads; ++tid)
    pool.emplace_back([&, tid]() {
      for (int p = tid; p < n_positions; p += threads)
        results[p] = iface.LoadAndGetInferenceGame(tid, games + static_cast<size_t>(p) * max_moves, num_moves[p], colors[p], 7.5f,
                                                   forbidden ? forbidden + static_cast<size_t>(p) * 361 : nullptr, p % 8);
      iface.UnregisterThread(tid);
    });
  for (auto& t : pool) t.join();
  if (n_inferences) *n_inferences = static_cast<long long>(iface.num_inferences());
  return 0;
}

int p3_host_iface_run(const char* weights_path, int device, int threads, int version, int precision, const p3_go_features* positions,
                      int n_positions, int use_sym, int timeout_us, p3_infer_result* results, long long* n_inferences) {
  return p3_host_iface_run_banks(weights_path, device, threads, version, precision, positions, n_positions, use_sym, timeout_us, 1,
                                 results, n_inferences, nullptr);
}

}  // extern "C"
