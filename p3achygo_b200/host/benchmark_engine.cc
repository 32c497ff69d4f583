// Engine micro-benchmark with the shape of the reference's nn::Benchmark
// (cc/nn/engine/benchmark_engine.cc:77-109): warm-up RunInference calls, then per batch
//     LoadBatch x B  ->  RunInference (timed with steady_clock, as the reference does)  ->  GetBatch x B
// over synthetic positions instead of the TFRecord dataset.  Also times the whole cycle, with the
// LoadBatch / GetBatch calls spread over `threads` host threads the way NNInterface's workers issue them
// (cc/nn/nn_interface.cc:245-277, cc/nn/nn_interface.h:251-290).
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <cstdio>
#include <thread>
#include <vector>

#include "b200_engine.h"
#include "go_dataset.h"

namespace {
using Clock = std::chrono::steady_clock;
double us_since(Clock::time_point t0) { return std::chrono::duration<double, std::micro>(Clock::now() - t0).count(); }

// Persistent worker threads, as the search threads that call NNInterface::LoadBatch / GetBatch are
// (cc/nn/nn_interface.cc:245-277): run(n, fn) hands out indices i = t, t + threads, ... and returns when all are done.
class WorkerPool {
 public:
  explicit WorkerPool(int threads) : threads_(threads < 1 ? 1 : threads) {
    for (int t = 1; t < threads_; ++t) pool_.emplace_back([this, t]() { Loop(t); });
  }
  ~WorkerPool() {
    {
      std::lock_guard<std::mutex> lk(mu_);
      stop_ = true;
      ++generation_;
    }
    cv_.notify_all();
    for (auto& th : pool_) th.join();
  }
  void run(int n, const std::function<void(int)>& fn) {
    {
      std::lock_guard<std::mutex> lk(mu_);
      fn_ = &fn;
      n_ = n;
      pending_ = threads_ - 1;
      ++generation_;
    }
    cv_.notify_all();
    for (int i = 0; i < n; i += threads_) fn(i);  // the caller is worker 0
    std::unique_lock<std::mutex> lk(mu_);
    done_cv_.wait(lk, [this]() { return pending_ == 0; });
  }

 private:
  void Loop(int t) {
    unsigned long long seen = 0;
    while (true) {
      const std::function<void(int)>* fn;
      int n;
      {
        std::unique_lock<std::mutex> lk(mu_);
        cv_.wait(lk, [&]() { return generation_ != seen; });
        seen = generation_;
        if (stop_) return;
        fn = fn_;
        n = n_;
      }
      for (int i = t; i < n; i += threads_) (*fn)(i);
      {
        std::lock_guard<std::mutex> lk(mu_);
        --pending_;
      }
      done_cv_.notify_one();
    }
  }
  const int threads_;
  std::vector<std::thread> pool_;
  std::mutex mu_;
  std::condition_variable cv_, done_cv_;
  const std::function<void(int)>* fn_ = nullptr;
  int n_ = 0, pending_ = 0;
  unsigned long long generation_ = 0;
  bool stop_ = false;
};
}  // namespace

extern "C" {

// positions: n_positions GoFeatures records (HOST). Cycles through them `steps` times (after `warmup` untimed
// cycles), batch slots filled from consecutive records.  Outputs (microseconds, per cycle averages):
//   out[0] RunInference only   out[1] LoadBatch x B   out[2] GetBatch x B   out[3] whole cycle
//   out[4] checksum of value_probs (keeps GetBatch honest)
int p3_host_benchmark(const char* weights_path, int device, int batch, int version, int precision,
                      const p3_go_features* positions, int n_positions, int warmup, int steps, int threads, double* out) {
  auto engine = nn::B200Engine::Create(weights_path, batch, version, device, precision);
  std::vector<nn::NNInferResult> results(batch);
  WorkerPool pool(threads);
  double t_run = 0, t_load = 0, t_get = 0, t_cycle = 0, checksum = 0;
  int cursor = 0;
  for (int it = 0; it < warmup + steps; ++it) {
    const bool timed = it >= warmup;
    const int base = cursor;
    cursor = (cursor + batch) % n_positions;
    auto c0 = Clock::now();
    pool.run(batch, [&](int b) { engine->LoadBatch(b, positions[(base + b) % n_positions]); });
    const double load_us = us_since(c0);
    auto r0 = Clock::now();
    engine->RunInference();
    const double run_us = us_since(r0);
    auto g0 = Clock::now();
    pool.run(batch, [&](int b) { engine->GetBatch(b, results[b]); });
    const double get_us = us_since(g0);
    const double cycle_us = us_since(c0);
    if (timed) {
      t_run += run_us;
      t_load += load_us;
      t_get += get_us;
      t_cycle += cycle_us;
      for (int b = 0; b < batch; ++b) checksum += results[b].value_probs[1];
    }
  }
  out[0] = t_run / steps;
  out[1] = t_load / steps;
  out[2] = t_get / steps;
  out[3] = t_cycle / steps;
  out[4] = checksum;
  return 0;
}

// The same cycle over the engine's two slot banks: while bank k's step is on the GPU, the workers read the previous
// results of bank 1-k and load its next positions, and both banks' copies run on the copy streams.  Every position is
// still loaded from host memory, evaluated, and read back into a caller-owned NNInferResult inside the timed region.
//   out[0] whole cycle per batch (us, total wall time / steps)   out[1] LoadBatch x B   out[2] GetBatch x B
//   out[3] Wait (time the host blocked on the GPU)               out[4] checksum of value_probs
// games != nullptr: the slots are loaded as game records instead (LoadGameBank: n_positions move lists of max_moves int16 codes,
// num_moves, colours; komi 7.5), so the board, the liberty grids and the laddered stones are derived on the GPU inside the step.
static int pipelined_cycle(const char* weights_path, int device, int batch, int version, int precision,
                           const p3_go_features* positions, const int16_t* games, const int32_t* num_moves, const int8_t* colors,
                           int max_moves, int n_positions, int warmup, int steps, int threads, double* out, bool leaf = false);

// The pipelined cycle with compact leaf records (P3_RESULT_LEAF): GetLeafBank x B reads 4360 B per slot - what
// mcts::LeafEvaluator's InitFields keeps of an NNInferResult - and only those cross PCIe.
int p3_host_benchmark_pipelined_leaf(const char* weights_path, int device, int batch, int version, int precision,
                                     const p3_go_features* positions, int n_positions, int warmup, int steps, int threads, double* out) {
  return pipelined_cycle(weights_path, device, batch, version, precision, positions, nullptr, nullptr, nullptr, 0, n_positions, warmup,
                         steps, threads, out, true);
}

int p3_host_benchmark_pipelined(const char* weights_path, int device, int batch, int version, int precision,
                                const p3_go_features* positions, int n_positions, int warmup, int steps, int threads, double* out) {
  return pipelined_cycle(weights_path, device, batch, version, precision, positions, nullptr, nullptr, nullptr, 0, n_positions, warmup,
                         steps, threads, out);
}

int p3_host_benchmark_games(const char* weights_path, int device, int batch, int version, int precision, const int16_t* games,
                            const int32_t* num_moves, const int8_t* colors, int max_moves, int n_games, int warmup, int steps,
                            int threads, double* out) {
  return pipelined_cycle(weights_path, device, batch, version, precision, nullptr, games, num_moves, colors, max_moves, n_games, warmup,
                         steps, threads, out);
}

static int pipelined_cycle(const char* weights_path, int device, int batch, int version, int precision,
                           const p3_go_features* positions, const int16_t* games, const int32_t* num_moves, const int8_t* colors,
                           int max_moves, int n_positions, int warmup, int steps, int threads, double* out, bool leaf) {
  auto engine = nn::B200Engine::Create(weights_path, batch, version, device, precision);
  engine->SetLeafResults(leaf);
  std::vector<nn::NNInferResult> results(leaf ? 0 : batch);
  std::vector<p3_leaf_result> leaves(leaf ? batch : 0);
  WorkerPool pool(threads);
  double t_load = 0, t_get = 0, t_wait = 0, checksum = 0;
  int cursor = 0;
  const int total = warmup + steps;
  auto load = [&](int bank) {
    const int base = cursor;
    cursor = (cursor + batch) % n_positions;
    if (games)
      pool.run(batch, [&](int b) {
        const int g = (base + b) % n_positions;
        engine->LoadGameBank(bank, b, games + static_cast<size_t>(g) * max_moves, num_moves[g], colors[g], 7.5f, nullptr, g % 8);
      });
    else
      pool.run(batch, [&](int b) { engine->LoadBatchBank(bank, b, positions[(base + b) % n_positions]); });
  };
  Clock::time_point t_start = Clock::now();
  load(0);
  engine->Submit(0);
  for (int it = 0; it < total; ++it) {
    if (it == warmup) {  // drain, then start the clock with bank `it` loaded and submitted inside the timed region
      engine->Wait(it & 1);
      t_start = Clock::now();
      load(it & 1);
      engine->Submit(it & 1);
      t_load = t_get = t_wait = 0;
    }
    const int cur = it & 1, nxt = cur ^ 1;
    if (it + 1 < total) {
      auto l0 = Clock::now();
      load(nxt);
      t_load += us_since(l0);
      engine->Submit(nxt);
    }
    auto w0 = Clock::now();
    engine->Wait(cur);
    t_wait += us_since(w0);
    auto g0 = Clock::now();
    if (leaf) pool.run(batch, [&](int b) { engine->GetLeafBank(cur, b, leaves[b]); });
    else pool.run(batch, [&](int b) { engine->GetBatchBank(cur, b, results[b]); });
    t_get += us_since(g0);
    if (it >= warmup)
      for (int b = 0; b < batch; ++b) checksum += leaf ? leaves[b].value : results[b].value_probs[1];
  }
  const double total_us = us_since(t_start);
  out[0] = total_us / steps;
  out[1] = t_load / steps;
  out[2] = t_get / steps;
  out[3] = t_wait / steps;
  out[4] = checksum;
  return 0;
}

// nn::Benchmark(engine, go_ds, stats) with nn::DefaultStats (cc/nn/engine/benchmark_engine.cc:24-109): warm-up runs, then per
// dataset batch LoadBatch x B -> RunInference (timed) -> GetBatch x B, accumulating the reference's accuracy statistics.
//   out[0] num_examples  out[1] avg_us  out[2] policy_loss  out[3] outcome_loss  out[4] policy_percent  out[5] outcome_percent
//   out[6] score_diff
int p3_host_benchmark_dataset(const char* weights_path, int device, int batch, int version, int precision, const char* ds_path,
                              int warmup, double* out) {
  auto engine = nn::B200Engine::Create(weights_path, batch, version, device, precision);
  nn::GoDataset ds(static_cast<size_t>(batch), ds_path);
  for (int i = 0; i < warmup; ++i) engine->RunInference();  // "Warming Up...", :79-82
  auto argmax = [](const float* v, int n) {
    int best = 0;
    for (int i = 1; i < n; ++i)
      if (v[i] > v[best]) best = i;
    return best;
  };
  auto avg = [](double a, double x, double cnt) { return ((cnt - 1) / cnt) * a + (1 / cnt) * x; };
  double cnt = 0, avg_us = 0, policy_loss = 0, outcome_loss = 0, policy_percent = 0, outcome_percent = 0, score_diff = 0;
  constexpr double kMaxLoss = 16;
  int num_inferences = 0;
  nn::NNInferResult result;
  size_t remaining = ds.num_examples();
  for (auto& rows : ds) {
    if (num_inferences > 1000) break;  // :88-90
    for (size_t b = 0; b < rows.size(); ++b) engine->LoadBatch(static_cast<int>(b), rows[b].features);
    auto t0 = Clock::now();
    engine->RunInference();
    const double elapsed_us = static_cast<double>(static_cast<long long>(us_since(t0)));  // duration_cast<microseconds>, :96-99
    for (size_t b = 0; b < rows.size() && remaining > 0; ++b, --remaining) {  // (the reference also scores the padding rows of the last batch)
      engine->GetBatch(static_cast<int>(b), result);
      const nn::GoDataset::Row& row = rows[b];
      const int mv_pred = argmax(result.move_probs, P3_MAX_MOVES);
      const int outcome_pred = argmax(result.value_probs, P3_NUM_VALUE_LOGITS);
      const int score_pred = static_cast<int>(argmax(result.score_probs, P3_NUM_SCORE_LOGITS) + 0.5 - 400);  // kScoreInflectionPoint
      const int mv = argmax(row.labels.policy.data(), P3_MAX_MOVES);
      const int won = row.labels.did_win ? 1 : 0;
      const float pi_ce = result.move_probs[mv] != 0.0f ? -std::log(result.move_probs[mv]) : static_cast<float>(kMaxLoss);
      const float v_ce = result.value_probs[won] != 0.0f ? -std::log(result.value_probs[won]) : static_cast<float>(kMaxLoss);
      cnt += 1;
      avg_us = avg(avg_us, elapsed_us, cnt);
      policy_loss = avg(policy_loss, pi_ce, cnt);
      outcome_loss = avg(outcome_loss, v_ce, cnt);
      policy_percent = avg(policy_percent, mv == mv_pred ? 1 : 0, cnt);
      outcome_percent = avg(outcome_percent, won == outcome_pred ? 1 : 0, cnt);
      score_diff = avg(score_diff, std::abs(row.labels.score_margin - score_pred), cnt);
    }
    ++num_inferences;
  }
  out[0] = cnt, out[1] = avg_us, out[2] = policy_loss, out[3] = outcome_loss, out[4] = policy_percent, out[5] = outcome_percent;
  out[6] = score_diff;
  return 0;
}

}  // extern "C"
