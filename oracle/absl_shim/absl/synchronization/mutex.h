// Build scaffolding for oracle/_ref ONLY: absl::Mutex / Condition / MutexLock over std::mutex + std::condition_variable,
// with the semantics the reference's cc/nn/nn_interface.{h,cc} relies on:
//   * LockWhen(cond) / LockWhenWithTimeout(cond, d) return with the mutex HELD (the latter also after the timeout);
//   * a condition is re-evaluated whenever the mutex is released (abseil does that implicitly; here Unlock notifies).
#pragma once
#include <chrono>
#include <condition_variable>
#include <mutex>
#include "absl/time/time.h"
#define ABSL_LOCKS_EXCLUDED(...)
#define ABSL_GUARDED_BY(x)
#define ABSL_EXCLUSIVE_LOCKS_REQUIRED(...)
#define ABSL_SHARED_LOCKS_REQUIRED(...)
#define ABSL_NO_THREAD_SAFETY_ANALYSIS
namespace absl {
class Condition {
 public:
  template <typename T>
  Condition(bool (*fn)(T*), T* arg) : fn_([](const Condition* c) { return reinterpret_cast<bool (*)(T*)>(c->raw_fn_)(static_cast<T*>(c->arg_)); }),
                                      raw_fn_(reinterpret_cast<void (*)()>(fn)), arg_(const_cast<void*>(static_cast<const void*>(arg))) {}
  template <typename T>
  Condition(const T* obj, bool (T::*method)() const) : fn_([](const Condition* c) {
                                                        return (static_cast<const T*>(c->arg_)->*reinterpret_cast<const MethodHolder<T>*>(c->method_)->m)();
                                                      }),
                                                      arg_(const_cast<void*>(static_cast<const void*>(obj))) {
    static_assert(sizeof(MethodHolder<T>) <= sizeof(method_), "member pointer too large");
    new (method_) MethodHolder<T>{method};
  }
  explicit Condition(const bool* flag) : fn_([](const Condition* c) { return *static_cast<const bool*>(c->arg_); }),
                                         arg_(const_cast<void*>(static_cast<const void*>(flag))) {}
  bool Eval() const { return fn_(this); }

 private:
  template <typename T>
  struct MethodHolder {
    bool (T::*m)() const;
  };
  bool (*fn_)(const Condition*);
  void (*raw_fn_)() = nullptr;
  void* arg_ = nullptr;
  alignas(16) unsigned char method_[32] = {};
};

class Mutex {
 public:
  void Lock() { mu_.lock(); }
  void Unlock() {
    mu_.unlock();
    cv_.notify_all();
  }
  void LockWhen(const Condition& cond) {
    std::unique_lock<std::mutex> l(mu_);
    cv_.wait(l, [&] { return cond.Eval(); });
    l.release();
  }
  bool LockWhenWithTimeout(const Condition& cond, Duration timeout) {
    std::unique_lock<std::mutex> l(mu_);
    const bool ok = cv_.wait_for(l, std::chrono::nanoseconds(timeout.ns), [&] { return cond.Eval(); });
    l.release();
    return ok;
  }
  void Await(const Condition& cond) {  // mutex held on entry and on return
    std::unique_lock<std::mutex> l(mu_, std::adopt_lock);
    cv_.wait(l, [&] { return cond.Eval(); });
    l.release();
  }
  void AssertHeld() const {}
  void ReaderLock() { Lock(); }
  void ReaderUnlock() { Unlock(); }

 private:
  std::mutex mu_;
  std::condition_variable cv_;
};

class MutexLock {
 public:
  explicit MutexLock(Mutex* mu) : mu_(mu) { mu_->Lock(); }
  ~MutexLock() { mu_->Unlock(); }
  MutexLock(const MutexLock&) = delete;
  MutexLock& operator=(const MutexLock&) = delete;

 private:
  Mutex* mu_;
};
using ReaderMutexLock = MutexLock;
using WriterMutexLock = MutexLock;
}  // namespace absl
