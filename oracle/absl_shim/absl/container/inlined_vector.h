// Build scaffolding for oracle/_ref ONLY. Must be a class template (not an
// alias) so the inline-capacity parameter N stays deducible in the reference's
// helper templates.
//
// Semantics the reference relies on and a bare std::vector does not give: while size() <= N an absl::InlinedVector never
// reallocates, so references to its elements survive push_back (cc/mcts/gumbel.cc:691-706 holds `auto& [action, node] =
// path.back()` across `path.push_back(...)`).  Every constructor / assignment therefore reserves N up front.
#pragma once
#include <cstddef>
#include <initializer_list>
#include <type_traits>
#include <utility>
#include <vector>
namespace absl {
template <typename T, size_t N>
class InlinedVector : public std::vector<T> {
  using Base = std::vector<T>;
  void Reserve(size_t n) {  // (a type that can be neither copied nor moved can never grow anyway)
    if constexpr (std::is_move_constructible_v<T> || std::is_copy_constructible_v<T>) this->reserve(n);
  }

 public:
  InlinedVector() { Reserve(N); }
  explicit InlinedVector(size_t n) : Base(n) { Reserve(n > N ? n : N); }
  InlinedVector(size_t n, const T& v) {
    Reserve(n > N ? n : N);
    this->assign(n, v);
  }
  InlinedVector(std::initializer_list<T> il) {
    Reserve(il.size() > N ? il.size() : N);
    this->insert(this->end(), il.begin(), il.end());
  }
  template <typename It, typename = decltype(*std::declval<It>())>
  InlinedVector(It first, It last) {
    Reserve(N);
    this->insert(this->end(), first, last);
  }
  InlinedVector(const InlinedVector& o) {
    Reserve(o.size() > N ? o.size() : N);
    this->insert(this->end(), o.begin(), o.end());
  }
  InlinedVector(InlinedVector&& o) noexcept : Base(std::move(static_cast<Base&>(o))) {
    if (this->capacity() < N) Reserve(N);
  }
  InlinedVector& operator=(const InlinedVector& o) {
    if (this != &o) {
      this->clear();
      Reserve(o.size() > N ? o.size() : N);
      this->insert(this->end(), o.begin(), o.end());
    }
    return *this;
  }
  InlinedVector& operator=(InlinedVector&& o) noexcept {
    Base::operator=(std::move(static_cast<Base&>(o)));
    if (this->capacity() < N) Reserve(N);
    return *this;
  }
};
}  // namespace absl
