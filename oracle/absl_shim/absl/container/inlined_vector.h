// Build scaffolding for oracle/_ref ONLY. Must be a class template (not an
// alias) so the inline-capacity parameter N stays deducible in the reference's
// helper templates.
#pragma once
#include <cstddef>
#include <initializer_list>
#include <vector>
namespace absl {
template <typename T, size_t N>
class InlinedVector : public std::vector<T> {
 public:
  using std::vector<T>::vector;
  InlinedVector() { this->reserve(N < 64 ? N : 64); }
  InlinedVector(std::initializer_list<T> il) : std::vector<T>(il) {}
};
}  // namespace absl
