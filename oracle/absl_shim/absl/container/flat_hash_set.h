// Build scaffolding for oracle/_ref ONLY (iteration order differs from abseil;
// the reference results on this path are order-independent).
#pragma once
#include <unordered_set>
#include "absl/hash/hash.h"
namespace absl {
template <typename T, typename H = absl::Hash<T>>
using flat_hash_set = std::unordered_set<T, H>;
template <typename T, typename H, typename E, typename A, typename Pred>
size_t erase_if(std::unordered_set<T, H, E, A>& c, Pred pred) {
  return std::erase_if(c, pred);
}
}  // namespace absl
