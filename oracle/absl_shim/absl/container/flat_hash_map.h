// Build scaffolding for oracle/_ref ONLY.
#pragma once
#include <unordered_map>
#include "absl/hash/hash.h"
namespace absl {
template <typename K, typename V, typename H = absl::Hash<K>>
using flat_hash_map = std::unordered_map<K, V, H>;
template <typename K, typename V, typename H, typename E, typename A, typename Pred>
size_t erase_if(std::unordered_map<K, V, H, E, A>& c, Pred pred) {
  return std::erase_if(c, pred);
}
}  // namespace absl
