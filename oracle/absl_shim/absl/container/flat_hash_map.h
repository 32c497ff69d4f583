// Build scaffolding for oracle/_ref ONLY.
#pragma once
#include <unordered_map>
#include "absl/hash/hash.h"
namespace absl {
template <typename K, typename V, typename H = absl::Hash<K>>
using flat_hash_map = std::unordered_map<K, V, H>;
}
