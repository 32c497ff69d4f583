// Build scaffolding for oracle/_ref ONLY (iteration order differs from abseil;
// the reference results on this path are order-independent).
#pragma once
#include <unordered_set>
#include "absl/hash/hash.h"
namespace absl {
template <typename T, typename H = absl::Hash<T>>
using flat_hash_set = std::unordered_set<T, H>;
}
