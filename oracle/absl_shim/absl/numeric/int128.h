// Build scaffolding for oracle/_ref ONLY (not product code): a minimal stand-in
// for the abseil headers the reference's cc/game + cc/core sources include, so
// those sources compile unmodified from /root/reference with plain g++.
#pragma once
#include <cstdint>
#include <functional>
#include <ostream>
namespace absl {
using uint128 = unsigned __int128;
inline constexpr uint128 MakeUint128(uint64_t hi, uint64_t lo) {
  return (static_cast<uint128>(hi) << 64) | lo;
}
inline constexpr uint64_t Uint128High64(uint128 v) { return static_cast<uint64_t>(v >> 64); }
inline constexpr uint64_t Uint128Low64(uint128 v) { return static_cast<uint64_t>(v); }
}  // namespace absl
inline std::ostream& operator<<(std::ostream& os, unsigned __int128 v) {
  return os << std::hex << static_cast<uint64_t>(v >> 64) << static_cast<uint64_t>(v) << std::dec;
}
