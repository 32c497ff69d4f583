// Build scaffolding for oracle/_ref ONLY.
#pragma once
#include <cstdint>
namespace absl {
struct Duration {
  int64_t ns = 0;
};
inline Duration Nanoseconds(int64_t v) { return Duration{v}; }
inline Duration Microseconds(int64_t v) { return Duration{v * 1000}; }
inline Duration Milliseconds(int64_t v) { return Duration{v * 1000000}; }
inline Duration Seconds(int64_t v) { return Duration{v * 1000000000}; }
}  // namespace absl
