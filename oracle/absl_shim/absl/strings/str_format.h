// Build scaffolding for oracle/_ref ONLY: absl::StrFormat over snprintf (printf-style specifiers only).
#pragma once
#include <cstdio>
#include <string>
namespace absl {
namespace shim {
template <typename T>
inline auto Arg(const T& v) { return v; }
inline const char* Arg(const std::string& v) { return v.c_str(); }
}  // namespace shim
template <typename... Args>
std::string StrFormat(const char* fmt, const Args&... args) {
  const int n = std::snprintf(nullptr, 0, fmt, shim::Arg(args)...);
  std::string s(n > 0 ? n : 0, '\0');
  if (n > 0) std::snprintf(s.data(), n + 1, fmt, shim::Arg(args)...);
  return s;
}
}  // namespace absl
