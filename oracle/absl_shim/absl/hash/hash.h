// Build scaffolding for oracle/_ref ONLY: absl::Hash / HashOf stand-in that
// dispatches to a type's AbslHashValue friend when it has one.
#pragma once
#include <cstddef>
#include <cstdint>
#include <array>
#include <functional>
#include <tuple>
#include <type_traits>
#include <utility>
namespace absl {
namespace shim {
struct HashState {
  size_t h = 0xcbf29ce484222325ull;
  static HashState combine(HashState s) { return s; }
  template <typename T, typename... Ts>
  static HashState combine(HashState s, const T& v, const Ts&... rest);
  template <typename T>
  static HashState combine_contiguous(HashState s, const T* p, size_t n) {
    for (size_t i = 0; i < n; ++i) s = combine(std::move(s), p[i]);
    return s;
  }
};
template <typename T, typename = void>
struct HasAbslHashValue : std::false_type {};
template <typename T>
struct HasAbslHashValue<T, std::void_t<decltype(AbslHashValue(std::declval<HashState>(), std::declval<const T&>()))>>
    : std::true_type {};
inline void Mix(HashState& s, size_t v) {
  s.h ^= v + 0x9e3779b97f4a7c15ull + (s.h << 6) + (s.h >> 2);
}
template <typename T> struct IsTuple : std::false_type {};
template <typename... Ts> struct IsTuple<std::tuple<Ts...>> : std::true_type {};
template <typename A, typename B> struct IsTuple<std::pair<A, B>> : std::true_type {};
template <typename T> struct IsStdArray : std::false_type {};
template <typename T, size_t N> struct IsStdArray<std::array<T, N>> : std::true_type {};
template <typename T>
HashState HashOne(HashState s, const T& v) {
  if constexpr (HasAbslHashValue<T>::value) {
    return AbslHashValue(std::move(s), v);
  } else if constexpr (IsTuple<T>::value) {
    return std::apply([&](const auto&... e) { return HashState::combine(std::move(s), e...); }, v);
  } else if constexpr (IsStdArray<T>::value) {
    for (const auto& e : v) s = HashState::combine(std::move(s), e);
    return s;
  } else if constexpr (std::is_same_v<T, unsigned __int128>) {
    Mix(s, static_cast<size_t>(v >> 64));
    Mix(s, static_cast<size_t>(v));
    return s;
  } else if constexpr (std::is_enum_v<T>) {
    Mix(s, static_cast<size_t>(v));
    return s;
  } else {
    Mix(s, std::hash<T>{}(v));
    return s;
  }
}
template <typename T, typename... Ts>
HashState HashState::combine(HashState s, const T& v, const Ts&... rest) {
  return combine(HashOne(std::move(s), v), rest...);
}
}  // namespace shim
template <typename T>
struct Hash {
  size_t operator()(const T& v) const { return shim::HashOne(shim::HashState{}, v).h; }
};
template <typename... Ts>
size_t HashOf(const Ts&... vs) {
  return shim::HashState::combine(shim::HashState{}, vs...).h;
}
}  // namespace absl
