// Build scaffolding for oracle/_ref ONLY: CHECK/DCHECK/LOG stream stand-ins.
#pragma once
#include <cstdlib>
#include <iostream>
#include <sstream>
namespace absl_shim {
struct FatalStream {
  std::ostringstream ss;
  const char* file; int line; const char* cond;
  FatalStream(const char* f, int l, const char* c) : file(f), line(l), cond(c) {}
  template <typename T> FatalStream& operator<<(const T& v) { ss << v; return *this; }
  [[noreturn]] ~FatalStream() {
    std::cerr << file << ":" << line << " CHECK failed: " << cond << " " << ss.str() << std::endl;
    std::abort();
  }
};
struct NullStream {
  template <typename T> NullStream& operator<<(const T&) { return *this; }
};
struct LogStream {
  std::ostringstream ss; bool fatal;
  explicit LogStream(bool f) : fatal(f) {}
  template <typename T> LogStream& operator<<(const T& v) { ss << v; return *this; }
  ~LogStream() { std::cerr << ss.str() << std::endl; if (fatal) std::abort(); }
};
}  // namespace absl_shim
#define CHECK(cond) \
  if (cond) {} else ::absl_shim::FatalStream(__FILE__, __LINE__, #cond)
#define CHECK_EQ(a, b) CHECK((a) == (b))
#define CHECK_NE(a, b) CHECK((a) != (b))
#define CHECK_LT(a, b) CHECK((a) < (b))
#define CHECK_LE(a, b) CHECK((a) <= (b))
#define CHECK_GT(a, b) CHECK((a) > (b))
#define CHECK_GE(a, b) CHECK((a) >= (b))
#ifdef NDEBUG
#define DCHECK(cond) while (false) ::absl_shim::NullStream()
#else
#define DCHECK(cond) CHECK(cond)
#endif
#define DCHECK_EQ(a, b) DCHECK((a) == (b))
#define DCHECK_NE(a, b) DCHECK((a) != (b))
#define DCHECK_LT(a, b) DCHECK((a) < (b))
#define DCHECK_LE(a, b) DCHECK((a) <= (b))
#define DCHECK_GT(a, b) DCHECK((a) > (b))
#define DCHECK_GE(a, b) DCHECK((a) >= (b))
