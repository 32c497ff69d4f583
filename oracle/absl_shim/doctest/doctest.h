// Build scaffolding for oracle/_ref ONLY: the handful of doctest macros the reference's tests use (TEST_CASE, CHECK*, REQUIRE*),
// so that an UNMODIFIED reference test file compiles into an executable.  Failures are counted and reported; exit code = failures.
// DOCTEST_SHIM_FILTER (env): run only the test cases whose name contains the string.
#pragma once
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>
namespace doctest_shim {
struct Case {
  const char* name;
  void (*fn)();
};
inline std::vector<Case>& Cases() {
  static std::vector<Case> c;
  return c;
}
inline int& Failures() {
  static int f = 0;
  return f;
}
struct Registrar {
  Registrar(const char* name, void (*fn)()) { Cases().push_back({name, fn}); }
};
inline void Fail(const char* file, int line, const char* expr, const std::string& msg) {
  ++Failures();
  std::cerr << file << ":" << line << ": CHECK failed: " << expr << (msg.empty() ? "" : " -- ") << msg << std::endl;
}
inline int RunAll() {
  const char* filter = std::getenv("DOCTEST_SHIM_FILTER");
  int ran = 0;
  for (const Case& c : Cases()) {
    if (filter && !std::strstr(c.name, filter)) continue;
    std::cerr << "[doctest-shim] TEST_CASE: " << c.name << std::endl;
    const int before = Failures();
    c.fn();
    std::cerr << "[doctest-shim]   " << (Failures() == before ? "ok" : "FAILED") << std::endl;
    ++ran;
  }
  std::cerr << "[doctest-shim] " << ran << " test case(s), " << Failures() << " failed check(s)" << std::endl;
  return Failures();
}
}  // namespace doctest_shim
#define DOCTEST_SHIM_CAT2(a, b) a##b
#define DOCTEST_SHIM_CAT(a, b) DOCTEST_SHIM_CAT2(a, b)
#define DOCTEST_SHIM_TEST(fn, name)                                        \
  static void fn();                                                        \
  static ::doctest_shim::Registrar DOCTEST_SHIM_CAT(fn, _reg)(name, &fn); \
  static void fn()
#define TEST_CASE(name) DOCTEST_SHIM_TEST(DOCTEST_SHIM_CAT(doctest_shim_case_, __LINE__), name)
#define CHECK_MESSAGE(cond, msg)                                        \
  do {                                                                  \
    if (!(cond)) {                                                      \
      std::ostringstream _ss;                                           \
      _ss << msg;                                                       \
      ::doctest_shim::Fail(__FILE__, __LINE__, #cond, _ss.str());      \
    }                                                                   \
  } while (0)
#define REQUIRE_MESSAGE(cond, msg) CHECK_MESSAGE(cond, msg)
#ifndef CHECK
#define CHECK(cond) CHECK_MESSAGE(cond, "")
#endif
#define REQUIRE(cond) CHECK_MESSAGE(cond, "")
#define CHECK_EQ_DT(a, b) CHECK_MESSAGE((a) == (b), "")
#ifdef DOCTEST_CONFIG_IMPLEMENT_WITH_MAIN
int main() { return ::doctest_shim::RunAll(); }
#endif
