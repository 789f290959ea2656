// check.hpp -- minimal gtest-shaped harness (gtest is not in this image):
// TEST(suite, name) registers a case, EXPECT_* record failures, main() runs
// everything between ipcl::initializeContext / terminateContext exactly as the
// reference's test/main.cpp:9-45 does.
#pragma once
#include <cstdio>
#include <exception>
#include <functional>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

namespace check {
struct Case {
  std::string name;
  std::function<void()> fn;
};
inline std::vector<Case>& cases() {
  static std::vector<Case> c;
  return c;
}
inline int& failures() {
  static int f = 0;
  return f;
}
struct Reg {
  Reg(const char* s, const char* n, std::function<void()> fn) {
    cases().push_back({std::string(s) + "." + n, fn});
  }
};
template <typename A, typename B>
void expect_eq(const A& a, const B& b, const char* ea, const char* eb,
               const char* file, int line) {
  if (!(a == b)) {
    std::ostringstream os;
    os << file << ":" << line << ": EXPECT_EQ(" << ea << ", " << eb << ") failed: "
       << a << " vs " << b;
    std::cout << os.str() << std::endl;
    failures()++;
  }
}
inline void expect_true(bool v, const char* e, const char* file, int line) {
  if (!v) {
    std::cout << file << ":" << line << ": EXPECT_TRUE(" << e << ") failed" << std::endl;
    failures()++;
  }
}
inline int run_all(const std::string& filter) {
  int ran = 0, failed_cases = 0;
  for (auto& c : cases()) {
    if (!filter.empty() && c.name.find(filter) == std::string::npos) continue;
    int before = failures();
    std::cout << "[ RUN      ] " << c.name << std::endl;
    try {
      c.fn();
    } catch (const std::exception& e) {
      std::cout << "exception: " << e.what() << std::endl;
      failures()++;
    }
    bool ok = failures() == before;
    if (!ok) failed_cases++;
    std::cout << (ok ? "[       OK ] " : "[  FAILED  ] ") << c.name << std::endl;
    ran++;
  }
  std::cout << "[==========] " << ran << " tests ran, " << failed_cases << " failed"
            << std::endl;
  return failed_cases;
}
}  // namespace check

#define TEST(suite, name)                                              \
  static void suite##_##name##_body();                                 \
  static check::Reg suite##_##name##_reg(#suite, #name, suite##_##name##_body); \
  static void suite##_##name##_body()
#define EXPECT_EQ(a, b) check::expect_eq((a), (b), #a, #b, __FILE__, __LINE__)
#define EXPECT_TRUE(a) check::expect_true((a), #a, __FILE__, __LINE__)
#define EXPECT_THROW(stmt)                                                 \
  do {                                                                     \
    bool thrown_ = false;                                                  \
    try {                                                                  \
      stmt;                                                                \
    } catch (const std::exception&) {                                      \
      thrown_ = true;                                                      \
    }                                                                      \
    check::expect_true(thrown_, "throws: " #stmt, __FILE__, __LINE__);     \
  } while (0)
