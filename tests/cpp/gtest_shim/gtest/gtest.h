// gtest.h -- a minimal stand-in for GoogleTest (not in this image), just large
// enough to compile the reference's own unit tests UNMODIFIED:
// /root/reference/test/{main,test_cryptography,test_ops,test_serialization}.cpp
// use TEST, EXPECT_EQ, ::testing::InitGoogleTest and RUN_ALL_TESTS only.
// Output mimics gtest's ("[ RUN      ]", "[       OK ]", "[  FAILED  ]") so the
// log reads the same.  Optional first argument: a substring filter on
// "Suite.Name" (or --gtest_filter=Suite.*).
#ifndef IPCL_B200_GTEST_SHIM_H_
#define IPCL_B200_GTEST_SHIM_H_

#include <algorithm>
#include <cstdio>
#include <functional>
#include <limits>
#include <memory>
#include <cstring>
#include <exception>
#include <iostream>
#include <sstream>
#include <string>
#include <type_traits>
#include <utility>
#include <vector>

namespace testing {

struct TestCase {
  const char* suite;
  const char* name;
  void (*fn)();
};

inline std::vector<TestCase>& registry() {
  static std::vector<TestCase> r;
  return r;
}
inline int& current_failures() {
  static int f = 0;
  return f;
}
inline std::string& filter() {
  static std::string f;
  return f;
}

struct Registrar {
  Registrar(const char* suite, const char* name, void (*fn)()) {
    registry().push_back({suite, name, fn});
  }
};

inline void InitGoogleTest(int* argc, char** argv) {
  for (int i = 1; i < *argc; i++) {
    std::string a = argv[i];
    const std::string pre = "--gtest_filter=";
    if (a.compare(0, pre.size(), pre) == 0) a = a.substr(pre.size());
    while (!a.empty() && (a.back() == '*' || a.back() == '.')) a.pop_back();
    if (!a.empty() && a[0] != '-') filter() = a;
  }
}

template <typename T, typename = void>
struct Streamable : std::false_type {};
template <typename T>
struct Streamable<T, decltype(void(std::declval<std::ostream&>() << std::declval<const T&>()))>
    : std::true_type {};

template <typename T>
typename std::enable_if<Streamable<T>::value, std::string>::type show(const T& v) {
  std::ostringstream os;
  os << v;
  return os.str();
}
template <typename T>
typename std::enable_if<!Streamable<T>::value, std::string>::type show(const T&) {
  return "<value>";
}

template <typename A, typename B>
void expect_eq(const A& a, const B& b, const char* ea, const char* eb, const char* file,
               int line) {
  if (a == b) return;
  current_failures()++;
  std::cout << file << ":" << line << ": Failure\nExpected equality of these values:\n  "
            << ea << "\n    Which is: " << show(a) << "\n  " << eb
            << "\n    Which is: " << show(b) << std::endl;
}

inline int run_all() {
  int ran = 0, failed = 0;
  std::vector<std::string> failed_names;
  for (const auto& t : registry()) {
    const std::string full = std::string(t.suite) + "." + t.name;
    if (!filter().empty() && full.find(filter()) == std::string::npos) continue;
    std::cout << "[ RUN      ] " << full << std::endl;
    current_failures() = 0;
    try {
      t.fn();
    } catch (const std::exception& e) {
      current_failures()++;
      std::cout << "unexpected exception: " << e.what() << std::endl;
    } catch (...) {
      current_failures()++;
      std::cout << "unexpected exception" << std::endl;
    }
    ran++;
    if (current_failures()) {
      failed++;
      failed_names.push_back(full);
      std::cout << "[  FAILED  ] " << full << std::endl;
    } else {
      std::cout << "[       OK ] " << full << std::endl;
    }
  }
  std::cout << "[==========] " << ran << " tests ran." << std::endl;
  std::cout << "[  PASSED  ] " << (ran - failed) << " tests." << std::endl;
  for (const auto& n : failed_names) std::cout << "[  FAILED  ] " << n << std::endl;
  return failed ? 1 : 0;
}

}  // namespace testing

#define TEST(suite, name)                                                        \
  static void suite##_##name##_body();                                           \
  static ::testing::Registrar suite##_##name##_reg(#suite, #name,                \
                                                   &suite##_##name##_body);      \
  static void suite##_##name##_body()

#define EXPECT_EQ(a, b) ::testing::expect_eq((a), (b), #a, #b, __FILE__, __LINE__)
#define RUN_ALL_TESTS() ::testing::run_all()

#endif  // IPCL_B200_GTEST_SHIM_H_
