// A minimal stand-in for <gtest/gtest.h> (GoogleTest is not in this image): just enough of the value-parameterised
// test machinery for the reference's unit tests (test/blas_test_macros.hpp, test/unittest/blas3/*.cpp) to compile and
// run UNCHANGED against this repository's headers: TestWithParam / TEST_P / INSTANTIATE_TEST_SUITE_P with
// Combine / Values / Range generators and a name generator, ASSERT_* / EXPECT_* / GTEST_SKIP, InitGoogleTest with
// --gtest_filter (substring or '*' wildcards) and --gtest_list_tests, RUN_ALL_TESTS with a GoogleTest-style report.
#pragma once
#include <cstdio>
#include <cstring>
#include <functional>
#include <iostream>
#include <memory>
#include <sstream>
#include <string>
#include <tuple>
#include <typeindex>
#include <utility>
#include <vector>

namespace testing {

class Test {
 public:
  virtual ~Test() {}
  virtual void SetUp() {}
  virtual void TearDown() {}
  virtual void TestBody() = 0;
};

template <typename T>
class WithParamInterface {
 public:
  using ParamType = T;
  static const ParamType& GetParam() { return *current(); }
  static const ParamType*& current() {
    static const ParamType* p = nullptr;
    return p;
  }
};

template <typename T>
class TestWithParam : public Test, public WithParamInterface<T> {};

template <typename T>
struct TestParamInfo {
  TestParamInfo(const T& p, size_t i) : param(p), index(i) {}
  T param;
  size_t index;
};

// ---- generators -------------------------------------------------------------------------------------------
template <typename T>
struct ParamGenerator {
  std::vector<T> values;
};

namespace internal {

template <typename To, typename From>
To convert_value(const From& f) {
  if constexpr (std::is_constructible_v<To, From>) return To(f);
  else return static_cast<To>(f);
}

template <typename... Ts>
struct ValueArray {
  std::tuple<Ts...> v;
  template <typename T>
  operator ParamGenerator<T>() const {
    ParamGenerator<T> g;
    std::apply([&](const auto&... e) { (g.values.push_back(convert_value<T>(e)), ...); }, v);
    return g;
  }
};

template <typename B>
struct RangeGen {
  B begin, end, step;
  template <typename T>
  operator ParamGenerator<T>() const {
    ParamGenerator<T> g;
    for (B x = begin; x < end; x = static_cast<B>(x + step)) g.values.push_back(static_cast<T>(x));
    return g;
  }
};

template <typename... Gs>
struct CartesianProduct {
  std::tuple<Gs...> gens;
  template <typename... Ts>
  operator ParamGenerator<std::tuple<Ts...>>() const {
    static_assert(sizeof...(Ts) == sizeof...(Gs), "Combine(): generator count must match the tuple arity");
    return build<Ts...>(std::index_sequence_for<Ts...>{});
  }

 private:
  template <typename... Ts, size_t... I>
  ParamGenerator<std::tuple<Ts...>> build(std::index_sequence<I...>) const {
    std::tuple<std::vector<Ts>...> cols{ParamGenerator<Ts>(std::get<I>(gens)).values...};
    ParamGenerator<std::tuple<Ts...>> out;
    const size_t sizes[] = {std::get<I>(cols).size()...};
    size_t total = 1;
    for (size_t s : sizes) total *= s;
    for (size_t lin = 0; lin < total; ++lin) {   // the LAST generator varies fastest, as in GoogleTest
      size_t idx[sizeof...(Ts)];
      size_t rem = lin;
      for (size_t d = sizeof...(Ts); d-- > 0;) { idx[d] = rem % sizes[d]; rem /= sizes[d]; }
      out.values.emplace_back(std::get<I>(cols)[idx[I]]...);
    }
    return out;
  }
};

// ---- registry ---------------------------------------------------------------------------------------------
struct Result {
  bool failed = false, skipped = false;
};
inline Result& current_result() {
  static Result r;
  return r;
}

struct RunnableTest {
  std::string full_name;
  std::function<void()> run;
};

struct SuiteBase {
  virtual ~SuiteBase() {}
  virtual void expand(std::vector<RunnableTest>& out) = 0;
};

inline std::vector<SuiteBase*>& suites() {
  static std::vector<SuiteBase*> s;
  return s;
}

template <typename Fixture>
struct Suite : SuiteBase {
  using P = typename Fixture::ParamType;
  struct Pattern { std::string name; std::function<Test*()> make; };
  struct Instantiation {
    std::string prefix;
    std::function<std::vector<P>()> params;
    std::function<std::string(const TestParamInfo<P>&)> namer;
  };
  std::string suite_name;
  std::vector<Pattern> patterns;
  std::vector<Instantiation> insts;
  static Suite& get(const char* name) {
    static Suite* s = [&] { auto* p = new Suite; p->suite_name = name; suites().push_back(p); return p; }();
    return *s;
  }
  void expand(std::vector<RunnableTest>& out) override {
    for (auto& in : insts) {
      auto params = std::make_shared<std::vector<P>>(in.params());
      for (auto& pat : patterns)
        for (size_t i = 0; i < params->size(); ++i) {
          std::string pname = in.namer ? in.namer(TestParamInfo<P>((*params)[i], i)) : std::to_string(i);
          RunnableTest t;
          t.full_name = in.prefix + "/" + suite_name + "." + pat.name + "/" + pname;
          auto make = pat.make;
          t.run = [params, i, make] {
            WithParamInterface<P>::current() = &(*params)[i];
            std::unique_ptr<Test> obj(make());
            obj->SetUp();
            if (!current_result().skipped && !current_result().failed) obj->TestBody();
            obj->TearDown();
          };
          out.push_back(std::move(t));
        }
    }
  }
};

struct MessageSink {   // `ASSERT_TRUE(x) << "why"` / `GTEST_SKIP() << "why"`
  std::ostringstream os;
  const char* kind;
  const char* file;
  int line;
  std::string expr;
  MessageSink(const char* k, const char* f, int l, const char* e) : kind(k), file(f), line(l), expr(e) {}
  template <typename T>
  MessageSink& operator<<(const T& v) { os << v; return *this; }
};
struct Reporter {      // `return Reporter() = MessageSink(...) << ...;` -- the GoogleTest idiom for void functions
  void operator=(const MessageSink& m) const {
    if (std::strcmp(m.kind, "skip") == 0) {
      current_result().skipped = true;
      if (!m.os.str().empty()) std::cout << m.file << ":" << m.line << ": Skipped: " << m.os.str() << std::endl;
    } else {
      current_result().failed = true;
      std::cout << m.file << ":" << m.line << ": Failure\n  " << m.expr << "\n  " << m.os.str() << std::endl;
    }
  }
};

inline bool wildcard_match(const char* pat, const char* s) {
  if (*pat == 0) return *s == 0;
  if (*pat == '*') return wildcard_match(pat + 1, s) || (*s && wildcard_match(pat, s + 1));
  if (*pat == '?') return *s && wildcard_match(pat + 1, s + 1);
  return *pat == *s && wildcard_match(pat + 1, s + 1);
}

struct Options {
  std::string filter = "*";
  bool list_only = false;
};
inline Options& options() {
  static Options o;
  return o;
}

inline bool filter_accepts(const std::string& name) {
  const std::string& f = options().filter;
  const size_t dash = f.find('-');
  const std::string pos = f.substr(0, dash), neg = dash == std::string::npos ? "" : f.substr(dash + 1);
  auto any = [&](const std::string& list) {
    std::stringstream ss(list);
    std::string pat;
    while (std::getline(ss, pat, ':'))
      if (!pat.empty() && wildcard_match(pat.c_str(), name.c_str())) return true;
    return false;
  };
  return (pos.empty() || any(pos)) && !(neg.size() && any(neg));
}

inline int run_all() {
  std::vector<RunnableTest> tests;
  for (SuiteBase* s : suites()) s->expand(tests);
  int ran = 0, failed = 0, skipped = 0;
  std::vector<std::string> failures;
  for (auto& t : tests) {
    if (!filter_accepts(t.full_name)) continue;
    if (options().list_only) { std::cout << t.full_name << "\n"; continue; }
    current_result() = Result();
    std::cout << "[ RUN      ] " << t.full_name << std::endl;
    try {
      t.run();
    } catch (const std::exception& e) {
      current_result().failed = true;
      std::cout << "unexpected exception: " << e.what() << std::endl;
    }
    ++ran;
    if (current_result().failed) { ++failed; failures.push_back(t.full_name); std::cout << "[  FAILED  ] "; }
    else if (current_result().skipped) { ++skipped; std::cout << "[  SKIPPED ] "; }
    else std::cout << "[       OK ] ";
    std::cout << t.full_name << std::endl;
  }
  if (options().list_only) return 0;
  std::cout << "[==========] " << ran << " tests ran.\n[  PASSED  ] " << (ran - failed - skipped) << " tests." << std::endl;
  if (skipped) std::cout << "[  SKIPPED ] " << skipped << " tests." << std::endl;
  if (failed) {
    std::cout << "[  FAILED  ] " << failed << " tests, listed below:" << std::endl;
    for (auto& n : failures) std::cout << "[  FAILED  ] " << n << std::endl;
  }
  return failed ? 1 : 0;
}

}  // namespace internal

template <typename... Ts>
internal::ValueArray<Ts...> Values(Ts... v) { return {std::tuple<Ts...>(v...)}; }
template <typename T>
internal::RangeGen<T> Range(T begin, T end) { return {begin, end, T(1)}; }
template <typename T, typename S>
internal::RangeGen<T> Range(T begin, T end, S step) { return {begin, end, static_cast<T>(step)}; }
template <typename... Gs>
internal::CartesianProduct<Gs...> Combine(const Gs&... g) { return {std::tuple<Gs...>(g...)}; }
inline internal::ValueArray<bool, bool> Bool() { return Values(false, true); }

inline void InitGoogleTest(int* argc, char** argv) {
  int w = 1;
  for (int i = 1; i < *argc; ++i) {
    const std::string a = argv[i];
    if (a.rfind("--gtest_filter=", 0) == 0) internal::options().filter = a.substr(15);
    else if (a == "--gtest_list_tests") internal::options().list_only = true;
    else if (a.rfind("--gtest_", 0) == 0) { /* other GoogleTest flags are accepted and ignored */ }
    else argv[w++] = argv[i];
  }
  *argc = w;
}

}  // namespace testing

#define RUN_ALL_TESTS() ::testing::internal::run_all()

#define GTEST_SHIM_CLASS_(suite, name) suite##_##name##_Test

#define TEST_P(suite, name)                                                                             \
  class GTEST_SHIM_CLASS_(suite, name) : public suite {                                                 \
   public:                                                                                              \
    void TestBody() override;                                                                           \
    static int registered_;                                                                             \
  };                                                                                                    \
  int GTEST_SHIM_CLASS_(suite, name)::registered_ = [] {                                                \
    ::testing::internal::Suite<suite>::get(#suite).patterns.push_back(                                  \
        {#name, [] { return static_cast<::testing::Test*>(new GTEST_SHIM_CLASS_(suite, name)); }});    \
    return 0;                                                                                           \
  }();                                                                                                  \
  void GTEST_SHIM_CLASS_(suite, name)::TestBody()

#define GTEST_SHIM_CAT2_(a, b) a##b
#define GTEST_SHIM_CAT_(a, b) GTEST_SHIM_CAT2_(a, b)

#define INSTANTIATE_TEST_SUITE_P(prefix_, suite_, ...) \
  GTEST_SHIM_INSTANTIATE_(prefix_, suite_, __VA_ARGS__, nullptr, nullptr)
#define INSTANTIATE_TEST_CASE_P INSTANTIATE_TEST_SUITE_P

#define GTEST_SHIM_INSTANTIATE_(prefix_, suite_, generator_, namer_, ...)                                          \
  static int GTEST_SHIM_CAT_(gtest_shim_inst_##prefix_##_##suite_##_, __LINE__) = [] {                             \
    using P_ = typename suite_::ParamType;                                                                         \
    typename ::testing::internal::Suite<suite_>::Instantiation in_;                                                \
    in_.prefix = #prefix_;                                                                                         \
    in_.params = [] { return ::testing::ParamGenerator<P_>(generator_).values; };                                  \
    in_.namer = ::testing::internal::make_namer<P_>(namer_);                                                       \
    ::testing::internal::Suite<suite_>::get(#suite_).insts.push_back(in_);                                         \
    return 0;                                                                                                      \
  }()

namespace testing {
namespace internal {
template <typename P, typename F>
std::function<std::string(const TestParamInfo<P>&)> make_namer(F f) {
  return [f](const TestParamInfo<P>& i) { return std::string(f(i)); };
}
template <typename P>
std::function<std::string(const TestParamInfo<P>&)> make_namer(std::nullptr_t) {
  return nullptr;
}
}  // namespace internal
}  // namespace testing

#define GTEST_SHIM_FAIL_(kind, expr) \
  return ::testing::internal::Reporter() = ::testing::internal::MessageSink(kind, __FILE__, __LINE__, expr)
#define GTEST_SHIM_NONFATAL_(expr) \
  ::testing::internal::Reporter() = ::testing::internal::MessageSink("fail", __FILE__, __LINE__, expr)

#define GTEST_SKIP() GTEST_SHIM_FAIL_("skip", "")
#define GTEST_FAIL() GTEST_SHIM_FAIL_("fail", "Failed")
#define FAIL() GTEST_FAIL()
#define SUCCEED() (void)0

#define ASSERT_TRUE(c) if (c) ; else GTEST_SHIM_FAIL_("fail", "Value of: " #c "\n    Actual: false\n  Expected: true")
#define ASSERT_FALSE(c) if (!(c)) ; else GTEST_SHIM_FAIL_("fail", "Value of: " #c "\n    Actual: true\n  Expected: false")
#define ASSERT_EQ(a, b) if ((a) == (b)) ; else GTEST_SHIM_FAIL_("fail", "Expected equality of " #a " and " #b)
#define ASSERT_NE(a, b) if ((a) != (b)) ; else GTEST_SHIM_FAIL_("fail", "Expected " #a " != " #b)
#define ASSERT_LT(a, b) if ((a) < (b)) ; else GTEST_SHIM_FAIL_("fail", "Expected " #a " < " #b)
#define ASSERT_LE(a, b) if ((a) <= (b)) ; else GTEST_SHIM_FAIL_("fail", "Expected " #a " <= " #b)
#define ASSERT_GT(a, b) if ((a) > (b)) ; else GTEST_SHIM_FAIL_("fail", "Expected " #a " > " #b)
#define ASSERT_GE(a, b) if ((a) >= (b)) ; else GTEST_SHIM_FAIL_("fail", "Expected " #a " >= " #b)
#define EXPECT_TRUE(c) if (c) ; else GTEST_SHIM_NONFATAL_("Value of: " #c "\n    Actual: false\n  Expected: true")
#define EXPECT_FALSE(c) if (!(c)) ; else GTEST_SHIM_NONFATAL_("Value of: " #c "\n    Actual: true\n  Expected: false")
#define EXPECT_EQ(a, b) if ((a) == (b)) ; else GTEST_SHIM_NONFATAL_("Expected equality of " #a " and " #b)
#define EXPECT_NE(a, b) if ((a) != (b)) ; else GTEST_SHIM_NONFATAL_("Expected " #a " != " #b)
