// A minimal stand-in for <benchmark/benchmark.h> (Google Benchmark is not in this image): just the part of its API the
// reference's benchmark harness uses (benchmark/portblas/main.cpp, blas3/*.cpp, common/include/common/common_utils.hpp),
// so that those sources compile and run UNCHANGED against this repository's library:
//   benchmark::Initialize / RunSpecifiedBenchmarks, RegisterBenchmark(name, fn, args...)->UseRealTime(),
//   benchmark::State as a range (`for (auto _ : state)`), counters, iterations(), PauseTiming / ResumeTiming,
//   SetLabel, SkipWithError, SetItemsProcessed, SetBytesProcessed;
//   --benchmark_filter=<regex>, --benchmark_min_time=<seconds>, --benchmark_format=console|json, --benchmark_out=<file>,
//   --benchmark_list_tests.  Reports follow Google Benchmark's console and JSON layouts (real_time in ns).
#pragma once
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <fstream>
#include <functional>
#include <iostream>
#include <map>
#include <memory>
#include <regex>
#include <sstream>
#include <string>
#include <vector>

namespace benchmark {

struct Counter {
  double value = 0.0;
  Counter() = default;
  Counter(double v) : value(v) {}
  operator double() const { return value; }
  Counter& operator=(double v) { value = v; return *this; }
  Counter& operator+=(double v) { value += v; return *this; }
  Counter& operator-=(double v) { value -= v; return *this; }
};

class State {
 public:
  using clock = std::chrono::steady_clock;
  explicit State(double min_time_s) : min_time_(min_time_s) {}

  struct Iterator {
    State* s;
    bool operator!=(const Iterator&) const { return s->keep_running(); }
    void operator++() {}
    struct Unused {
      ~Unused() {}   // non-trivial: silences -Wunused-variable on `for (auto _ : state)`
    };
    Unused operator*() const { return Unused(); }
  };
  Iterator begin() { start(); return Iterator{this}; }
  Iterator end() { return Iterator{this}; }
  bool KeepRunning() {
    if (!started_) start();
    return keep_running();
  }

  void PauseTiming() { if (running_) { elapsed_ += clock::now() - t0_; running_ = false; } }
  void ResumeTiming() { if (!running_) { t0_ = clock::now(); running_ = true; } }
  void SkipWithError(const char* msg) { error_ = true; error_msg_ = msg ? msg : ""; }
  void SetLabel(const std::string& l) { label_ = l; }
  void SetItemsProcessed(int64_t n) { items_ = n; }
  void SetBytesProcessed(int64_t n) { bytes_ = n; }
  int64_t iterations() const { return iters_; }
  int64_t items_processed() const { return items_; }
  int64_t bytes_processed() const { return bytes_; }

  std::map<std::string, Counter> counters;

  // results (read by the runner)
  double real_seconds() const { return std::chrono::duration<double>(elapsed_).count(); }
  bool error() const { return error_; }
  const std::string& error_message() const { return error_msg_; }
  const std::string& label() const { return label_; }

 private:
  void start() { started_ = true; iters_ = 0; elapsed_ = clock::duration::zero(); t0_ = clock::now(); running_ = true; }
  bool keep_running() {
    if (error_) { finish(); return false; }
    if (iters_ > 0) {
      const double el = std::chrono::duration<double>(elapsed_ + (running_ ? clock::now() - t0_ : clock::duration::zero())).count();
      if ((el >= min_time_ && iters_ >= 1) || iters_ >= 1000000) { finish(); return false; }
    }
    ++iters_;
    return true;
  }
  void finish() { PauseTiming(); }
  double min_time_;
  bool started_ = false, running_ = false, error_ = false;
  int64_t iters_ = 0, items_ = 0, bytes_ = 0;
  clock::time_point t0_;
  clock::duration elapsed_ = clock::duration::zero();
  std::string error_msg_, label_;
};

namespace internal {

class Benchmark {
 public:
  Benchmark(std::string n, std::function<void(State&)> f) : name(std::move(n)), fn(std::move(f)) {}
  Benchmark* UseRealTime() { return this; }
  Benchmark* UseManualTime() { return this; }
  Benchmark* Unit(int) { return this; }
  Benchmark* Iterations(int64_t) { return this; }
  Benchmark* MinTime(double) { return this; }
  std::string name;
  std::function<void(State&)> fn;
  std::unique_ptr<State> done;   // BENCHMARK_SHIM_RUN_AT_REGISTRATION: the finished run, reported later
};

inline std::vector<std::unique_ptr<Benchmark>>& registry() {
  static std::vector<std::unique_ptr<Benchmark>> r;
  return r;
}

struct Options {
  std::string filter = ".";
  std::string format = "console";
  std::string out_file;
  double min_time = 0.5;
  bool list_only = false;
};
inline Options& options() {
  static Options o;
  return o;
}

inline std::string json_escape(const std::string& s) {
  std::string o;
  for (char c : s) {
    if (c == '"' || c == '\\') { o += '\\'; o += c; }
    else if (c == '\n') o += "\\n";
    else o += c;
  }
  return o;
}

}  // namespace internal

// BENCHMARK_SHIM_RUN_AT_REGISTRATION: the benchmark body runs inside RegisterBenchmark and RunSpecifiedBenchmarks only
// reports it.  Needed for the reference's benchmark/portblas/blas3/symm.cpp, whose registered lambda captures the loop
// locals `side_c` / `uplo_c` BY REFERENCE (symm.cpp:139-147, `[&]`): by the time a deferred runner calls it they are dead
// stack slots (upstream bug, independent of the BLAS underneath).  Running at registration reads them while they live.
template <typename Lambda, typename... Args>
internal::Benchmark* RegisterBenchmark(const char* name, Lambda&& fn, Args&&... args) {
  auto bound = [fn, args...](State& st) mutable { fn(st, args...); };
  internal::registry().emplace_back(new internal::Benchmark(name, bound));
  internal::Benchmark* b = internal::registry().back().get();
#ifdef BENCHMARK_SHIM_RUN_AT_REGISTRATION
  auto& o = internal::options();
  if (!o.list_only && std::regex_search(b->name, std::regex(o.filter))) {
    b->done.reset(new State(o.min_time));
    b->fn(*b->done);
  }
#endif
  return b;
}
template <typename Lambda, typename... Args>
internal::Benchmark* RegisterBenchmark(const std::string& name, Lambda&& fn, Args&&... args) {
  return RegisterBenchmark(name.c_str(), std::forward<Lambda>(fn), std::forward<Args>(args)...);
}

inline void Initialize(int* argc, char** argv) {
  int w = 1;
  auto& o = internal::options();
  for (int i = 1; i < *argc; ++i) {
    const std::string a = argv[i];
    auto val = [&](const char* key) { return a.substr(std::string(key).size()); };
    if (a.rfind("--benchmark_filter=", 0) == 0) o.filter = val("--benchmark_filter=");
    else if (a.rfind("--benchmark_format=", 0) == 0) o.format = val("--benchmark_format=");
    else if (a.rfind("--benchmark_out=", 0) == 0) o.out_file = val("--benchmark_out=");
    else if (a.rfind("--benchmark_out_format=", 0) == 0) { /* the output file is always JSON */ }
    else if (a.rfind("--benchmark_min_time=", 0) == 0) o.min_time = std::atof(val("--benchmark_min_time=").c_str());
    else if (a == "--benchmark_list_tests" || a == "--benchmark_list_tests=true") o.list_only = true;
    else if (a.rfind("--benchmark_", 0) == 0) { /* other Google Benchmark flags are accepted and ignored */ }
    else argv[w++] = argv[i];
  }
  *argc = w;
}

inline size_t RunSpecifiedBenchmarks() {
  auto& o = internal::options();
  std::regex re(o.filter);
  std::ostringstream json;
  json << "{\n  \"context\": {\n    \"library_build_type\": \"shim\",\n    \"executable\": \"portblas benchmark (reference sources)\"\n  },\n"
       << "  \"benchmarks\": [\n";
  bool first = true;
  size_t ran = 0;
  if (o.format == "console" && !o.list_only)
    std::printf("%-90s %15s %15s %12s\n", "Benchmark", "Time", "CPU", "Iterations");
  for (auto& b : internal::registry()) {
    if (!std::regex_search(b->name, re)) continue;
    if (o.list_only) { std::cout << b->name << "\n"; continue; }
    State fresh(o.min_time);
    if (!b->done) b->fn(fresh);
    State& st = b->done ? *b->done : fresh;
    ++ran;
    const double it = st.iterations() > 0 ? (double)st.iterations() : 1.0;
    const double real_ns = st.real_seconds() * 1e9 / it;
    if (o.format == "console") {
      if (st.error()) std::printf("%-90s ERROR OCCURRED: '%s'\n", b->name.c_str(), st.error_message().c_str());
      else {
        std::printf("%-90s %12.0f ns %12.0f ns %12lld", b->name.c_str(), real_ns, real_ns, (long long)st.iterations());
        for (auto& kv : st.counters) std::printf(" %s=%g", kv.first.c_str(), (double)kv.second);
        if (st.items_processed() && st.real_seconds() > 0) std::printf(" items_per_second=%g", st.items_processed() / st.real_seconds());
        std::printf("\n");
      }
      std::fflush(stdout);
    }
    json << (first ? "" : ",\n") << "    {\n      \"name\": \"" << internal::json_escape(b->name) << "\",\n      \"run_name\": \""
         << internal::json_escape(b->name) << "\",\n      \"run_type\": \"iteration\",\n      \"repetitions\": 1,\n"
         << "      \"repetition_index\": 0,\n      \"threads\": 1,\n      \"iterations\": " << st.iterations() << ",\n";
    if (st.error())
      json << "      \"error_occurred\": true,\n      \"error_message\": \"" << internal::json_escape(st.error_message()) << "\",\n";
    json << "      \"real_time\": " << real_ns << ",\n      \"cpu_time\": " << real_ns << ",\n      \"time_unit\": \"ns\"";
    if (st.items_processed() && st.real_seconds() > 0) json << ",\n      \"items_per_second\": " << st.items_processed() / st.real_seconds();
    if (st.bytes_processed() && st.real_seconds() > 0) json << ",\n      \"bytes_per_second\": " << st.bytes_processed() / st.real_seconds();
    if (!st.label().empty()) json << ",\n      \"label\": \"" << internal::json_escape(st.label()) << "\"";
    for (auto& kv : st.counters) json << ",\n      \"" << internal::json_escape(kv.first) << "\": " << (double)kv.second;
    json << "\n    }";
    first = false;
  }
  json << "\n  ]\n}\n";
  if (!o.list_only) {
    if (o.format == "json") std::cout << json.str();
    if (!o.out_file.empty()) std::ofstream(o.out_file) << json.str();
  }
  return ran;
}

inline void Shutdown() {}

}  // namespace benchmark
