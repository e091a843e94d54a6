// The reference generates git_config.h at configure time (benchmark/make_git_config.sh) and compiles it into the small
// bench_info library whose two symbols label every benchmark (common/include/common/set_benchmark_label.hpp:43-44,
// 193-194).  Building its harness here without its build system needs those symbols from somewhere: this file.
const char* commit_date = "n/a (portblas_b200 build of the reference harness)";
const char* commit_hash = "portblas_b200";
