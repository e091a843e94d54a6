// gemm_b200.cpp -- C++ caller of the B200 GEMM path through the reference's interface
// (blas::SB_Handle, BufferIterator / USM containers, blas::helper, blas::_gemm*).
// Exercises buffer and USM containers, float / double / half / half->float, the three entry
// points, dependency events and event profiling, and checks every result against a plain host
// triple loop.  Exit code 0 == all checks passed.
#include "portblas.hpp"
#include <sycl/sycl.hpp>

#include <cmath>
#include <complex>
#include <cstdio>
#include <random>
#include <vector>

namespace {

template <typename T> double to_d(T v) { return static_cast<double>(v); }
template <> double to_d<sycl::half>(sycl::half v) { return static_cast<double>(__half2float(v)); }
template <typename T> T from_d(double v) { return static_cast<T>(v); }
template <> sycl::half from_d<sycl::half>(double v) { return __float2half(static_cast<float>(v)); }

template <typename Tin, typename Tout>
void host_gemm(char ta, char tb, int m, int n, int k, double alpha, const std::vector<Tin>& a, int lda,
               const std::vector<Tin>& b, int ldb, double beta, std::vector<Tout>& c, int ldc, size_t oa = 0,
               size_t ob = 0, size_t oc = 0) {
  for (int j = 0; j < n; ++j)
    for (int i = 0; i < m; ++i) {
      double acc = 0;
      for (int l = 0; l < k; ++l) {
        const double av = to_d(ta == 'n' ? a[oa + i + (size_t)l * lda] : a[oa + l + (size_t)i * lda]);
        const double bv = to_d(tb == 'n' ? b[ob + l + (size_t)j * ldb] : b[ob + j + (size_t)l * ldb]);
        acc += av * bv;
      }
      const double prev = beta == 0.0 ? 0.0 : beta * to_d(c[oc + i + (size_t)j * ldc]);
      c[oc + i + (size_t)j * ldc] = from_d<Tout>(alpha * acc + prev);
    }
}

template <typename T> void fill(std::vector<T>& v, unsigned seed) {
  std::mt19937 gen(seed);
  std::uniform_real_distribution<double> dis(-2.0, 5.0);
  for (auto& e : v) e = from_d<T>(dis(gen));
}

template <typename T>
bool close(const std::vector<T>& got, const std::vector<T>& want, double rel, const char* what) {
  double worst = 0;
  for (size_t i = 0; i < got.size(); ++i) {
    const double g = to_d(got[i]), w = to_d(want[i]);
    worst = std::max(worst, std::fabs(g - w) / (std::fabs(w) + 1.0));
  }
  const bool ok = worst <= rel;
  std::printf("%-46s max rel err %.3e  %s\n", what, worst, ok ? "PASS" : "FAIL");
  return ok;
}

template <typename Tin, typename Tout, blas::helper::AllocType mem>
bool run_gemm(blas::SB_Handle& sb, char ta, char tb, int m, int n, int k, double rel, const char* what) {
  using namespace blas;
  auto q = sb.get_queue();
  const int lda = (ta == 'n' ? m : k) + 3, ldb = (tb == 'n' ? k : n) + 1, ldc = m + 2;
  std::vector<Tin> a((size_t)lda * (ta == 'n' ? k : m)), b((size_t)ldb * (tb == 'n' ? n : k));
  std::vector<Tout> c((size_t)ldc * n), want;
  fill(a, 1); fill(b, 2); fill(c, 3);
  want = c;
  const Tout alpha = from_d<Tout>(1.5), beta = from_d<Tout>(0.5);
  host_gemm(ta, tb, m, n, k, 1.5, a, lda, b, ldb, 0.5, want, ldc);
  auto a_d = helper::allocate<mem, Tin>((int)a.size(), q);
  auto b_d = helper::allocate<mem, Tin>((int)b.size(), q);
  auto c_d = helper::allocate<mem, Tout>((int)c.size(), q);
  auto e1 = helper::copy_to_device(q, a.data(), a_d, a.size());
  auto e2 = helper::copy_to_device(q, b.data(), b_d, b.size());
  auto e3 = helper::copy_to_device(q, c.data(), c_d, c.size());
  auto ev = _gemm(sb, ta, tb, m, n, k, alpha, a_d, lda, b_d, ldb, beta, c_d, ldc, {e1, e2, e3});
  sb.wait(ev);
  auto e4 = helper::copy_to_host(q, c_d, c.data(), c.size());
  sb.wait(e4);
  helper::deallocate<mem>(a_d, q); helper::deallocate<mem>(b_d, q); helper::deallocate<mem>(c_d, q);
  return close(c, want, rel, what);
}

bool run_batched(blas::SB_Handle& sb) {
  using namespace blas;
  auto q = sb.get_queue();
  const int m = 63, n = 40, k = 128, batch = 5, lda = m, ldb = k, ldc = m;
  std::vector<float> a((size_t)lda * k * batch), b((size_t)ldb * n * batch), c((size_t)ldc * n * batch), want;
  fill(a, 4); fill(b, 5); fill(c, 6);
  want = c;
  for (int i = 0; i < batch; ++i)
    host_gemm('n', 'n', m, n, k, 3.0, a, lda, b, ldb, 7.0, want, ldc, (size_t)i * lda * k, (size_t)i * ldb * n,
              (size_t)i * ldc * n);
  auto a_d = make_sycl_iterator_buffer<float>(a.data(), a.size());
  auto b_d = make_sycl_iterator_buffer<float>(b.data(), b.size());
  bool ok = true;
  {
    auto c_d = make_sycl_iterator_buffer<float>(c.size());
    helper::copy_to_device(q, c.data(), c_d, c.size());
    auto ev = _gemm_batched(sb, 'n', 'n', m, n, k, 3.0f, a_d, lda, b_d, ldb, 7.0f, c_d, ldc, batch,
                            gemm_batch_type_t::strided);
    sb.wait(ev);
    std::vector<float> got(c.size());
    sb.wait(helper::copy_to_host(q, c_d, got.data(), got.size()));
    ok &= close(got, want, 2e-5, "_gemm_batched float strided (buffers)");
  }
  {
    float* c_u = sycl::malloc_device<float>(c.size(), q);
    helper::copy_to_device(q, c.data(), c_u, c.size());
    auto ev = _gemm_strided_batched(sb, 'n', 'n', m, n, k, 3.0f, a_d, lda, lda * k, b_d, ldb, ldb * n, 7.0f, c_u,
                                    ldc, ldc * n, batch);
    sb.wait(ev);
    const auto t0 = ev[0].get_profiling_info<sycl::info::event_profiling::command_start>();
    const auto t1 = ev[0].get_profiling_info<sycl::info::event_profiling::command_end>();
    std::printf("  event profiling: %.1f us on device\n", (double)(t1 - t0) / 1e3);
    std::vector<float> got(c.size());
    sb.wait(helper::copy_to_host(q, c_u, got.data(), got.size()));
    sycl::free(c_u, q);
    ok &= close(got, want, 2e-5, "_gemm_strided_batched float (buffer A,B + USM C)") && t1 >= t0;
  }
  return ok;
}

// _symm, _trsm and complex _gemm: the routines built on the GEMM path
bool run_symm_trsm_complex(blas::SB_Handle& sb) {
  using namespace blas;
  auto q = sb.get_queue();
  bool ok = true;
  {  // _symm left / upper: only the upper triangle of A is referenced
    const int m = 200, n = 130, lda = m + 1, ldb = m, ldc = m + 3;
    std::vector<float> a((size_t)lda * m), b((size_t)ldb * n), c((size_t)ldc * n), want;
    fill(a, 11); fill(b, 12); fill(c, 13);
    std::vector<float> full(a);
    for (int j = 0; j < m; ++j)
      for (int i = j + 1; i < m; ++i) { full[i + (size_t)j * lda] = a[j + (size_t)i * lda]; a[i + (size_t)j * lda] = NAN; }
    want = c;
    host_gemm('n', 'n', m, n, m, 1.5, full, lda, b, ldb, 0.5, want, ldc);
    auto a_d = make_sycl_iterator_buffer<float>(a.data(), a.size());
    auto b_d = make_sycl_iterator_buffer<float>(b.data(), b.size());
    float* c_u = sycl::malloc_device<float>(c.size(), q);
    helper::copy_to_device(q, c.data(), c_u, c.size());
    sb.wait(_symm(sb, 'l', 'u', m, n, 1.5f, a_d, lda, b_d, ldb, 0.5f, c_u, ldc));
    std::vector<float> got(c.size());
    sb.wait(helper::copy_to_host(q, c_u, got.data(), got.size()));
    sycl::free(c_u, q);
    ok &= close(got, want, 2e-5, "_symm float left/upper 200x130");
  }
  {  // _trsm left / lower / no-trans / non-unit: check the residual A*X = alpha*B
    const int m = 300, n = 70, lda = m, ldb = m + 2;
    std::vector<double> a((size_t)lda * m, NAN), b((size_t)ldb * n), x;
    std::mt19937 gen(21);
    std::uniform_real_distribution<double> dis(-1.0, 1.0);
    for (int j = 0; j < m; ++j)
      for (int i = j; i < m; ++i) a[i + (size_t)j * lda] = (i == j) ? 4.0 + dis(gen) : dis(gen) / m;
    fill(b, 22);
    double* a_u = sycl::malloc_device<double>(a.size(), q);
    double* b_u = sycl::malloc_device<double>(b.size(), q);
    helper::copy_to_device(q, a.data(), a_u, a.size());
    helper::copy_to_device(q, b.data(), b_u, b.size());
    sb.wait(_trsm(sb, 'l', 'l', 'n', 'n', m, n, 2.0, a_u, lda, b_u, ldb));
    x.resize(b.size());
    sb.wait(helper::copy_to_host(q, b_u, x.data(), x.size()));
    sycl::free(a_u, q); sycl::free(b_u, q);
    double worst = 0;
    for (int j = 0; j < n; ++j)
      for (int i = 0; i < m; ++i) {
        double acc = 0;
        for (int l = 0; l <= i; ++l) acc += a[i + (size_t)l * lda] * x[l + (size_t)j * ldb];
        worst = std::max(worst, std::fabs(acc - 2.0 * b[i + (size_t)j * ldb]) / (std::fabs(2.0 * b[i + (size_t)j * ldb]) + 1.0));
      }
    const bool pass = worst < 1e-12 && x[m] == b[m];   // padding row untouched
    std::printf("%-46s max residual %.3e  %s\n", "_trsm double left/lower 300x70", worst, pass ? "PASS" : "FAIL");
    ok &= pass;
  }
  {  // complex _gemm (BLAS_ENABLE_COMPLEX): C = alpha*A^T*B + beta*C
    using cf = std::complex<float>;
    const int m = 65, n = 40, k = 100, lda = k, ldb = k + 1, ldc = m;
    std::vector<cf> a((size_t)lda * m), b((size_t)ldb * n), c((size_t)ldc * n), want;
    std::mt19937 gen(31);
    std::uniform_real_distribution<float> dis(-2.0f, 5.0f);
    for (auto* v : {&a, &b, &c}) for (auto& e : *v) e = cf(dis(gen), dis(gen));
    want = c;
    const cf alpha(1.5f, 1.0f), beta(1.5f, 3.0f);
    for (int j = 0; j < n; ++j)
      for (int i = 0; i < m; ++i) {
        std::complex<double> acc = 0;
        for (int l = 0; l < k; ++l) acc += std::complex<double>(a[l + (size_t)i * lda]) * std::complex<double>(b[l + (size_t)j * ldb]);
        want[i + (size_t)j * ldc] = cf(std::complex<double>(alpha) * acc + std::complex<double>(beta) * std::complex<double>(c[i + (size_t)j * ldc]));
      }
    cf* a_u = sycl::malloc_device<cf>(a.size(), q);
    cf* b_u = sycl::malloc_device<cf>(b.size(), q);
    cf* c_u = sycl::malloc_device<cf>(c.size(), q);
    helper::copy_to_device(q, a.data(), a_u, a.size());
    helper::copy_to_device(q, b.data(), b_u, b.size());
    helper::copy_to_device(q, c.data(), c_u, c.size());
    sb.wait(_gemm(sb, 't', 'n', m, n, k, alpha, a_u, lda, b_u, ldb, beta, c_u, ldc));
    std::vector<cf> got(c.size());
    sb.wait(helper::copy_to_host(q, c_u, got.data(), got.size()));
    sycl::free(a_u, q); sycl::free(b_u, q); sycl::free(c_u, q);
    double worst = 0;
    for (size_t i = 0; i < got.size(); ++i) worst = std::max(worst, (double)std::abs(got[i] - want[i]) / (std::abs(want[i]) + 1.0));
    std::printf("%-46s max rel err %.3e  %s\n", "_gemm complex<float> tn 65x40x100", worst, worst < 2e-5 ? "PASS" : "FAIL");
    ok &= worst < 2e-5;
  }
  bool threw = false;
  try {
    float* p = sycl::malloc_device<float>(16, q);
    try { _symm(sb, 'l', 'x', 4, 4, 1.0f, p, 4, p, 4, 0.0f, p, 4); } catch (const std::invalid_argument& e) { threw = std::string(e.what()) == "invalid _uplo"; }
    bool t2 = false;
    try { _trsm(sb, 'l', 'u', 'c', 'n', 4, 4, 1.0f, p, 4, p, 4); } catch (const std::invalid_argument& e) { t2 = std::string(e.what()) == "invalid Transpose argument"; }
    threw = threw && t2;
    sycl::free(p, q);
  } catch (...) { threw = false; }
  std::printf("%-46s %s\n", "_symm / _trsm invalid arguments", threw ? "PASS" : "FAIL");
  return ok && threw;
}

bool run_errors(blas::SB_Handle& sb) {
  auto q = sb.get_queue();
  float* p = sycl::malloc_device<float>(64, q);
  bool ok = false;
  try {
    blas::_gemm(sb, 'x', 'n', 4, 4, 4, 1.0f, p, 4, p, 4, 0.0f, p, 4);
  } catch (const std::invalid_argument& e) {
    ok = std::string(e.what()) == "invalid _TransA";
  }
  sycl::free(p, q);
  std::printf("%-46s %s\n", "invalid trans -> std::invalid_argument", ok ? "PASS" : "FAIL");
  return ok;
}

}  // namespace

int main() {
  sycl::queue q(sycl::default_selector_v, {sycl::property::queue::enable_profiling()});
  blas::SB_Handle sb(q);
  std::printf("device: %s, %zu compute units\n",
              q.get_device().get_info<sycl::info::device::name>().c_str(), sb.get_num_compute_units());
  using blas::helper::AllocType;
  bool ok = true;
  ok &= run_gemm<float, float, AllocType::buffer>(sb, 'n', 'n', 7, 5, 9, 2e-5, "_gemm float nn 7x5x9 (buffers)");
  ok &= run_gemm<float, float, AllocType::usm>(sb, 't', 'n', 300, 200, 520, 2e-5, "_gemm float tn 300x200x520 (usm)");
  ok &= run_gemm<float, float, AllocType::usm>(sb, 'n', 't', 1024, 512, 1024, 2e-5, "_gemm float nt 1024x512x1024 (usm)");
  ok &= run_gemm<double, double, AllocType::buffer>(sb, 'n', 't', 253, 257, 511, 1e-12, "_gemm double nt 253x257x511 (buffers)");
  ok &= run_gemm<double, double, AllocType::usm>(sb, 't', 't', 512, 384, 640, 1e-12, "_gemm double tt 512x384x640 (usm)");
  ok &= run_gemm<sycl::half, sycl::half, AllocType::usm>(sb, 'n', 'n', 256, 128, 192, 3e-3, "_gemm half nn 256x128x192 (usm)");
  ok &= run_gemm<sycl::half, float, AllocType::buffer>(sb, 't', 'n', 125, 131, 192, 2e-5, "_gemm half->float tn 125x131x192 (buffers)");
  ok &= run_batched(sb);
  ok &= run_symm_trsm_complex(sb);
  ok &= run_errors(sb);
  std::printf("%s\n", ok ? "ALL PASS" : "SOME FAILED");
  return ok ? 0 : 1;
}
