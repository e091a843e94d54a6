/*
 * gemm_oracle.c -- CPU restatement of portBLAS's GEMM path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg
 * may load this.  The product (libpbx_gemm.so) never links or calls it.
 *
 * PARITY STATUS: PINNED against outputs of the reference itself.  The reference (codeplaysoftware/portBLAS @ 6cf5e58)
 * is SYCL and this image has no SYCL compiler, but its GEMM path is header-only: `make ref` compiles it unchanged from
 * /root/reference over a host stand-in for the SYCL runtime (ref_host_driver.cpp, sycl_host/sycl/sycl.hpp) into
 * oracle/_ref/, and tests/test_oracle_ref.py checks that the "gemm_local ordering" below and the DEFAULT-backend kernel
 * reproduce it BIT FOR BIT on the reference's grids (single, strided-batched, interleaved; default and NVIDIA backend
 * heuristics), the front-end rules included; committed outputs of it (tests/golden/ref_host_golden.npz) keep the pin
 * where oracle/_ref is absent.  The reference ships no golden vectors of its own (its tests compare against a system
 * CBLAS at run time on random inputs, test/unittest/blas3/blas3_gemm_common.hpp:164-169,223-225), so this file is also
 * pinned (tests/test_oracle.py) against that CBLAS -- OpenBLAS 0.3.30 via numpy -- on the reference's parameter grids,
 * input distribution U(-2,5) and tolerance function (common/include/common/float_comparison.hpp:163-188).
 *
 * What is restated, with the reference lines each function follows:
 *   oracle_gemm_ref_*     naive kernel            src/operations/blas3/gemm_ref.hpp:204-260
 *                         (acc=0; k ascending mad; C = alpha*acc [+ beta*C]; trans indexing :230-243;
 *                          strided batch pointer arithmetic :217-223,254-259)
 *   oracle_gemm_local_*   production ordering     src/operations/blas3/gemm_local.hpp:182 (beta/alpha),
 *                         :377-409 (acc0 = beta'*C or 0), :752-772 (k ascending mad), :523,534 (x alpha)
 *   oracle_gemm_truth_*   long-double accumulation (the value both of the above approximate)
 *   interleaved layout    src/operations/blas3/gemm_interleaved.hpp:265-271
 *   oracle_gemm_frontend  src/interface/gemm_interface.hpp:105-185 (alpha==0 first -> scal; trans /
 *                         stride validation; 'c'=='t'), blas1_interface.hpp:438-510 (_scal, _scal_matrix)
 *   oracle_gemm_default_cpu_*  the kernel the DEFAULT (CPU) backend picks for M*N >= 524288:
 *                         Tile<4,4,4,4>, no local memory (src/interface/blas3/backend/default.hpp:98-112;
 *                         src/operations/blas3/gemm_no_local_partial_vec.hpp:129 beta/alpha, :502 mul_add,
 *                         :510 x alpha): 16x16 block per work-group, 4x4 register tile per work-item.
 *                         OpenMP over work-groups; this is the timed "port" CPU baseline.
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORACLE_OK 0
#define ORACLE_INVALID_TRANSA 1
#define ORACLE_INVALID_TRANSB 2
#define ORACLE_INVALID_STRIDEC 3
#define ORACLE_INVALID_STRIDEA 4
#define ORACLE_INVALID_STRIDEB 5

static int lower(int c) { return (c >= 'A' && c <= 'Z') ? c + 32 : c; }

int oracle_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* element (r,c) of batch entry b of a stored column-major matrix */
static inline int64_t idx(int interleaved, int64_t r, int64_t c, int64_t ld, int64_t b, int64_t stride,
                          int64_t batch) {
  return interleaved ? (c * ld + r) * batch + b : b * stride + c * ld + r;
}

#define DEFINE_KERNELS(SUFFIX, T, FMA)                                                               \
  /* mode 0: gemm_ref ordering, 1: gemm_local ordering, 2: long double truth */                    \
  void oracle_gemm_core_##SUFFIX(int mode, int ta, int tb, int64_t m, int64_t n, int64_t k, T alpha,  \
                                 const T* A, int64_t lda, int64_t sa, const T* B, int64_t ldb,       \
                                 int64_t sb, T beta, T* C, int64_t ldc, int64_t sc, int64_t batch,   \
                                 int interleaved) {                                                  \
    const int beta0 = (beta == (T)0);                                                                \
    _Pragma("omp parallel for collapse(2) schedule(static)")                                         \
    for (int64_t b = 0; b < batch; ++b)                                                              \
      for (int64_t j = 0; j < n; ++j)                                                                \
        for (int64_t i = 0; i < m; ++i) {                                                            \
          T* c = &C[idx(interleaved, i, j, ldc, b, sc, batch)];                                      \
          if (mode == 2) {                                                                           \
            long double acc = 0.0L;                                                                  \
            for (int64_t l = 0; l < k; ++l) {                                                        \
              const T a = ta ? A[idx(interleaved, l, i, lda, b, sa, batch)]                          \
                             : A[idx(interleaved, i, l, lda, b, sa, batch)];                         \
              const T bb = tb ? B[idx(interleaved, j, l, ldb, b, sb, batch)]                         \
                              : B[idx(interleaved, l, j, ldb, b, sb, batch)];                        \
              acc += (long double)a * (long double)bb;                                               \
            }                                                                                        \
            long double r = (long double)alpha * acc;                                                \
            if (!beta0) r += (long double)beta * (long double)(*c);                                  \
            *c = (T)r;                                                                               \
          } else {                                                                                   \
            T acc = (T)0;                                                                            \
            if (mode == 1 && !beta0) acc = (beta / alpha) * (*c);                                    \
            for (int64_t l = 0; l < k; ++l) {                                                        \
              const T a = ta ? A[idx(interleaved, l, i, lda, b, sa, batch)]                          \
                             : A[idx(interleaved, i, l, lda, b, sa, batch)];                         \
              const T bb = tb ? B[idx(interleaved, j, l, ldb, b, sb, batch)]                         \
                              : B[idx(interleaved, l, j, ldb, b, sb, batch)];                        \
              acc = FMA(a, bb, acc);                                                                 \
            }                                                                                        \
            if (mode == 1) *c = alpha * acc;                                                         \
            else *c = beta0 ? alpha * acc : alpha * acc + beta * (*c);                               \
          }                                                                                          \
        }                                                                                            \
  }                                                                                                  \
                                                                                                     \
  /* _scal_matrix: beta==1 no-op; else C = beta*C (0*C when beta==0 -> NaN propagates);           */ \
  /* _scal (contiguous case): exact zeros when beta==0.   blas1_interface.hpp:444-448,498-505     */ \
  static void scal_window_##SUFFIX(T beta, T* C, int64_t m, int64_t n, int64_t ldc, int exact_zero) { \
    if (!exact_zero && beta == (T)1) return;                                                         \
    for (int64_t j = 0; j < n; ++j)                                                                  \
      for (int64_t i = 0; i < m; ++i)                                                                \
        C[j * ldc + i] = (exact_zero && beta == (T)0) ? (T)0 : beta * C[j * ldc + i];                \
  }                                                                                                  \
                                                                                                     \
  /* Front end.  Returns ORACLE_* status (the reference throws std::invalid_argument).            */ \
  int oracle_gemm_frontend_##SUFFIX(int mode, char transa, char transb, int64_t m, int64_t n,        \
                                    int64_t k, T alpha, const T* A, int64_t lda, int64_t sa,         \
                                    const T* B, int64_t ldb, int64_t sb, T beta, T* C, int64_t ldc,  \
                                    int64_t sc, int64_t batch, int interleaved) {                    \
    if (alpha == (T)0) { /* gemm_interface.hpp:112-139 -- before validation */                       \
      if (interleaved && batch > 1) {                                                                \
        /* mathematically intended result (the reference is untested/ill-defined here) */           \
        scal_window_##SUFFIX(beta, C, m * batch, n, ldc * batch, 0);                                 \
        return ORACLE_OK;                                                                            \
      }                                                                                              \
      const int64_t size_c = ldc * n;                                                                \
      if (size_c == sc) {                                                                            \
        if (ldc == m) scal_window_##SUFFIX(beta, C, size_c * batch, 1, size_c * batch, 1);           \
        else scal_window_##SUFFIX(beta, C, m, n * batch, ldc, 0);                                    \
      } else {                                                                                       \
        for (int64_t b = 0; b < batch; ++b) scal_window_##SUFFIX(beta, C + b * sc, m, n, ldc, 0);    \
      }                                                                                              \
      return ORACLE_OK;                                                                              \
    }                                                                                                \
    const int tac = lower(transa), tbc = lower(transb);                                              \
    if (tac != 'n' && tac != 't' && tac != 'c') return ORACLE_INVALID_TRANSA;                        \
    if (tbc != 'n' && tbc != 't' && tbc != 'c') return ORACLE_INVALID_TRANSB;                        \
    if (batch > 1 && !interleaved) {                                                                 \
      if (sc < ldc * n || sc < 0) return ORACLE_INVALID_STRIDEC;                                     \
      if (sa < 0) return ORACLE_INVALID_STRIDEA;                                                     \
      if (sb < 0) return ORACLE_INVALID_STRIDEB;                                                     \
    }                                                                                                \
    oracle_gemm_core_##SUFFIX(mode, tac != 'n', tbc != 'n', m, n, k, alpha, A, lda, sa, B, ldb, sb,  \
                              beta, C, ldc, sc, batch, interleaved);                                 \
    return ORACLE_OK;                                                                                \
  }                                                                                                  \
                                                                                                     \
  /* DEFAULT-backend CPU kernel: Tile<4,4,4,4> no-local; one 16x16 block per work-group.          */ \
  void oracle_gemm_default_cpu_##SUFFIX(int ta, int tb, int64_t m, int64_t n, int64_t k, T alpha,    \
                                        const T* A, int64_t lda, const T* B, int64_t ldb, T beta,    \
                                        T* C, int64_t ldc) {                                         \
    const int64_t bm = (m + 15) / 16, bn = (n + 15) / 16;                                            \
    const int beta0 = (beta == (T)0);                                                                \
    const T betap = beta0 ? (T)0 : beta / alpha;                                                     \
    _Pragma("omp parallel for collapse(2) schedule(static)")                                         \
    for (int64_t wj = 0; wj < bn; ++wj)                                                              \
      for (int64_t wi = 0; wi < bm; ++wi)                                                            \
        for (int item = 0; item < 16; ++item) { /* 4x4 work-items, each a 4x4 register tile */       \
          const int64_t r0 = wi * 16 + (item % 4) * 4, c0 = wj * 16 + (item / 4) * 4;                \
          T reg[4][4];                                                                               \
          for (int jj = 0; jj < 4; ++jj)                                                             \
            for (int ii = 0; ii < 4; ++ii)                                                           \
              reg[jj][ii] = (!beta0 && r0 + ii < m && c0 + jj < n)                                   \
                                ? betap * C[(c0 + jj) * ldc + r0 + ii] : (T)0;                       \
          for (int64_t l = 0; l < k; ++l) {                                                          \
            T ra[4], rb[4];                                                                          \
            for (int ii = 0; ii < 4; ++ii)                                                           \
              ra[ii] = (r0 + ii < m) ? (ta ? A[(r0 + ii) * lda + l] : A[l * lda + r0 + ii]) : (T)0;  \
            for (int jj = 0; jj < 4; ++jj)                                                           \
              rb[jj] = (c0 + jj < n) ? (tb ? B[l * ldb + c0 + jj] : B[(c0 + jj) * ldb + l]) : (T)0;  \
            for (int jj = 0; jj < 4; ++jj)                                                           \
              for (int ii = 0; ii < 4; ++ii) reg[jj][ii] = FMA(ra[ii], rb[jj], reg[jj][ii]);         \
          }                                                                                          \
          for (int jj = 0; jj < 4; ++jj)                                                             \
            for (int ii = 0; ii < 4; ++ii)                                                           \
              if (r0 + ii < m && c0 + jj < n) C[(c0 + jj) * ldc + r0 + ii] = alpha * reg[jj][ii];    \
        }                                                                                            \
  }

DEFINE_KERNELS(f32, float, fmaf)
DEFINE_KERNELS(f64, double, fma)

/* The reference's comparison predicate, float_comparison.hpp:163-188.  Returns the number of
 * mismatching elements (0 == compare_vectors passes).  kind: 0 float, 1 double, 2 half margins.  */
int64_t oracle_compare_f64(const double* a, const double* b, int64_t count, int kind, int margin_mul) {
  double rel, absm;
  if (kind == 1) { rel = 1e-10; absm = 1e-10; }
  else if (kind == 2) { rel = 0.05; absm = 1.0; }
  else { rel = 0.005 * margin_mul; absm = 0.001 * margin_mul; }
  int64_t bad = 0;
  for (int64_t i = 0; i < count; ++i) {
    const double x = a[i], y = b[i];
    if (x == y) continue;
    if ((isnan(x) && isnan(y)) || (isinf(x) && isinf(y))) continue;
    const double d = fabs(x - y);
    if (x == 0.0 || y == 0.0 || d < absm) { if (!(d < absm)) ++bad; continue; }
    if (!(d / (fabs(x) + fabs(y)) < rel)) ++bad;
  }
  return bad;
}
