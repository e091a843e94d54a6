// blas3_ext.cu -- the BLAS-3 routines the reference builds on top of its GEMM path (SURVEY.md section 8,
// rows f1-f3), re-designed so that all of their O(n^3) work runs through the same tensor-core GEMM
// (pbx_gemm) and everything else is one HBM-bound pass:
//
//   pbx_symm   src/interface/symm_interface.hpp:35-75  (reference: GEMM kernels with a mirroring loader,
//              src/operations/blas3/gemm_local.hpp:813-873)
//              here: symmetrize_kernel mirrors the referenced triangle into a pooled full matrix -> GEMM.
//   pbx_trsm   src/interface/trsm_interface.hpp:105-387 + DiagonalBlocksInverter (src/operations/blas3/trsm.hpp)
//              here: trtri_diag_kernel inverts NB-wide diagonal blocks (NB = 128 fp32 / 64 fp64 instead of 16),
//              then a recursive block substitution whose trailing updates are large GEMMs.
//   pbx_cgemm / pbx_zgemm   BLAS_ENABLE_COMPLEX GEMM (backend/default.hpp:202-246, nvidia_gpu.hpp:237-260)
//              here: [Cr; Ci] = [Ar -Ai; Ai Ar] * [Br; Bi] -- one real GEMM of size 2M x N x 2K between a planar
//              split pass and a combine pass that applies the complex alpha / beta.
#include <ctype.h>
#include <stdio.h>

#include <functional>

#include "pbx_internal.cuh"

namespace {

inline int64_t round_up(int64_t x, int64_t q) { return (x + q - 1) / q * q; }

// ------------------------------------------------------------------------------------------------
// symm: out(r, c) = A(r, c) if (r, c) lies in the referenced triangle else A(c, r).  32x32 tiles through
// shared memory so that both the straight and the mirrored reads are coalesced.
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) symmetrize_kernel(const T* __restrict__ A, T* __restrict__ out, int64_t k,
                                                         int64_t lda, int64_t ldo, int lower) {
  __shared__ T s[32][33];
  const int64_t bi = blockIdx.x, bj = blockIdx.y;   // tile row / column of the output
  // source tile: the one of (bi,bj), (bj,bi) that lies in the referenced triangle
  const bool straight = lower ? (bi >= bj) : (bi <= bj);
  const int64_t si = straight ? bi : bj, sj = straight ? bj : bi;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
  for (int c = ty; c < 32; c += 8) {
    const int64_t gr = si * 32 + tx, gc = sj * 32 + c;
    // on the diagonal tile only the referenced half may be read (the other half may hold anything, e.g. NaN)
    const bool ref = lower ? (gr >= gc) : (gr <= gc);
    if (gr < k && gc < k && ref) s[c][tx] = A[gr + gc * lda];
  }
  __syncthreads();
  for (int c = ty; c < 32; c += 8) {
    const int64_t gr = bi * 32 + tx, gc = bj * 32 + c;
    if (gr < k && gc < k) {
      const bool valid = lower ? (gr >= gc) : (gr <= gc);
      out[gr + gc * ldo] = valid ? s[c][tx] : s[tx][c];
    }
  }
}

// ------------------------------------------------------------------------------------------------
// complex GEMM helpers
// ------------------------------------------------------------------------------------------------
// Planar expansion of an interleaved complex matrix S (rows x cols, ld_s complex elements) into a real matrix:
//   block (0,0) = Sr                 always
//   block (0,1) = tr * Si            if tr != 0     (columns cols .. 2 cols-1)
//   block (1,0) = bl * Si            if bl != 0     (rows rows .. 2 rows-1)
//   block (1,1) = Sr                 if br
template <typename T, typename T2>
__global__ void __launch_bounds__(256) cplx_expand_kernel(const T2* __restrict__ src, T* __restrict__ dst,
                                                          int64_t rows, int64_t cols, int64_t ld_s, int64_t stride_s,
                                                          int64_t ld_d, int64_t stride_d, int64_t copies, T tr, T bl,
                                                          int br) {
  const int64_t per = rows * cols, total = per * copies;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = i / per, e = i % per, r = e % rows, c = e / rows;
    const T2 v = src[b * stride_s + r + c * ld_s];
    T* d = dst + b * stride_d;
    d[r + c * ld_d] = v.x;
    if (tr != T(0)) d[r + (cols + c) * ld_d] = tr * v.y;
    if (bl != T(0)) d[rows + r + c * ld_d] = bl * v.y;
    if (br) d[rows + r + (cols + c) * ld_d] = v.x;
  }
}

// C <- alpha * (Tr + i Ti) + beta * C  on interleaved complex C; T is real, 2m x n (rows 0..m-1 real parts,
// m..2m-1 imaginary parts).  has_t == 0: C <- beta*C (the alpha == 0 / k == 0 shortcut).  beta == 0 never reads C.
template <typename T, typename T2>
__global__ void __launch_bounds__(256) cplx_combine_kernel(const T* __restrict__ t, T2* __restrict__ C, int64_t m,
                                                           int64_t n, int64_t ldt, int64_t stride_t, int64_t ldc,
                                                           int64_t stride_c, int64_t batch, T ar, T ai, T br, T bi,
                                                           int has_t) {
  const int64_t per = m * n, total = per * batch;
  const bool beta0 = (br == T(0) && bi == T(0));
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = i / per, e = i % per, r = e % m, c = e / m;
    T2* dst = C + b * stride_c + r + c * ldc;
    T2 o;
    o.x = T(0); o.y = T(0);
    if (has_t) {
      const T* tp = t + b * stride_t + r + c * ldt;
      const T xr = tp[0], xi = tp[m];
      o.x = ar * xr - ai * xi;
      o.y = ar * xi + ai * xr;
    }
    if (!beta0) {
      const T2 cv = *dst;
      o.x += br * cv.x - bi * cv.y;
      o.y += br * cv.y + bi * cv.x;
    }
    *dst = o;
  }
}

inline unsigned grid_for(pbx_handle_t h, int64_t total, int threads) {
  int64_t blocks = (total + threads - 1) / threads;
  const int64_t cap = (int64_t)h->sm_count * 16;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (unsigned)blocks;
}

template <typename T, typename T2>
int gemm_complex(pbx_handle_t h, int real_dtype, char transa, char transb, int64_t m, int64_t n, int64_t k,
                 const T* alpha, const void* A, int64_t lda, int64_t stridea, const void* B, int64_t ldb,
                 int64_t strideb, const T* beta, void* C, int64_t ldc, int64_t stridec, int64_t batch) {
  if (!h) return PBX_ERR_INVALID_ARG;
  if (!alpha || !beta || m < 0 || n < 0 || k < 0 || batch < 0) {
    h->last_error = "complex gemm: invalid argument";
    return PBX_ERR_INVALID_ARG;
  }
  PBX_DEVICE_GUARD(h);
  const T ar = alpha[0], ai = alpha[1], br = beta[0], bi = beta[1];
  h->last_split_k = 1;
  h->last_repack = 0;
  const int64_t sc = (batch > 1) ? stridec : 0;
  auto scal_only = [&]() -> int {   // C <- beta*C
    if (m == 0 || n == 0 || batch == 0 || (br == T(1) && bi == T(0))) { h->last_kernel = PBX_KERNEL_NONE; return PBX_OK; }
    h->last_kernel = PBX_KERNEL_SCAL;
    cplx_combine_kernel<T, T2><<<grid_for(h, m * n * batch, 256), 256, 0, h->stream>>>(
        nullptr, (T2*)C, m, n, 0, 0, ldc, sc, batch, T(0), T(0), br, bi, 0);
    h->launches++;
    PBX_CUDA_CHECK(h, cudaGetLastError());
    return PBX_OK;
  };
  // (1) alpha == 0 first, before validation (gemm_interface.hpp:112-139 with the complex isZero, :58-66)
  if (ar == T(0) && ai == T(0)) return scal_only();
  // (2) trans and stride validation (gemm_interface.hpp:141-166)
  const int ta_c = tolower((unsigned char)transa), tb_c = tolower((unsigned char)transb);
  if (ta_c != 'n' && ta_c != 't' && ta_c != 'c') return PBX_ERR_INVALID_TRANSA;
  if (tb_c != 'n' && tb_c != 't' && tb_c != 'c') return PBX_ERR_INVALID_TRANSB;
  if (batch > 1) {
    if (stridec < ldc * n || stridec < 0) return PBX_ERR_INVALID_STRIDEC;
    if (stridea < 0) return PBX_ERR_INVALID_STRIDEA;
    if (strideb < 0) return PBX_ERR_INVALID_STRIDEB;
  }
  if (m == 0 || n == 0 || batch == 0) { h->last_kernel = PBX_KERNEL_NONE; return PBX_OK; }
  if (k == 0) return scal_only();
  if (!A || !B || !C) { h->last_error = "complex gemm: null matrix pointer"; return PBX_ERR_INVALID_ARG; }

  const bool ta = ta_c != 'n', tb = tb_c != 'n';
  // sign of the imaginary part under the transpose: 'c' conjugates only when the handle asks for BLAS semantics
  const T sa = (ta_c == 'c' && h->conj_transpose) ? T(-1) : T(1);
  const T sb = (tb_c == 'c' && h->conj_transpose) ? T(-1) : T(1);
  const int64_t q = 16 / (int64_t)sizeof(T);          // keep every planar operand TMA-legal
  const int64_t a_rows = ta ? k : m, a_cols = ta ? m : k;   // stored complex A
  const int64_t b_rows = tb ? n : k, b_cols = tb ? k : n;   // stored complex B
  const int64_t sa_in = (batch > 1) ? stridea : 0, sb_in = (batch > 1) ? strideb : 0;
  const int64_t a_copies = sa_in > 0 ? batch : 1, b_copies = sb_in > 0 ? batch : 1;
  // A~ : [Sr tr*Si; bl*Si Sr]   (2 a_rows x 2 a_cols)
  const int64_t lda2 = round_up(2 * a_rows, q), sa2 = lda2 * 2 * a_cols;
  // B~ : op N -> [Sr; Si] (2K x N)      op T -> [Sr  s*Si] (N x 2K)
  const int64_t ldb2 = round_up(tb ? b_rows : 2 * b_rows, q), sb2 = ldb2 * (tb ? 2 * b_cols : b_cols);
  const int64_t ldt = round_up(2 * m, q), st = ldt * n;
  int s;
  if ((s = pbx_ensure_aux(h, 0, sa2 * a_copies * (int64_t)sizeof(T))) != PBX_OK) return s;
  if ((s = pbx_ensure_aux(h, 1, sb2 * b_copies * (int64_t)sizeof(T))) != PBX_OK) return s;
  if ((s = pbx_ensure_aux(h, 2, st * batch * (int64_t)sizeof(T))) != PBX_OK) return s;
  T* A2 = (T*)h->aux[0];
  T* B2 = (T*)h->aux[1];
  T* T2buf = (T*)h->aux[2];
  // op N: [Ar -Ai; Ai Ar].  op T (stored S = op(A)^T): [Sr s*Si; -s*Si Sr]
  const T a_tr = ta ? sa : T(-1), a_bl = -a_tr;
  cplx_expand_kernel<T, T2><<<grid_for(h, a_rows * a_cols * a_copies, 256), 256, 0, h->stream>>>(
      (const T2*)A, A2, a_rows, a_cols, lda, sa_in, lda2, sa2, a_copies, a_tr, a_bl, 1);
  cplx_expand_kernel<T, T2><<<grid_for(h, b_rows * b_cols * b_copies, 256), 256, 0, h->stream>>>(
      (const T2*)B, B2, b_rows, b_cols, ldb, sb_in, ldb2, sb2, b_copies, tb ? sb : T(0), tb ? T(0) : T(1), 0);
  h->launches += 2;
  PBX_CUDA_CHECK(h, cudaGetLastError());
  const T one = T(1), zero = T(0);
  s = pbx_gemm(h, real_dtype, ta ? 't' : 'n', tb ? 't' : 'n', 2 * m, n, 2 * k, &one, A2, lda2,
               sa_in > 0 ? sa2 : 0, B2, ldb2, sb_in > 0 ? sb2 : 0, &zero, T2buf, ldt, st, batch, 0);
  if (s != PBX_OK) return s;
  cplx_combine_kernel<T, T2><<<grid_for(h, m * n * batch, 256), 256, 0, h->stream>>>(
      T2buf, (T2*)C, m, n, ldt, st, ldc, sc, batch, ar, ai, br, bi, 1);
  h->launches++;
  PBX_CUDA_CHECK(h, cudaGetLastError());
  return PBX_OK;
}

// ------------------------------------------------------------------------------------------------
// trsm: inverses of the NB x NB diagonal blocks of a triangular matrix.
// One CTA per block, one thread per column of the inverse.  The block is read as a LOWER triangle L
// (an upper triangle is read transposed, (U^-1) = ((U^T)^-1)^T), padded with the identity past the edge of A.
// Row-oriented forward substitution: X(i, j) = (delta_ij - sum_{k=j..i-1} L(i,k) X(k,j)) / L(i,i); thread j only
// ever reads its own column of X, so the sweep needs no barrier; L(i,k) is a shared-memory broadcast.
// Output block b: invA + b*NB*NB, leading dimension NB, same triangle as the input, exact zeros elsewhere.
// ------------------------------------------------------------------------------------------------
template <typename T, int NB>
__global__ void __launch_bounds__(NB) trtri_diag_kernel(const T* __restrict__ A, T* __restrict__ invA, int64_t k,
                                                        int64_t lda, int lower, int unit) {
  extern __shared__ __align__(16) unsigned char trtri_smem[];
  T (*L)[NB + 1] = reinterpret_cast<T (*)[NB + 1]>(trtri_smem);
  T (*X)[NB + 1] = reinterpret_cast<T (*)[NB + 1]>(trtri_smem + sizeof(T) * NB * (NB + 1));
  const int64_t b0 = (int64_t)blockIdx.x * NB;
  const int t = threadIdx.x;
  // load: consecutive threads read consecutive memory (rows of a column of A)
  for (int c = 0; c < NB; ++c) {
    const int64_t gr = b0 + t, gc = b0 + c;
    // element (t, c) of the stored block lies in the referenced triangle?
    const bool ref = lower ? (t >= c) : (t <= c);
    T v = (t == c) ? T(1) : T(0);
    if (ref && gr < k && gc < k && !(unit && t == c)) v = A[gr + gc * lda];
    if (lower) L[t][c] = v; else L[c][t] = v;   // upper: L = U^T
  }
  __syncthreads();
  const int j = t;
  for (int i = 0; i < NB; ++i) {
    T x = T(0);
    if (i >= j) {
      T s0 = (i == j) ? T(1) : T(0), s1 = T(0);
      int kk = j;
      for (; kk + 1 < i; kk += 2) {
        s0 -= L[i][kk] * X[kk][j];
        s1 -= L[i][kk + 1] * X[kk + 1][j];
      }
      if (kk < i) s0 -= L[i][kk] * X[kk][j];
      x = (s0 + s1) / L[i][i];
    }
    X[i][j] = x;
  }
  __syncthreads();
  T* out = invA + (int64_t)blockIdx.x * NB * NB;
  for (int c = 0; c < NB; ++c) out[t + c * NB] = lower ? X[t][c] : X[c][t];
}

template <typename T, int NB>
int trsm_impl(pbx_handle_t h, int dtype, bool left, bool lower, bool trans, bool unit, int64_t m, int64_t n, T alpha,
              const T* A, int64_t lda, T* B, int64_t ldb) {
  const int64_t K = left ? m : n;
  const int64_t nblk = (K + NB - 1) / NB;
  const int64_t q = 16 / (int64_t)sizeof(T);
  const int64_t ldx = round_up(m, q);
  int s;
  if ((s = pbx_ensure_aux(h, 0, nblk * NB * NB * (int64_t)sizeof(T))) != PBX_OK) return s;
  if ((s = pbx_ensure_aux(h, 1, ldx * n * (int64_t)sizeof(T))) != PBX_OK) return s;
  T* invA = (T*)h->aux[0];
  T* X = (T*)h->aux[1];
  constexpr int SMEM = 2 * (int)sizeof(T) * NB * (NB + 1);
  auto kern = trtri_diag_kernel<T, NB>;
  PBX_CUDA_CHECK(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
  kern<<<(unsigned)nblk, NB, SMEM, h->stream>>>(A, invA, K, lda, lower ? 1 : 0, unit ? 1 : 0);
  h->launches++;
  PBX_CUDA_CHECK(h, cudaGetLastError());

  const char tc = trans ? 't' : 'n';
  const T one = T(1), zero = T(0), minus = T(-1);
  const bool op_lower = (lower != trans);
  // left: blocks are solved top-down when op(A) is lower; right: left-to-right when op(A) is upper
  const bool forward = left ? op_lower : !op_lower;
  auto start = [&](int64_t blk) { return blk * NB; };
  auto stop = [&](int64_t blk) { return blk * NB < K ? blk * NB : K; };   // exclusive end of block range [.., blk)
  // op(A)[rows of range P, cols of range Q] as a GEMM operand pointer (trans applied by the GEMM)
  auto a_sub = [&](int64_t p0, int64_t q0) { return trans ? A + q0 + p0 * lda : A + p0 + q0 * lda; };

  // solve blocks [b0, b1); `scaled` = alpha has already been applied to B over this range
  std::function<int(int64_t, int64_t, bool)> rec = [&](int64_t b0, int64_t b1, bool scaled) -> int {
    if (b1 - b0 == 1) {
      const int64_t i0 = start(b0), bs = ((b0 + 1) * NB <= K ? NB : K - i0);
      const T a = scaled ? one : alpha;
      const T* inv = invA + b0 * NB * NB;
      if (left)   // X_i = a * op(invA_ii) * B_i
        return pbx_gemm(h, dtype, tc, 'n', bs, n, bs, &a, inv, NB, 0, B + i0, ldb, 0, &zero, X + i0, ldx, 0, 1, 0);
      // X_j = a * B_j * op(invA_jj)
      return pbx_gemm(h, dtype, 'n', tc, m, bs, bs, &a, B + i0 * ldb, ldb, 0, inv, NB, 0, &zero, X + i0 * ldx, ldx, 0,
                      1, 0);
    }
    const int64_t mid = b0 + (b1 - b0 + 1) / 2;
    // P = the half solved first, Q = the half updated with P's solution
    const int64_t p0 = forward ? b0 : mid, p1 = forward ? mid : b1;
    const int64_t q0 = forward ? mid : b0, q1 = forward ? b1 : mid;
    int st = rec(p0, p1, scaled);
    if (st != PBX_OK) return st;
    const int64_t pi = start(p0), plen = stop(p1) - pi;
    const int64_t qi = start(q0), qlen = stop(q1) - qi;
    const T be = scaled ? one : alpha;
    if (left)   // B_Q <- -op(A)[Q, P] * X_P + be * B_Q
      st = pbx_gemm(h, dtype, tc, 'n', qlen, n, plen, &minus, a_sub(qi, pi), lda, 0, X + pi, ldx, 0, &be, B + qi, ldb,
                    0, 1, 0);
    else        // B_Q <- -X_P * op(A)[P, Q] + be * B_Q
      st = pbx_gemm(h, dtype, 'n', tc, m, qlen, plen, &minus, X + pi * ldx, ldx, 0, a_sub(pi, qi), lda, 0, &be,
                    B + qi * ldb, ldb, 0, 1, 0);
    if (st != PBX_OK) return st;
    return rec(q0, q1, true);
  };
  s = rec(0, nblk, false);
  if (s != PBX_OK) return s;
  // X -> B (the reference's final _copy, trsm_interface.hpp:378-384)
  return pbx_launch_repack(h, (int)sizeof(T), X, B, m, n, ldx, ldb, 0, 0, 1);
}

}  // namespace

extern "C" {

int pbx_set_conj_transpose(pbx_handle_t h, int enable) {
  if (!h) return PBX_ERR_INVALID_ARG;
  h->conj_transpose = enable ? 1 : 0;
  return PBX_OK;
}

int pbx_symm(pbx_handle_t h, int dtype, char side, char uplo, int64_t m, int64_t n, const void* alpha, const void* A,
             int64_t lda, const void* B, int64_t ldb, const void* beta, void* C, int64_t ldc) {
  if (!h) return PBX_ERR_INVALID_ARG;
  if ((dtype != PBX_F32 && dtype != PBX_F64 && dtype != PBX_F16 && dtype != PBX_BF16) || !alpha || !beta || m < 0 ||
      n < 0) {
    h->last_error = "pbx_symm: invalid argument";
    return PBX_ERR_INVALID_ARG;
  }
  const int sd = tolower((unsigned char)side), ul = tolower((unsigned char)uplo);
  if (ul != 'u' && ul != 'l') return PBX_ERR_INVALID_UPLO;   // symm_interface.hpp:51-53 (checked before side)
  if (sd != 'l' && sd != 'r') return PBX_ERR_INVALID_SIDE;   // symm_interface.hpp:70-72
  PBX_DEVICE_GUARD(h);
  const double al = (dtype == PBX_F64) ? *(const double*)alpha : (double)*(const float*)alpha;
  const int64_t k = (sd == 'l') ? m : n;
  const void* Afull = A;
  int64_t ldf = lda;
  if (al != 0.0 && m > 0 && n > 0) {
    if (!A || !B || !C) { h->last_error = "pbx_symm: null matrix pointer"; return PBX_ERR_INVALID_ARG; }
    const int64_t es = (int64_t)pbx_in_size(dtype);
    ldf = round_up(k, 16 / es);
    int s = pbx_ensure_aux(h, 0, ldf * k * es);
    if (s != PBX_OK) return s;
    const dim3 grid((unsigned)((k + 31) / 32), (unsigned)((k + 31) / 32));
    const int lower = (ul == 'l');
    if (es == 2) symmetrize_kernel<uint16_t><<<grid, 256, 0, h->stream>>>((const uint16_t*)A, (uint16_t*)h->aux[0], k, lda, ldf, lower);
    else if (es == 4) symmetrize_kernel<uint32_t><<<grid, 256, 0, h->stream>>>((const uint32_t*)A, (uint32_t*)h->aux[0], k, lda, ldf, lower);
    else symmetrize_kernel<uint64_t><<<grid, 256, 0, h->stream>>>((const uint64_t*)A, (uint64_t*)h->aux[0], k, lda, ldf, lower);
    h->launches++;
    PBX_CUDA_CHECK(h, cudaGetLastError());
    Afull = h->aux[0];
  }
  if (sd == 'l') return pbx_gemm(h, dtype, 'n', 'n', m, n, m, alpha, Afull, ldf, 0, B, ldb, 0, beta, C, ldc, 0, 1, 0);
  return pbx_gemm(h, dtype, 'n', 'n', m, n, n, alpha, B, ldb, 0, Afull, ldf, 0, beta, C, ldc, 0, 1, 0);
}

int pbx_trsm(pbx_handle_t h, int dtype, char side, char uplo, char trans, char diag, int64_t m, int64_t n,
             const void* alpha, const void* A, int64_t lda, void* B, int64_t ldb) {
  if (!h) return PBX_ERR_INVALID_ARG;
  if ((dtype != PBX_F32 && dtype != PBX_F64) || !alpha || m < 0 || n < 0) {
    h->last_error = "pbx_trsm: invalid argument";
    return PBX_ERR_INVALID_ARG;
  }
  if (m == 0 || n == 0 || lda == 0 || ldb == 0) return PBX_ERR_TRSM_SIZE;   // trsm_interface.hpp:112-114
  const int sd = tolower((unsigned char)side), ul = tolower((unsigned char)uplo);
  const int tr = tolower((unsigned char)trans), dg = tolower((unsigned char)diag);
  if (sd != 'l' && sd != 'r') return PBX_ERR_TRSM_SIDE;
  if (ul != 'u' && ul != 'l') return PBX_ERR_TRSM_UPLO;
  if (tr != 'n' && tr != 't') return PBX_ERR_TRSM_TRANS;
  if (dg != 'u' && dg != 'n') return PBX_ERR_TRSM_DIAG;
  if (!A || !B) { h->last_error = "pbx_trsm: null matrix pointer"; return PBX_ERR_INVALID_ARG; }
  PBX_DEVICE_GUARD(h);
  if (dtype == PBX_F64)
    return trsm_impl<double, 64>(h, dtype, sd == 'l', ul == 'l', tr == 't', dg == 'u', m, n, *(const double*)alpha,
                                 (const double*)A, lda, (double*)B, ldb);
  return trsm_impl<float, 128>(h, dtype, sd == 'l', ul == 'l', tr == 't', dg == 'u', m, n, *(const float*)alpha,
                               (const float*)A, lda, (float*)B, ldb);
}

int pbx_cgemm(pbx_handle_t h, char transa, char transb, int64_t m, int64_t n, int64_t k, const float* alpha,
              const void* A, int64_t lda, int64_t stridea, const void* B, int64_t ldb, int64_t strideb,
              const float* beta, void* C, int64_t ldc, int64_t stridec, int64_t batch) {
  return gemm_complex<float, float2>(h, PBX_F32, transa, transb, m, n, k, alpha, A, lda, stridea, B, ldb, strideb,
                                     beta, C, ldc, stridec, batch);
}

int pbx_zgemm(pbx_handle_t h, char transa, char transb, int64_t m, int64_t n, int64_t k, const double* alpha,
              const void* A, int64_t lda, int64_t stridea, const void* B, int64_t ldb, int64_t strideb,
              const double* beta, void* C, int64_t ldc, int64_t stridec, int64_t batch) {
  return gemm_complex<double, double2>(h, PBX_F64, transa, transb, m, n, k, alpha, A, lda, stridea, B, ldb, strideb,
                                       beta, C, ldc, stridec, batch);
}

}  // extern "C"
