// gemm_simt.cu -- CUDA-core kernels of the GEMM path:
//   * gemm_simt_kernel       : shared-memory tiled GEMM for ANY alignment / ld / transpose.
//                              It is the shape-robust member of the family (the role
//                              Gemm<...local...> plays in the reference,
//                              src/operations/blas3/gemm_local.hpp:263-355,427-517) and takes
//                              the calls the TMA/tcgen05 path cannot (unaligned base, odd ld).
//   * gemm_interleaved_kernel: batch-interleaved layout, batch is the fastest dimension
//                              (reference src/operations/blas3/gemm_interleaved.hpp:219-312).
//   * scal_matrix_kernel     : C <- beta*C  (reference blas1_interface.hpp:438-510).
//   * splitk_reduce_kernel   : sum of K-slice partials + alpha/beta epilogue (the role of
//                              Reduction<Add,outer> + the axpby tree in
//                              src/sb_handle/portblas_handle.hpp:354-398).
#include <stdio.h>

#include <stdlib.h>

#include "pbx_internal.cuh"

namespace {

template <typename T> struct Cvt;
template <> struct Cvt<float> {
  __device__ static float to_f(float v) { return v; }
  __device__ static float from_f(float v) { return v; }
};
template <> struct Cvt<double> {
  __device__ static double to_f(double v) { return v; }
  __device__ static double from_f(double v) { return v; }
};
template <> struct Cvt<__half> {
  __device__ static float to_f(__half v) { return __half2float(v); }
  __device__ static __half from_f(float v) { return __float2half_rn(v); }
};
template <> struct Cvt<__nv_bfloat16> {
  __device__ static float to_f(__nv_bfloat16 v) { return __bfloat162float(v); }
  __device__ static __nv_bfloat16 from_f(float v) { return __float2bfloat16_rn(v); }
};

struct SimtParams {
  const void* A;
  const void* B;
  void* C;
  void* ws;  // split-K partials (TAcc)
  int64_t a_rs, a_cs, b_rs, b_cs;  // op(A)(m,k) = A[m*a_rs + k*a_cs], op(B)(k,n) = B[k*b_rs + n*b_cs]
  int64_t m, n, k, ldc, sa, sb, sc, batch;
  int64_t k_per_slice;
  int slices, m_tiles, n_tiles;
  double alpha, beta;
};

constexpr int SBM = 64, SBN = 64, SBK = 16;

template <typename TIn, typename TOut, typename TAcc>
__global__ void __launch_bounds__(256) gemm_simt_kernel(SimtParams p) {
  __shared__ TAcc As[SBK][SBM + 4];
  __shared__ TAcc Bs[SBK][SBN + 4];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int mt = blockIdx.x % p.m_tiles, nt = blockIdx.x / p.m_tiles;
  const int64_t m0 = (int64_t)mt * SBM, n0 = (int64_t)nt * SBN;
  const int slice = blockIdx.y;
  const int64_t kbeg = (int64_t)slice * p.k_per_slice;
  const int64_t kend = min(p.k, kbeg + p.k_per_slice);
  const bool a_mcontig = (p.a_rs == 1);
  const bool b_kcontig = (p.b_rs == 1);

  for (int64_t b = blockIdx.z; b < p.batch; b += gridDim.z) {
    const TIn* A = reinterpret_cast<const TIn*>(p.A) + b * p.sa;
    const TIn* B = reinterpret_cast<const TIn*>(p.B) + b * p.sb;
    TAcc acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = TAcc(0);

    // global -> registers for K block k0 (issued one block ahead, so the loads fly under the FMAs)
    TAcc ra[4], rb[4];
    auto gload = [&](int64_t k0) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int e = tid + i * 256;
        int mm, kk;
        if (a_mcontig) { mm = e % SBM; kk = e / SBM; } else { kk = e % SBK; mm = e / SBK; }
        ra[i] = TAcc(0);
        if (m0 + mm < p.m && k0 + kk < kend)
          ra[i] = (TAcc)Cvt<TIn>::to_f(A[(m0 + mm) * p.a_rs + (k0 + kk) * p.a_cs]);
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int e = tid + i * 256;
        int nn, kk;
        if (b_kcontig) { kk = e % SBK; nn = e / SBK; } else { nn = e % SBN; kk = e / SBN; }
        rb[i] = TAcc(0);
        if (n0 + nn < p.n && k0 + kk < kend)
          rb[i] = (TAcc)Cvt<TIn>::to_f(B[(k0 + kk) * p.b_rs + (n0 + nn) * p.b_cs]);
      }
    };
    gload(kbeg);
    for (int64_t k0 = kbeg; k0 < kend; k0 += SBK) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int e = tid + i * 256;
        int mm, kk;
        if (a_mcontig) { mm = e % SBM; kk = e / SBM; } else { kk = e % SBK; mm = e / SBK; }
        As[kk][mm] = ra[i];
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int e = tid + i * 256;
        int nn, kk;
        if (b_kcontig) { kk = e % SBK; nn = e / SBK; } else { nn = e % SBN; kk = e / SBN; }
        Bs[kk][nn] = rb[i];
      }
      __syncthreads();
      if (k0 + SBK < kend) gload(k0 + SBK);
#pragma unroll
      for (int kk = 0; kk < SBK; ++kk) {
        TAcc a[4], bb[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) a[i] = As[kk][tx * 4 + i];
#pragma unroll
        for (int j = 0; j < 4; ++j) bb[j] = Bs[kk][ty * 4 + j];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fma(a[i], bb[j], acc[i][j]);
      }
      __syncthreads();
    }

    if (p.slices > 1) {
      TAcc* ws = reinterpret_cast<TAcc*>(p.ws) + ((b * p.slices + slice) * p.n) * p.m;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int64_t nn = n0 + ty * 4 + j;
        if (nn >= p.n) continue;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int64_t mm = m0 + tx * 4 + i;
          if (mm < p.m) ws[nn * p.m + mm] = acc[i][j];
        }
      }
    } else {
      TOut* C = reinterpret_cast<TOut*>(p.C) + b * p.sc;
      const TAcc alpha = (TAcc)p.alpha, beta = (TAcc)p.beta;
      const bool beta0 = (p.beta == 0.0);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int64_t nn = n0 + ty * 4 + j;
        if (nn >= p.n) continue;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int64_t mm = m0 + tx * 4 + i;
          if (mm >= p.m) continue;
          TOut* dst = C + mm + nn * p.ldc;
          TAcc r = alpha * acc[i][j];
          if (!beta0) r += beta * (TAcc)Cvt<TOut>::to_f(*dst);
          *dst = Cvt<TOut>::from_f(r);
        }
      }
    }
  }
}

// ---- interleaved batch ------------------------------------------------------
struct IlvParams {
  const void* A;
  const void* B;
  void* C;
  int64_t a_rs, a_cs, b_rs, b_cs;  // in elements of the *interleaved* matrix (before x batch)
  int64_t m, n, k, ldc, batch;
  int64_t m_groups, n_groups, chunks;   // block tiles along m / n, 32-entry batch chunks
  double alpha, beta;
};

// Every batch entry is an independent small GEMM and the batch index is the contiguous dimension, so one
// LANE = one batch entry: every operand load of a warp is one 128-byte line, no lane ever shares data
// with another.  A block is 32 consecutive batch entries x 8 warps; warp w owns the RM x RN register
// tile (w % 4, w / 4) of a (4 RM) x (2 RN) block tile, so the A rows / B columns the warps of a block
// need overlap and are served by L1 after the first touch: per K step a block fetches 4 RM + 2 RN lines
// from L2 for 8 RM RN warp-FMAs (8x8: 48 lines per 512), instead of RM + RN per RM RN.
// Blocks walk m fastest, then n, then the batch chunk, so neighbouring blocks re-read the same B lines
// out of L2.
template <typename TIn, typename TOut, typename TAcc, int RM, int RN, int UNROLL>
__global__ void __launch_bounds__(256) gemm_interleaved_kernel(IlvParams p) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int64_t per_chunk = p.m_groups * p.n_groups, total = per_chunk * p.chunks;
  for (int64_t blk = blockIdx.x; blk < total; blk += gridDim.x) {
    const int64_t chunk = blk / per_chunk, t = blk % per_chunk;
    const int64_t m0 = ((t % p.m_groups) * 4 + (w & 3)) * RM, n0 = ((t / p.m_groups) * 2 + (w >> 2)) * RN;
    if (m0 >= p.m || n0 >= p.n) continue;   // warp-uniform
    const int64_t b = chunk * 32 + lane;
    const int64_t bc = min(b, p.batch - 1);  // lanes past the batch read entry batch-1 and store nothing
    const TIn* A = reinterpret_cast<const TIn*>(p.A) + bc;
    const TIn* B = reinterpret_cast<const TIn*>(p.B) + bc;
    TAcc acc[RM][RN];
#pragma unroll
    for (int i = 0; i < RM; ++i)
#pragma unroll
      for (int j = 0; j < RN; ++j) acc[i][j] = TAcc(0);
    // rows / columns past the edge are clamped (their results are never stored): every load is
    // unconditional, so UNROLL K steps' worth of loads can be in flight per thread
    int64_t a_off[RM], b_off[RN];
#pragma unroll
    for (int i = 0; i < RM; ++i) a_off[i] = min(m0 + i, p.m - 1) * p.a_rs * p.batch;
#pragma unroll
    for (int j = 0; j < RN; ++j) b_off[j] = min(n0 + j, p.n - 1) * p.b_cs * p.batch;
    const int64_t a_kstep = p.a_cs * p.batch, b_kstep = p.b_rs * p.batch;
    int64_t kk = 0;
    for (; kk + UNROLL <= p.k; kk += UNROLL) {
      TIn av[UNROLL][RM], bv[UNROLL][RN];
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) {
#pragma unroll
        for (int i = 0; i < RM; ++i) av[u][i] = A[a_off[i] + (kk + u) * a_kstep];
#pragma unroll
        for (int j = 0; j < RN; ++j) bv[u][j] = B[b_off[j] + (kk + u) * b_kstep];
      }
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) {
        TAcc a[RM], bb[RN];
#pragma unroll
        for (int i = 0; i < RM; ++i) a[i] = (TAcc)Cvt<TIn>::to_f(av[u][i]);
#pragma unroll
        for (int j = 0; j < RN; ++j) bb[j] = (TAcc)Cvt<TIn>::to_f(bv[u][j]);
#pragma unroll
        for (int i = 0; i < RM; ++i)
#pragma unroll
          for (int j = 0; j < RN; ++j) acc[i][j] = fma(a[i], bb[j], acc[i][j]);
      }
    }
    for (; kk < p.k; ++kk) {
      TAcc a[RM], bb[RN];
#pragma unroll
      for (int i = 0; i < RM; ++i) a[i] = (TAcc)Cvt<TIn>::to_f(A[a_off[i] + kk * a_kstep]);
#pragma unroll
      for (int j = 0; j < RN; ++j) bb[j] = (TAcc)Cvt<TIn>::to_f(B[b_off[j] + kk * b_kstep]);
#pragma unroll
      for (int i = 0; i < RM; ++i)
#pragma unroll
        for (int j = 0; j < RN; ++j) acc[i][j] = fma(a[i], bb[j], acc[i][j]);
    }
    if (b >= p.batch) continue;
    TOut* C = reinterpret_cast<TOut*>(p.C) + b;
    const TAcc alpha = (TAcc)p.alpha, beta = (TAcc)p.beta;
    const bool beta0 = (p.beta == 0.0);
#pragma unroll
    for (int j = 0; j < RN; ++j)
#pragma unroll
      for (int i = 0; i < RM; ++i) {
        if (m0 + i >= p.m || n0 + j >= p.n) continue;
        TOut* dst = C + ((n0 + j) * p.ldc + (m0 + i)) * p.batch;
        TAcc r = alpha * acc[i][j];
        if (!beta0) r += beta * (TAcc)Cvt<TOut>::to_f(*dst);
        *dst = Cvt<TOut>::from_f(r);
      }
  }
}

// fp32 interleaved batches, shared-memory staged.  Same mapping as gemm_interleaved_kernel (lane = batch entry, warp
// (w % 4, w / 4) owns an RM x RN register tile of a (4 RM) x (2 RN) block tile), but the 4 RM + 2 RN operand lines a block
// needs per k are fetched ONCE by cp.async into a STAGES-deep ring (KB k-steps per stage) and read back conflict-free
// (a warp reads 32 consecutive floats), so L2 sees every line once per block instead of once per warp and the loads of
// later stages fly under the FMAs.  Out-of-range rows / columns / lanes are clamped (never stored), the K tail is
// zero-filled through the cp.async source size.
template <int RM, int RN, int KB, int STAGES>
__global__ void __launch_bounds__(256, 2) gemm_interleaved_f32_smem_kernel(IlvParams p) {
  constexpr int TM = 4 * RM, TN = 2 * RN, ROWS = TM + TN, RPW = ROWS / 8;
  static_assert(ROWS % 8 == 0, "rows are dealt to the 8 warps");
  extern __shared__ __align__(16) float ilv_smem[];   // [STAGES][KB][ROWS][32]
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int wm = w & 3, wn = w >> 2;
  const float* Ag = reinterpret_cast<const float*>(p.A);
  const float* Bg = reinterpret_cast<const float*>(p.B);
  const int64_t per_chunk = p.m_groups * p.n_groups, total = per_chunk * p.chunks;
  const int nk = (int)((p.k + KB - 1) / KB);
  const int64_t a_kstep = p.a_cs * p.batch, b_kstep = p.b_rs * p.batch;
  const uint32_t smem0 = (uint32_t)__cvta_generic_to_shared(ilv_smem);
  for (int64_t blk = blockIdx.x; blk < total; blk += gridDim.x) {
    const int64_t chunk = blk / per_chunk, t = blk % per_chunk;
    const int64_t m_base = (t % p.m_groups) * TM, n_base = (t / p.m_groups) * TN;
    const int64_t b = chunk * 32 + lane;
    const int64_t bc = min(b, p.batch - 1);
    // the RPW operand rows this warp fetches: block row r = w + 8 j is an A row (r < TM) or a B column
    const float* rbase[RPW];
    int64_t rstep[RPW];
#pragma unroll
    for (int j = 0; j < RPW; ++j) {
      const int r = w + 8 * j;
      if (r < TM) {
        rbase[j] = Ag + min(m_base + r, p.m - 1) * p.a_rs * p.batch + bc;
        rstep[j] = a_kstep;
      } else {
        rbase[j] = Bg + min(n_base + (r - TM), p.n - 1) * p.b_cs * p.batch + bc;
        rstep[j] = b_kstep;
      }
    }
    auto issue = [&](int ks, int stage) {
      const uint32_t sdst = smem0 + (uint32_t)(stage * KB * ROWS * 32 + lane) * 4u;
#pragma unroll
      for (int kk = 0; kk < KB; ++kk) {
        const int64_t kg = (int64_t)ks * KB + kk;
        const uint32_t nbytes = kg < p.k ? 4u : 0u;        // K tail: zero fill
        const int64_t kc = min(kg, p.k - 1);
#pragma unroll
        for (int j = 0; j < RPW; ++j) {
          const uint32_t d = sdst + (uint32_t)((kk * ROWS + w + 8 * j) * 32) * 4u;
          asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(d), "l"(rbase[j] + kc * rstep[j]), "r"(nbytes)
                       : "memory");
        }
      }
    };
    float acc[RM][RN];
#pragma unroll
    for (int i = 0; i < RM; ++i)
#pragma unroll
      for (int j = 0; j < RN; ++j) acc[i][j] = 0.0f;
#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) {
      if (s < nk) issue(s, s);
      asm volatile("cp.async.commit_group;" ::: "memory");
    }
    for (int ks = 0; ks < nk; ++ks) {
      asm volatile("cp.async.wait_group %0;" ::"n"(STAGES - 2) : "memory");
      __syncthreads();   // stage ks has landed for every warp; stage ks-1 is no longer being read
      const int nxt = ks + STAGES - 1;
      if (nxt < nk) issue(nxt, nxt % STAGES);
      asm volatile("cp.async.commit_group;" ::: "memory");
      const float* st = ilv_smem + (size_t)(ks % STAGES) * KB * ROWS * 32 + lane;
#pragma unroll
      for (int kk = 0; kk < KB; ++kk) {
        float a[RM], bb[RN];
#pragma unroll
        for (int i = 0; i < RM; ++i) a[i] = st[(kk * ROWS + wm * RM + i) * 32];
#pragma unroll
        for (int j = 0; j < RN; ++j) bb[j] = st[(kk * ROWS + TM + wn * RN + j) * 32];
#pragma unroll
        for (int i = 0; i < RM; ++i)
#pragma unroll
          for (int j = 0; j < RN; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
      }
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();   // the ring is reused by the next block tile
    const int64_t m0 = m_base + wm * RM, n0 = n_base + wn * RN;
    if (b >= p.batch || m0 >= p.m || n0 >= p.n) continue;
    float* C = reinterpret_cast<float*>(p.C) + b;
    const float alpha = (float)p.alpha, beta = (float)p.beta;
    const bool beta0 = (p.beta == 0.0);
#pragma unroll
    for (int j = 0; j < RN; ++j)
#pragma unroll
      for (int i = 0; i < RM; ++i) {
        if (m0 + i >= p.m || n0 + j >= p.n) continue;
        float* dst = C + ((n0 + j) * p.ldc + (m0 + i)) * p.batch;
        float r = alpha * acc[i][j];
        if (!beta0) r += beta * *dst;
        *dst = r;
      }
  }
}

// ---- operand re-layout -----------------------------------------------------------
// rows x cols column-major window (x batch) copied to a new leading dimension / batch stride, so
// that an operand with an odd ld or a base that is not 16-byte aligned becomes TMA-legal.  One warp
// per (column, 2048-row chunk): every load and store instruction covers 32 consecutive elements.
template <typename T>
__global__ void __launch_bounds__(256) repack_kernel(const T* __restrict__ src, T* __restrict__ dst, int64_t rows,
                                                     int64_t cols, int64_t ld_s, int64_t ld_d, int64_t stride_s,
                                                     int64_t stride_d, int64_t batch, int64_t chunks) {
  const int lane = threadIdx.x & 31;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int64_t units = batch * cols * chunks;
  for (int64_t u = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; u < units; u += nwarps) {
    const int64_t chunk = u % chunks, col = (u / chunks) % cols, b = u / (chunks * cols);
    const int64_t r0 = chunk * 2048, r1 = min(rows, r0 + 2048);
    const T* s = src + b * stride_s + col * ld_s;
    T* d = dst + b * stride_d + col * ld_d;
    int64_t r = r0 + lane;
    for (; r + 96 < r1; r += 128) {
      const T v0 = s[r], v1 = s[r + 32], v2 = s[r + 64], v3 = s[r + 96];
      d[r] = v0; d[r + 32] = v1; d[r + 64] = v2; d[r + 96] = v3;
    }
    for (; r < r1; r += 32) d[r] = s[r];
  }
}

// lo half of the 3xTF32 split, computed once per operand: lo = rn_tf32(x - trunc_tf32(x)) (the tensor core reads
// the raw fp32 word as the hi half, i.e. truncates).  Same window walk as repack_kernel, same ld / stride as the
// source so that the lo copy shares the operand's tensor-map geometry.
__global__ void __launch_bounds__(256) tf32_lo_kernel(const float* __restrict__ src, float* __restrict__ dst,
                                                      int64_t rows, int64_t cols, int64_t ld, int64_t stride,
                                                      int64_t batch, int64_t chunks) {
  const int lane = threadIdx.x & 31;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int64_t units = batch * cols * chunks;
  auto lo_of = [](float x) {
    const float d = x - __uint_as_float(__float_as_uint(x) & 0xffffe000u);
    uint32_t lb;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(lb) : "f"(d == d ? d : 0.0f));   // inf - inf -> 0
    return __uint_as_float(lb);
  };
  for (int64_t u = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; u < units; u += nwarps) {
    const int64_t chunk = u % chunks, col = (u / chunks) % cols, b = u / (chunks * cols);
    const int64_t r0 = chunk * 2048, r1 = min(rows, r0 + 2048);
    const float* s = src + b * stride + col * ld;
    float* d = dst + b * stride + col * ld;
    int64_t r = r0 + lane;
    for (; r + 96 < r1; r += 128) {
      const float v0 = s[r], v1 = s[r + 32], v2 = s[r + 64], v3 = s[r + 96];
      d[r] = lo_of(v0); d[r + 32] = lo_of(v1); d[r + 64] = lo_of(v2); d[r + 96] = lo_of(v3);
    }
    for (; r < r1; r += 32) d[r] = lo_of(s[r]);
  }
}

// ---- fp32 operand -> two bf16 copies for the tf32 + 2 x bf16 product (EXPERIMENTAL, PBX_F32_SPLIT16=1) ----------
// hi16 = bf16(a), lo16 = bf16(a - trunc_tf32(a)); the copies have their own leading dimension / batch stride (multiples
// of 8 elements, so that they are TMA-legal whatever the source's were).  One warp per 2048-row chunk of a column.
// VEC8: eight consecutive rows per lane -- two 16-byte loads, one 16-byte store per copy (needs a 16-byte-aligned source:
// base, ld and stride multiples of 4 floats; the copies' ld16 / st16 are multiples of 8 by construction).
template <bool VEC8>
__global__ void __launch_bounds__(256) split16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ hi,
                                                      __nv_bfloat16* __restrict__ lo, int64_t rows, int64_t cols,
                                                      int64_t ld, int64_t stride, int64_t ld16, int64_t st16,
                                                      int64_t batch, int64_t chunks) {
  const int lane = threadIdx.x & 31;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int64_t units = batch * cols * chunks;
  auto lo_of = [](float x) {
    const float d = x - __uint_as_float(__float_as_uint(x) & 0xffffe000u);
    return d == d ? d : 0.0f;   // inf - inf -> 0
  };
  for (int64_t u = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; u < units; u += nwarps) {
    const int64_t chunk = u % chunks, col = (u / chunks) % cols, b = u / (chunks * cols);
    const int64_t r0 = chunk * 2048, r1 = min(rows, r0 + 2048);
    const float* s = src + b * stride + col * ld;
    __nv_bfloat16* dh = hi + b * st16 + col * ld16;
    __nv_bfloat16* dl = lo + b * st16 + col * ld16;
    if (VEC8) {
      int64_t r = r0 + lane * 8;
      for (; r + 8 <= r1; r += 256) {
        const float4 x0 = *reinterpret_cast<const float4*>(s + r), x1 = *reinterpret_cast<const float4*>(s + r + 4);
        __nv_bfloat162 h[4], l[4];
        h[0] = __floats2bfloat162_rn(x0.x, x0.y); h[1] = __floats2bfloat162_rn(x0.z, x0.w);
        h[2] = __floats2bfloat162_rn(x1.x, x1.y); h[3] = __floats2bfloat162_rn(x1.z, x1.w);
        l[0] = __floats2bfloat162_rn(lo_of(x0.x), lo_of(x0.y)); l[1] = __floats2bfloat162_rn(lo_of(x0.z), lo_of(x0.w));
        l[2] = __floats2bfloat162_rn(lo_of(x1.x), lo_of(x1.y)); l[3] = __floats2bfloat162_rn(lo_of(x1.z), lo_of(x1.w));
        *reinterpret_cast<uint4*>(dh + r) = *reinterpret_cast<const uint4*>(h);
        *reinterpret_cast<uint4*>(dl + r) = *reinterpret_cast<const uint4*>(l);
      }
      for (int64_t t = r; t < min(r + 8, r1); ++t) {   // ragged end of the column (at most one lane)
        const float x = s[t];
        dh[t] = __float2bfloat16_rn(x);
        dl[t] = __float2bfloat16_rn(lo_of(x));
      }
    } else {
      for (int64_t r = r0 + lane; r < r1; r += 32) {
        const float x = s[r];
        dh[r] = __float2bfloat16_rn(x);
        dl[r] = __float2bfloat16_rn(lo_of(x));
      }
    }
  }
}

// ---- C <- beta*C ---------------------------------------------------------------
template <typename TOut, typename TAcc>
__global__ void __launch_bounds__(256) scal_matrix_kernel(TOut* C, int64_t m, int64_t n, int64_t ldc,
                                                          int64_t sc, int64_t batch, double beta_d) {
  const TAcc beta = (TAcc)beta_d;
  const bool beta0 = (beta_d == 0.0);
  const int64_t per = m * n;
  const int64_t total = per * batch;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = i / per, r = i % per;
    TOut* dst = C + b * sc + (r % m) + (r / m) * ldc;
    // beta == 0 stores exact zeros without reading C (blas1_interface.hpp:444-448)
    *dst = beta0 ? Cvt<TOut>::from_f(TAcc(0)) : Cvt<TOut>::from_f(beta * (TAcc)Cvt<TOut>::to_f(*dst));
  }
}

// ---- split-K epilogue ------------------------------------------------------------
// ws holds [batch][slice][n][m] partial sums.  VEC consecutive rows per thread (VEC = 4 needs m % 4 == 0, so a
// vector never crosses a column; the workspace base is 256-byte aligned); the slice loop is unrolled by four so
// that four independent loads are in flight, the sum itself keeps the fixed slice order (deterministic).
template <typename TOut, typename TAcc, int VEC>
__global__ void __launch_bounds__(256) splitk_reduce_kernel(const TAcc* __restrict__ ws, TOut* C,
                                                            int64_t m, int64_t n, int64_t ldc,
                                                            int64_t sc, int64_t batch, int slices,
                                                            double alpha_d, double beta_d) {
  const TAcc alpha = (TAcc)alpha_d, beta = (TAcc)beta_d;
  const bool beta0 = (beta_d == 0.0);
  const int64_t per = m * n;
  const int64_t total = per * batch / VEC;
  struct alignas(sizeof(TAcc) * VEC) Vec { TAcc v[VEC]; };
  // launched with programmatic stream serialization: the grid is resident while the GEMM that writes the partials
  // drains; nothing of the workspace is read before that GEMM has completed
  asm volatile("griddepcontrol.wait;" ::: "memory");
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t e = i * VEC;
    const int64_t b = e / per, r = e % per;
    const TAcc* src = ws + b * slices * per + r;
    TAcc s[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) s[v] = TAcc(0);
    int sl = 0;
    for (; sl + 4 <= slices; sl += 4) {
      Vec x[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) x[u] = *reinterpret_cast<const Vec*>(src + (int64_t)(sl + u) * per);
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int v = 0; v < VEC; ++v) s[v] += x[u].v[v];
    }
    for (; sl < slices; ++sl) {
      const Vec x = *reinterpret_cast<const Vec*>(src + (int64_t)sl * per);
#pragma unroll
      for (int v = 0; v < VEC; ++v) s[v] += x.v[v];
    }
    TOut* dst = C + b * sc + (r % m) + (r / m) * ldc;
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
      TAcc o = alpha * s[v];
      if (!beta0) o += beta * (TAcc)Cvt<TOut>::to_f(dst[v]);
      dst[v] = Cvt<TOut>::from_f(o);
    }
  }
}

template <typename F>
int dispatch_dtype(int dtype, F&& f) {
  switch (dtype) {
    case PBX_F32: return f((float*)0, (float*)0, (float*)0);
    case PBX_F64: return f((double*)0, (double*)0, (double*)0);
    case PBX_F16: return f((__half*)0, (__half*)0, (float*)0);
    case PBX_F16_F32: return f((__half*)0, (float*)0, (float*)0);
    case PBX_BF16: return f((__nv_bfloat16*)0, (__nv_bfloat16*)0, (float*)0);
    case PBX_BF16_F32: return f((__nv_bfloat16*)0, (float*)0, (float*)0);
  }
  return PBX_ERR_INVALID_ARG;
}

}  // namespace

// ---- interleaved <-> strided re-layout of a batch of column-major matrices ---------------------------------------
// interleaved: element (r, c, b) at (c*ld_i + r)*batch + b (reference gemm_interleaved.hpp:265-271: the batch index is
// the contiguous one); strided: at b*stride + c*ld_s + r.  For a fixed column c this is the transpose of a rows x batch
// matrix: 32 x 32 tiles through shared memory, both sides coalesced (128-byte rows of 4-byte elements).
template <typename T, bool TO_STRIDED>
__global__ void __launch_bounds__(256) ilv_relayout_kernel(const T* __restrict__ src, T* __restrict__ dst, int64_t rows,
                                                           int64_t ld_i, int64_t ld_s, int64_t stride, int64_t batch) {
  __shared__ T tile[32][33];
  const int64_t c = blockIdx.y;
  const int64_t r0 = (int64_t)blockIdx.x * 32, b0 = (int64_t)blockIdx.z * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8 threads
  if (TO_STRIDED) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {   // read: b contiguous
      const int64_t r = r0 + ty + 8 * j, b = b0 + tx;
      if (r < rows && b < batch) tile[ty + 8 * j][tx] = src[(c * ld_i + r) * batch + b];
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 4; ++j) {   // write: r contiguous
      const int64_t b = b0 + ty + 8 * j, r = r0 + tx;
      if (r < rows && b < batch) dst[b * stride + c * ld_s + r] = tile[tx][ty + 8 * j];
    }
  } else {
#pragma unroll
    for (int j = 0; j < 4; ++j) {   // read: r contiguous
      const int64_t b = b0 + ty + 8 * j, r = r0 + tx;
      if (r < rows && b < batch) tile[ty + 8 * j][tx] = src[b * stride + c * ld_s + r];
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 4; ++j) {   // write: b contiguous
      const int64_t r = r0 + ty + 8 * j, b = b0 + tx;
      if (r < rows && b < batch) dst[(c * ld_i + r) * batch + b] = tile[tx][ty + 8 * j];
    }
  }
}

int pbx_launch_ilv_relayout(pbx_handle_t h, int elem_bytes, const void* src, void* dst, int64_t rows, int64_t cols,
                            int64_t ld_i, int64_t ld_s, int64_t stride, int64_t batch, bool to_strided) {
  if (rows <= 0 || cols <= 0 || batch <= 0) return PBX_OK;
  if (cols > 65535 || (batch + 31) / 32 > 65535) return PBX_ERR_INVALID_ARG;
  dim3 grid((unsigned)((rows + 31) / 32), (unsigned)cols, (unsigned)((batch + 31) / 32));
  auto go = [&](auto tag) {
    using T = decltype(tag);
    if (to_strided) ilv_relayout_kernel<T, true><<<grid, 256, 0, h->stream>>>((const T*)src, (T*)dst, rows, ld_i, ld_s, stride, batch);
    else ilv_relayout_kernel<T, false><<<grid, 256, 0, h->stream>>>((const T*)src, (T*)dst, rows, ld_i, ld_s, stride, batch);
  };
  if (elem_bytes == 2) go(uint16_t{});
  else if (elem_bytes == 4) go(uint32_t{});
  else if (elem_bytes == 8) go(uint64_t{});
  else return PBX_ERR_INVALID_ARG;
  h->launches++;
  PBX_CUDA_CHECK(h, cudaGetLastError());
  return PBX_OK;
}

int pbx_launch_repack(pbx_handle_t h, int elem_bytes, const void* src, void* dst, int64_t rows, int64_t cols,
                      int64_t ld_src, int64_t ld_dst, int64_t stride_src, int64_t stride_dst, int64_t batch) {
  if (rows <= 0 || cols <= 0 || batch <= 0) return PBX_OK;
  const int64_t chunks = (rows + 2047) / 2048;
  const int64_t units = batch * cols * chunks;
  int64_t blocks = (units + 7) / 8;   // 8 warps per block
  const int64_t cap = (int64_t)h->sm_count * 16;
  if (blocks > cap) blocks = cap;
  if (elem_bytes == 2)
    repack_kernel<uint16_t><<<(unsigned)blocks, 256, 0, h->stream>>>((const uint16_t*)src, (uint16_t*)dst, rows, cols,
                                                                  ld_src, ld_dst, stride_src, stride_dst, batch, chunks);
  else if (elem_bytes == 4)
    repack_kernel<uint32_t><<<(unsigned)blocks, 256, 0, h->stream>>>((const uint32_t*)src, (uint32_t*)dst, rows, cols,
                                                                  ld_src, ld_dst, stride_src, stride_dst, batch, chunks);
  else if (elem_bytes == 8)
    repack_kernel<uint64_t><<<(unsigned)blocks, 256, 0, h->stream>>>((const uint64_t*)src, (uint64_t*)dst, rows, cols,
                                                                  ld_src, ld_dst, stride_src, stride_dst, batch, chunks);
  else
    return PBX_ERR_INVALID_ARG;
  h->launches++;
  PBX_CUDA_CHECK(h, cudaGetLastError());
  return PBX_OK;
}

int pbx_launch_tf32_lo(pbx_handle_t h, const float* src, float* dst, int64_t rows, int64_t cols, int64_t ld,
                       int64_t stride, int64_t batch) {
  if (rows <= 0 || cols <= 0 || batch <= 0) return PBX_OK;
  const int64_t chunks = (rows + 2047) / 2048;
  const int64_t units = batch * cols * chunks;
  int64_t blocks = (units + 7) / 8;   // 8 warps per block
  const int64_t cap = (int64_t)h->sm_count * 16;
  if (blocks > cap) blocks = cap;
  tf32_lo_kernel<<<(unsigned)blocks, 256, 0, h->stream>>>(src, dst, rows, cols, ld, stride, batch, chunks);
  h->launches++;
  PBX_CUDA_CHECK(h, cudaGetLastError());
  return PBX_OK;
}

int pbx_launch_split16(pbx_handle_t h, const float* src, void* hi, void* lo, int64_t rows, int64_t cols, int64_t ld,
                       int64_t stride, int64_t ld16, int64_t st16, int64_t batch) {
  if (rows <= 0 || cols <= 0 || batch <= 0) return PBX_OK;
  const int64_t chunks = (rows + 2047) / 2048;
  const int64_t units = batch * cols * chunks;
  int64_t blocks = (units + 7) / 8;   // 8 warps per block
  const int64_t cap = (int64_t)h->sm_count * 16;
  if (blocks > cap) blocks = cap;
  const bool vec8 = ((uintptr_t)src % 16 == 0) && (ld % 4 == 0) && (stride % 4 == 0);
  if (vec8)
    split16_kernel<true><<<(unsigned)blocks, 256, 0, h->stream>>>(src, (__nv_bfloat16*)hi, (__nv_bfloat16*)lo, rows, cols, ld,
                                                                  stride, ld16, st16, batch, chunks);
  else
    split16_kernel<false><<<(unsigned)blocks, 256, 0, h->stream>>>(src, (__nv_bfloat16*)hi, (__nv_bfloat16*)lo, rows, cols, ld,
                                                                   stride, ld16, st16, batch, chunks);
  h->launches++;
  PBX_CUDA_CHECK(h, cudaGetLastError());
  return PBX_OK;
}

int pbx_launch_simt(pbx_handle_t h, const PbxGemmCall& c, int slices) {
  SimtParams p;
  p.A = c.A; p.B = c.B; p.C = c.C; p.ws = h->ws;
  p.a_rs = c.ta ? c.lda : 1; p.a_cs = c.ta ? 1 : c.lda;
  p.b_rs = c.tb ? c.ldb : 1; p.b_cs = c.tb ? 1 : c.ldb;
  p.m = c.m; p.n = c.n; p.k = c.k; p.ldc = c.ldc;
  p.sa = c.sa; p.sb = c.sb; p.sc = c.sc; p.batch = c.batch;
  p.slices = slices;
  int64_t kps = (c.k + slices - 1) / slices;
  kps = ((kps + SBK - 1) / SBK) * SBK;
  p.k_per_slice = kps;
  p.m_tiles = (int)((c.m + SBM - 1) / SBM);
  p.n_tiles = (int)((c.n + SBN - 1) / SBN);
  p.alpha = c.alpha; p.beta = c.beta;
  dim3 grid((unsigned)(p.m_tiles * (int64_t)p.n_tiles), (unsigned)slices,
            (unsigned)(c.batch < 65535 ? c.batch : 65535));
  return dispatch_dtype(c.dtype, [&](auto* ti, auto* to, auto* ta) -> int {
    using TIn = std::remove_pointer_t<decltype(ti)>;
    using TOut = std::remove_pointer_t<decltype(to)>;
    using TAcc = std::remove_pointer_t<decltype(ta)>;
    gemm_simt_kernel<TIn, TOut, TAcc><<<grid, 256, 0, h->stream>>>(p);
    h->launches++;
    PBX_CUDA_CHECK(h, cudaGetLastError());
    return PBX_OK;
  });
}

int pbx_launch_interleaved(pbx_handle_t h, const PbxGemmCall& c) {
  IlvParams p;
  p.A = c.A; p.B = c.B; p.C = c.C;
  p.a_rs = c.ta ? c.lda : 1; p.a_cs = c.ta ? 1 : c.lda;
  p.b_rs = c.tb ? c.ldb : 1; p.b_cs = c.tb ? 1 : c.ldb;
  p.m = c.m; p.n = c.n; p.k = c.k; p.ldc = c.ldc; p.batch = c.batch;
  p.chunks = (c.batch + 31) / 32;
  p.alpha = c.alpha; p.beta = c.beta;
  // register tile per lane: 8x8 when that still gives every SM a few blocks, else 4x4 (more, smaller blocks);
  // fp64 stays on 4x4 (64 double accumulators would spill)
  auto blocks_for = [&](int rm, int rn) {
    return ((c.m + 4 * rm - 1) / (4 * rm)) * ((c.n + 2 * rn - 1) / (2 * rn)) * p.chunks;
  };
  static const int env_tile = getenv("PBX_ILV_TILE") ? atoi(getenv("PBX_ILV_TILE")) : 0;   // 4 or 8 (testing)
  int tile = (c.dtype != PBX_F64 && blocks_for(8, 8) >= 2 * (int64_t)h->sm_count) ? 8 : 4;
  if (env_tile == 4 || (env_tile == 8 && c.dtype != PBX_F64)) tile = env_tile;
  p.m_groups = (c.m + 4 * tile - 1) / (4 * tile);
  p.n_groups = (c.n + 2 * tile - 1) / (2 * tile);
  int64_t blocks = p.m_groups * p.n_groups * p.chunks;
  const int64_t cap = (int64_t)h->sm_count * 64;
  if (blocks > cap) blocks = cap;
  static const int env_smem = getenv("PBX_ILV_SMEM") ? atoi(getenv("PBX_ILV_SMEM")) : 1;
  if (c.dtype == PBX_F32 && env_smem) {   // shared-memory staged ring (fp32)
    constexpr int KB = 4, STAGES = 4;
    if (tile == 8) {
      constexpr int SMEM = STAGES * KB * (4 * 8 + 2 * 8) * 32 * 4;
      auto kern = gemm_interleaved_f32_smem_kernel<8, 8, KB, STAGES>;
      PBX_CUDA_CHECK(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
      kern<<<(unsigned)blocks, 256, SMEM, h->stream>>>(p);
    } else {
      constexpr int SMEM = STAGES * KB * (4 * 4 + 2 * 4) * 32 * 4;
      auto kern = gemm_interleaved_f32_smem_kernel<4, 4, KB, STAGES>;
      PBX_CUDA_CHECK(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
      kern<<<(unsigned)blocks, 256, SMEM, h->stream>>>(p);
    }
    h->launches++;
    PBX_CUDA_CHECK(h, cudaGetLastError());
    return PBX_OK;
  }
  return dispatch_dtype(c.dtype, [&](auto* ti, auto* to, auto* ta) -> int {
    using TIn = std::remove_pointer_t<decltype(ti)>;
    using TOut = std::remove_pointer_t<decltype(to)>;
    using TAcc = std::remove_pointer_t<decltype(ta)>;
    // K steps whose loads are issued together (in-flight loads per thread = (RM + RN) x unroll)
    static const int env_unroll = getenv("PBX_ILV_UNROLL") ? atoi(getenv("PBX_ILV_UNROLL")) : 0;
    if constexpr (!std::is_same<TAcc, double>::value) {
      if (tile == 8) {
        if (env_unroll == 1) gemm_interleaved_kernel<TIn, TOut, TAcc, 8, 8, 1><<<(unsigned)blocks, 256, 0, h->stream>>>(p);
        else gemm_interleaved_kernel<TIn, TOut, TAcc, 8, 8, 2><<<(unsigned)blocks, 256, 0, h->stream>>>(p);
        h->launches++;
        PBX_CUDA_CHECK(h, cudaGetLastError());
        return PBX_OK;
      }
    }
    const int unroll = env_unroll ? env_unroll : 4;
    if (unroll >= 4) gemm_interleaved_kernel<TIn, TOut, TAcc, 4, 4, 4><<<(unsigned)blocks, 256, 0, h->stream>>>(p);
    else if (unroll >= 2) gemm_interleaved_kernel<TIn, TOut, TAcc, 4, 4, 2><<<(unsigned)blocks, 256, 0, h->stream>>>(p);
    else gemm_interleaved_kernel<TIn, TOut, TAcc, 4, 4, 1><<<(unsigned)blocks, 256, 0, h->stream>>>(p);
    h->launches++;
    PBX_CUDA_CHECK(h, cudaGetLastError());
    return PBX_OK;
  });
}

int pbx_launch_scal(pbx_handle_t h, int dtype, int64_t m, int64_t n, double beta, void* C,
                    int64_t ldc, int64_t stridec, int64_t batch, int interleaved) {
  if (interleaved) {
    // (c*ldc + r)*batch + b  ==  column-major matrix with m*batch rows and ld = ldc*batch
    m = m * batch; ldc = ldc * batch; batch = 1; stridec = 0;
  }
  const int64_t total = m * n * batch;
  if (total == 0) return PBX_OK;
  int64_t blocks = (total + 255) / 256;
  const int64_t cap = (int64_t)h->sm_count * 16;
  if (blocks > cap) blocks = cap;
  return dispatch_dtype(dtype, [&](auto* ti, auto* to, auto* ta) -> int {
    using TOut = std::remove_pointer_t<decltype(to)>;
    using TAcc = std::remove_pointer_t<decltype(ta)>;
    scal_matrix_kernel<TOut, TAcc><<<(unsigned)blocks, 256, 0, h->stream>>>(
        (TOut*)C, m, n, ldc, stridec, batch, beta);
    h->launches++;
    PBX_CUDA_CHECK(h, cudaGetLastError());
    return PBX_OK;
  });
}

int pbx_launch_splitk_reduce(pbx_handle_t h, const PbxGemmCall& c, int slices) {
  const bool vec4 = (c.m % 4 == 0);
  const int64_t total = c.m * c.n * c.batch / (vec4 ? 4 : 1);
  int64_t blocks = (total + 255) / 256;
  const int64_t cap = (int64_t)h->sm_count * 16;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return dispatch_dtype(c.dtype, [&](auto* ti, auto* to, auto* ta) -> int {
    using TOut = std::remove_pointer_t<decltype(to)>;
    using TAcc = std::remove_pointer_t<decltype(ta)>;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)blocks);
    cfg.blockDim = dim3(256);
    cfg.stream = h->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    // Programmatic launch lets this grid become resident while the GEMM drains -- a gain when the GEMM leaves SMs free
    // (SGEMM 1024^3: 19.3 -> 18.1 us), a loss when it fills the machine (64 x 147 x 13225 on 206 CTAs: 35 -> 46 us), so
    // it is requested only behind a tcgen05 grid smaller than the SM count
    cfg.numAttrs = (h->pdl && h->pdl_reduce && h->last_grid_ctas < h->sm_count) ? 1 : 0;
    h->last_grid_ctas = 1 << 30;
    const TAcc* ws = (const TAcc*)h->ws;
    TOut* C = (TOut*)c.C;
    if (vec4)
      PBX_CUDA_CHECK(h, cudaLaunchKernelEx(&cfg, splitk_reduce_kernel<TOut, TAcc, 4>, ws, C, c.m, c.n, c.ldc, c.sc, c.batch,
                                           slices, c.alpha, c.beta));
    else
      PBX_CUDA_CHECK(h, cudaLaunchKernelEx(&cfg, splitk_reduce_kernel<TOut, TAcc, 1>, ws, C, c.m, c.n, c.ldc, c.sc, c.batch,
                                           slices, c.alpha, c.beta));
    h->launches++;
    PBX_CUDA_CHECK(h, cudaGetLastError());
    return PBX_OK;
  });
}
