// gemm_dmma.cu -- fp64 GEMM on the FP64 tensor path (DMMA.8x8x4; tcgen05 has no f64 kind).
//
// Replaces the reference's double instantiation of Gemm<...local...>
// (src/operations/blas3/gemm_local.hpp:427-517,738-773: smem tiles + scalar mad) for
// BASELINE cfg 2 (DGEMM 8192^3, all four transposes, alpha/beta).
//
// Design: 128x128x16 block tile, 8 warps (2 x 4), 64x32 warp tile of m8n8k4 fragments,
// 4-stage cp.async ring (16 B copies when base/ld allow, else 8 B), zero-fill at ragged
// edges through the cp.async src-size operand.  Shared-memory row strides are == 4 (mod 16)
// doubles so every 64-bit fragment load is bank-conflict free per half-warp.
#include <stdio.h>

#include <type_traits>

#include "pbx_internal.cuh"

namespace {

constexpr int DBM = 128, DBN = 128, DBK = 16;
constexpr int DSTAGES = 4;
constexpr int LD_MN = DBM + 4;  // [k][mn] layout, mn contiguous (132 doubles)
constexpr int LD_K = DBK + 4;   // [mn][k] layout, k contiguous  (20 doubles)
constexpr int TILE_MN_ELEMS = DBK * LD_MN;   // 2112
constexpr int TILE_K_ELEMS = DBM * LD_K;     // 2560
constexpr int TILE_ELEMS = TILE_K_ELEMS;     // max of both
constexpr int DMMA_SMEM_BYTES = DSTAGES * 2 * TILE_ELEMS * 8;  // 163840

struct DmmaParams {
  const double* A;
  const double* B;
  double* C;
  double* ws;
  int64_t m, n, k, lda, ldb, ldc, sa, sb, sc, batch;
  int64_t kb_per_slice;  // in units of DBK
  int slices, m_tiles, n_tiles, group_m;
  double alpha, beta;
};

__device__ __forceinline__ void cp_async_16(uint32_t dst, const void* src, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_8(uint32_t dst, const void* src, int src_bytes) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(dst), "l"(src), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N)); }

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// Per-thread copy plan of one operand: chunk i of this thread covers VEC doubles at
//   KContig : (mn = mn_t + i*MN_STEP, k = k_t)        MN_STEP = 256 / (DBK/VEC)
//   else    : (mn = mn_t,             k = k_t + i*K_STEP)   K_STEP = 256 / (DBM/VEC)
// The global pointer of chunk 0 advances by a constant per K block, so the interior fast path
// is one 64-bit add per operand per K block plus PER_THREAD cp.async with immediate offsets.
template <bool KContig, int VEC>
struct TilePlan {
  static constexpr int CHUNKS = DBM * DBK / VEC;
  static constexpr int PER_THREAD = CHUNKS / 256;
  static constexpr int MN_STEP = KContig ? 256 / (DBK / VEC) : 0;
  static constexpr int K_STEP = KContig ? 0 : 256 / (DBM / VEC);
  static constexpr int DST_STEP = (KContig ? MN_STEP * LD_K : K_STEP * LD_MN) * 8;  // bytes between chunks
  const double* src;     // chunk 0 at the current K block
  int64_t chunk_stride;  // elements between consecutive chunks of this thread
  int64_t kb_stride;     // elements per K block
  uint32_t dst0;         // byte offset of chunk 0 inside a stage tile
  int mn_t, k_t;
  __device__ __forceinline__ void init(const double* X, int64_t ld, int64_t mn0, int64_t k0, int tid) {
    if (KContig) { k_t = (tid % (DBK / VEC)) * VEC; mn_t = tid / (DBK / VEC); }
    else         { mn_t = (tid % (DBM / VEC)) * VEC; k_t = tid / (DBM / VEC); }
    src = KContig ? X + (k0 + k_t) + (mn0 + mn_t) * ld : X + (mn0 + mn_t) + (k0 + k_t) * ld;
    chunk_stride = KContig ? (int64_t)MN_STEP * ld : (int64_t)K_STEP * ld;
    kb_stride = KContig ? (int64_t)DBK : (int64_t)DBK * ld;
    dst0 = (uint32_t)((KContig ? mn_t * LD_K + k_t : k_t * LD_MN + mn_t) * 8);
  }
  // chunks [I0, I1) of the tile, all elements known to exist
  template <int I0, int I1>
  __device__ __forceinline__ void copy_full(uint32_t stage_base) const {
#pragma unroll
    for (int i = I0; i < I1; ++i) {
      if (VEC == 2) cp_async_16(stage_base + dst0 + i * DST_STEP, src + i * chunk_stride, 16);
      else cp_async_8(stage_base + dst0 + i * DST_STEP, src + i * chunk_stride, 8);
    }
  }
  // same with per-chunk edge predication (zero fill)
  template <int I0, int I1>
  __device__ __forceinline__ void copy_edge(uint32_t stage_base, const double* X, int64_t mn0, int64_t k0,
                                            int64_t mn_total, int64_t k_end) const {
#pragma unroll
    for (int i = I0; i < I1; ++i) {
      const int64_t gmn = mn0 + mn_t + (KContig ? i * MN_STEP : 0);
      const int64_t gk = k0 + k_t + (KContig ? 0 : i * K_STEP);
      int valid;
      if (KContig) valid = (gmn < mn_total) ? (int)min((int64_t)VEC, max((int64_t)0, k_end - gk)) : 0;
      else         valid = (gk < k_end) ? (int)min((int64_t)VEC, max((int64_t)0, mn_total - gmn)) : 0;
      const double* sp = valid > 0 ? src + i * chunk_stride : X;
      if (VEC == 2) cp_async_16(stage_base + dst0 + i * DST_STEP, sp, valid * 8);
      else cp_async_8(stage_base + dst0 + i * DST_STEP, sp, valid * 8);
    }
  }
};

template <bool AK, bool BK_, int VEC>
__global__ void __launch_bounds__(256, 1) gemm_dmma_kernel(DmmaParams p) {
  extern __shared__ __align__(16) double dsmem[];
  double* sA = dsmem;
  double* sB = dsmem + DSTAGES * TILE_ELEMS;
  const uint32_t sA_u32 = (uint32_t)__cvta_generic_to_shared(sA);
  const uint32_t sB_u32 = (uint32_t)__cvta_generic_to_shared(sB);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm0 = (warp & 1) * 64, wn0 = (warp >> 1) * 32;
  const int lr = lane >> 2, lc = lane & 3;

  // grouped rasterisation: GROUP_M row-tiles share each B panel while it is hot in L2
  const int tiles_per_group = p.group_m * p.n_tiles;
  const int t = blockIdx.x;
  const int g = t / tiles_per_group;
  const int first_m = g * p.group_m;
  const int gsize = min(p.m_tiles - first_m, p.group_m);
  const int mt = first_m + (t % tiles_per_group) % gsize;
  const int nt = (t % tiles_per_group) / gsize;
  const int64_t m0 = (int64_t)mt * DBM, n0 = (int64_t)nt * DBN;
  const int slice = blockIdx.y;
  const int64_t kb_total = (p.k + DBK - 1) / DBK;
  const int64_t kb_beg = (int64_t)slice * p.kb_per_slice;
  const int64_t kb_end = min(kb_total, kb_beg + p.kb_per_slice);
  const int nkb = (int)(kb_end - kb_beg);
  // interior tile: every row/column of the 128x128 tile exists -> only the last K block may need predication
  const bool a_full = (m0 + DBM <= p.m), b_full = (n0 + DBN <= p.n);
  constexpr int PT = TilePlan<AK, VEC>::PER_THREAD;   // cp.async per thread per operand per K block (4 or 8)
  constexpr int Q = PT / 4;                            // issued per k4 step

  for (int64_t b = blockIdx.z; b < p.batch; b += gridDim.z) {
    const double* A = p.A + b * p.sa;
    const double* B = p.B + b * p.sb;
    double acc[8][4][2];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    TilePlan<AK, VEC> pa;
    TilePlan<BK_, VEC> pb;
    pa.init(A, p.lda, m0, kb_beg * DBK, tid);
    pb.init(B, p.ldb, n0, kb_beg * DBK, tid);
    int issued = 0;  // K blocks whose copies have been issued (plans point at block `issued`)

    // quarter q (0..3) of the copies of K block `issued` into its ring slot
    auto issue_part = [&](auto qc) {
      constexpr int q = decltype(qc)::value;
      if (issued < nkb) {
        const int s = issued % DSTAGES;
        const int64_t k0 = (kb_beg + issued) * DBK;
        const uint32_t da = sA_u32 + (uint32_t)(s * TILE_ELEMS * 8), db = sB_u32 + (uint32_t)(s * TILE_ELEMS * 8);
        const bool k_full = (k0 + DBK <= p.k);
        if (a_full && k_full) pa.template copy_full<q * Q, (q + 1) * Q>(da);
        else pa.template copy_edge<q * Q, (q + 1) * Q>(da, A, m0, k0, p.m, p.k);
        if (b_full && k_full) pb.template copy_full<q * Q, (q + 1) * Q>(db);
        else pb.template copy_edge<q * Q, (q + 1) * Q>(db, B, n0, k0, p.n, p.k);
      }
    };
    auto issue_done = [&]() {
      if (issued < nkb) { pa.src += pa.kb_stride; pb.src += pb.kb_stride; }
      ++issued;
      cp_async_commit();
    };
    using Q0 = std::integral_constant<int, 0>; using Q1 = std::integral_constant<int, 1>;
    using Q2 = std::integral_constant<int, 2>; using Q3 = std::integral_constant<int, 3>;
#pragma unroll
    for (int s = 0; s < DSTAGES - 1; ++s) {
      issue_part(Q0{}); issue_part(Q1{}); issue_part(Q2{}); issue_part(Q3{});
      issue_done();
    }

    for (int kb = 0; kb < nkb; ++kb) {
      cp_async_wait<DSTAGES - 2>();
      __syncthreads();
      // the copies of block kb+DSTAGES-1 (into the slot consumed in iteration kb-1) are spread over the
      // four k4 steps so that the tensor pipe never waits behind a burst of address arithmetic
      const double* tA = sA + (kb % DSTAGES) * TILE_ELEMS;
      const double* tB = sB + (kb % DSTAGES) * TILE_ELEMS;
#pragma unroll
      for (int kk = 0; kk < DBK; kk += 4) {
        double af[8], bf[4];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int r = wm0 + i * 8 + lr;
          af[i] = AK ? tA[r * LD_K + kk + lc] : tA[(kk + lc) * LD_MN + r];
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int c = wn0 + j * 8 + lr;
          bf[j] = BK_ ? tB[c * LD_K + kk + lc] : tB[(kk + lc) * LD_MN + c];
        }
        if (kk == 0) issue_part(Q0{});
        else if (kk == 4) issue_part(Q1{});
        else if (kk == 8) issue_part(Q2{});
        else issue_part(Q3{});
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) dmma884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
      }
      issue_done();
    }
    cp_async_wait<0>();
    __syncthreads();

    // epilogue: fragment (i,j) holds C[wm0+i*8+lr][wn0+j*8+lc*2 + {0,1}]
    if (p.slices > 1) {
      double* ws = p.ws + ((b * p.slices + slice) * p.n) * p.m;
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int64_t col = n0 + wn0 + j * 8 + lc * 2 + e;
          if (col >= p.n) continue;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int64_t row = m0 + wm0 + i * 8 + lr;
            if (row < p.m) ws[col * p.m + row] = acc[i][j][e];
          }
        }
    } else {
      double* C = p.C + b * p.sc;
      const bool beta0 = (p.beta == 0.0);
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int64_t col = n0 + wn0 + j * 8 + lc * 2 + e;
          if (col >= p.n) continue;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int64_t row = m0 + wm0 + i * 8 + lr;
            if (row >= p.m) continue;
            double* dst = C + row + col * p.ldc;
            double r = p.alpha * acc[i][j][e];
            if (!beta0) r += p.beta * (*dst);
            *dst = r;
          }
        }
    }
  }
}

template <bool AK, bool BK_, int VEC>
int launch_variant(pbx_handle_t h, const DmmaParams& p, dim3 grid) {
  auto kern = gemm_dmma_kernel<AK, BK_, VEC>;
  PBX_CUDA_CHECK(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, DMMA_SMEM_BYTES));
  kern<<<grid, 256, DMMA_SMEM_BYTES, h->stream>>>(p);
  h->launches++;
  PBX_CUDA_CHECK(h, cudaGetLastError());
  return PBX_OK;
}

}  // namespace

int pbx_launch_dmma(pbx_handle_t h, const PbxGemmCall& c, int slices) {
  DmmaParams p;
  p.A = (const double*)c.A; p.B = (const double*)c.B; p.C = (double*)c.C; p.ws = (double*)h->ws;
  p.m = c.m; p.n = c.n; p.k = c.k; p.lda = c.lda; p.ldb = c.ldb; p.ldc = c.ldc;
  p.sa = c.sa; p.sb = c.sb; p.sc = c.sc; p.batch = c.batch;
  p.slices = slices;
  const int64_t kb_total = (c.k + DBK - 1) / DBK;
  p.kb_per_slice = (kb_total + slices - 1) / slices;
  p.m_tiles = (int)((c.m + DBM - 1) / DBM);
  p.n_tiles = (int)((c.n + DBN - 1) / DBN);
  p.group_m = 8;
  p.alpha = c.alpha; p.beta = c.beta;
  dim3 grid((unsigned)((int64_t)p.m_tiles * p.n_tiles), (unsigned)slices,
            (unsigned)(c.batch < 65535 ? c.batch : 65535));
  // 16-byte copies need 16 B aligned bases and even leading dimensions / batch strides
  const bool vec2 = (((uintptr_t)c.A | (uintptr_t)c.B) % 16 == 0) && (c.lda % 2 == 0) &&
                    (c.ldb % 2 == 0) && (c.sa % 2 == 0) && (c.sb % 2 == 0);
  const bool ak = c.ta;    // op(A)=A^T is stored K x M: k contiguous
  const bool bk = !c.tb;   // op(B)=B   is stored K x N: k contiguous
#define PBX_DMMA_CASE(AKV, BKV)                                                        \
  if (ak == AKV && bk == BKV)                                                          \
    return vec2 ? launch_variant<AKV, BKV, 2>(h, p, grid) : launch_variant<AKV, BKV, 1>(h, p, grid);
  PBX_DMMA_CASE(false, false)
  PBX_DMMA_CASE(false, true)
  PBX_DMMA_CASE(true, false)
  PBX_DMMA_CASE(true, true)
#undef PBX_DMMA_CASE
  return PBX_ERR_INVALID_ARG;
}
