// gemm_dmma.cu -- fp64 GEMM on the FP64 tensor path (DMMA.8x8x4; tcgen05 has no f64 kind).
//
// Replaces the reference's double instantiation of Gemm<...local...>
// (src/operations/blas3/gemm_local.hpp:427-517,738-773: smem tiles + scalar mad) for
// BASELINE cfg 2 (DGEMM 8192^3, all four transposes, alpha/beta).
//
// Design: 128x128x16 block tile, warp-specialised.
//   * 8 consumer warps (2 x 4), each a 64x32 warp tile of m8n8k4 fragments.  They execute nothing
//     but fragment loads (LDS.64, double-buffered in registers) and DMMAs: measured on B200
//     (tools/micro/dmma_rate.cu) the FP64 pipe takes one DMMA.8x8x4 per 16 cycles per SM
//     sub-partition and two such warps per sub-partition keep it 98% busy -- but only while neither
//     of them is off doing address arithmetic or sitting in a CTA-wide barrier (that was worth 15%).
//   * 4 producer warps (one warp group) fill a 4-stage shared-memory ring with cp.async (16 B copies when base / ld
//     allow, else 8 B; zero-fill at ragged edges through the src-size operand) and signals each
//     stage through an mbarrier (cp.async.mbarrier.arrive); consumers hand stages back through a
//     second mbarrier.  There is no __syncthreads in the main loop.
// Shared-memory row strides are == 4 (mod 16) doubles so every 64-bit fragment load is
// bank-conflict free per half-warp.
#include <stdio.h>

#include <type_traits>

#include <atomic>

#include "pbx_internal.cuh"
#include "tc_ptx.cuh"

namespace {

using tcx::mbar_arrive;
using tcx::mbar_init;
using tcx::mbar_wait;

constexpr int DBM = 128, DBN = 128, DBK = 16;
constexpr int DSTAGES = 4;
constexpr int DPROD = 4;                       // producer warps = warp group 0 (runs on a trimmed register budget)
constexpr int DCONS = 8;                       // consumer warps = warp groups 1 and 2
constexpr int DTHREADS = (DPROD + DCONS) * 32;
// Registers: the file is split per sub-partition (16K each, 3 of these warps apiece).  The kernel
// launches at <= 168 regs/thread; consumers then raise themselves to 232 (128 accumulator + 48
// double-buffered fragment registers) and the producers drop to 40:  2*32*232 + 32*40 = 16128.
constexpr int DREG_CONS = 232, DREG_PROD = 40;
constexpr int LD_MN = DBM + 4;  // [k][mn] layout, mn contiguous (132 doubles)
constexpr int LD_K = DBK + 4;   // [mn][k] layout, k contiguous  (20 doubles)
constexpr int TILE_MN_ELEMS = DBK * LD_MN;   // 2112
constexpr int TILE_K_ELEMS = DBM * LD_K;     // 2560
constexpr int TILE_ELEMS = TILE_K_ELEMS;     // max of both
constexpr int DMMA_SMEM_BYTES = DSTAGES * 2 * TILE_ELEMS * 8 + 2 * DSTAGES * 8;  // tiles + full/empty barriers

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
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_8(uint32_t dst, const void* src, int src_bytes) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
// the mbarrier receives one (pre-counted) arrival once all cp.async issued so far by this thread have landed
__device__ __forceinline__ void cp_async_mbar_arrive(uint32_t bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// Producer-side copy of one 128 x 16 operand tile by ONE warp.  X(mn,k) lives at
// KContig ? X[k + mn*ld] : X[mn + k*ld].  KContig tiles go to smem as [mn][k] (stride LD_K), the
// others as [k][mn] (stride LD_MN).  Lane l owns a fixed column of VEC-double chunks and walks the
// other dimension, so the interior path is one pointer add + one cp.async per chunk.
template <bool KContig, int VEC>
struct TileCopy {
  static constexpr int CH_PER_ROW = (KContig ? DBK : DBM) / VEC;   // chunks along the contiguous dimension
  static constexpr int ROWS = KContig ? DBM : DBK;                 // strided dimension
  static constexpr int LD_S = KContig ? LD_K : LD_MN;
  static constexpr int LANES_PER_ROW = CH_PER_ROW < 32 ? CH_PER_ROW : 32;
  static constexpr int ROWS_PER_PASS = 32 / LANES_PER_ROW;         // rows covered by one warp-wide cp.async
  static constexpr int COL_PASSES = CH_PER_ROW / LANES_PER_ROW;    // > 1 when a row has more than 32 chunks
  static constexpr int PASSES = ROWS / ROWS_PER_PASS;

  // mn0/k0: tile origin; mn_total/k_end: matrix extents
  __device__ static __forceinline__ void run(uint32_t dst_tile, const double* __restrict__ X, int64_t ld,
                                             int64_t mn0, int64_t k0, int64_t mn_total, int64_t k_end, int lane,
                                             int pw, bool full) {
    const int c_lane = (lane % LANES_PER_ROW) * VEC;    // offset along the contiguous dimension
    const int r_lane = lane / LANES_PER_ROW;            // row inside a pass
    const int64_t c0 = KContig ? k0 : mn0, r0 = KContig ? mn0 : k0;
    const int64_t c_end = KContig ? k_end : mn_total, r_end = KContig ? mn_total : k_end;
#pragma unroll
    for (int cp = 0; cp < COL_PASSES; ++cp) {
      const int cc = c_lane + cp * 32 * VEC;
      constexpr int PPW = PASSES / DPROD;               // passes per producer warp
      const int row_first = r_lane + pw * PPW * ROWS_PER_PASS;
      const double* src = X + (c0 + cc) + (r0 + row_first) * ld;
      uint32_t dst = dst_tile + (uint32_t)((row_first * LD_S + cc) * 8);
      const int64_t step = (int64_t)ROWS_PER_PASS * ld;
      if (full) {
#pragma unroll
        for (int ps = 0; ps < PPW; ++ps) {
          if (VEC == 2) cp_async_16(dst, src, 16); else cp_async_8(dst, src, 8);
          src += step;
          dst += ROWS_PER_PASS * LD_S * 8;
        }
      } else {
        const int cvalid = (int)min((int64_t)VEC, max((int64_t)0, c_end - (c0 + cc)));
#pragma unroll 4
        for (int ps = 0; ps < PPW; ++ps) {
          const int64_t r = r0 + row_first + ps * ROWS_PER_PASS;
          const int valid = (r < r_end) ? cvalid : 0;
          const double* sp = valid > 0 ? src : X;
          if (VEC == 2) cp_async_16(dst, sp, valid * 8); else cp_async_8(dst, sp, valid * 8);
          src += step;
          dst += ROWS_PER_PASS * LD_S * 8;
        }
      }
    }
  }
};

template <bool AK, bool BK_, int VEC>
__global__ void __launch_bounds__(DTHREADS, 1) gemm_dmma_kernel(DmmaParams p) {
  extern __shared__ __align__(16) double dsmem[];
  double* sA = dsmem;
  double* sB = dsmem + DSTAGES * TILE_ELEMS;
  const uint32_t sA_u32 = (uint32_t)__cvta_generic_to_shared(sA);
  const uint32_t sB_u32 = (uint32_t)__cvta_generic_to_shared(sB);
  const uint32_t bar_u32 = sA_u32 + (uint32_t)(2 * DSTAGES * TILE_ELEMS * 8);
  auto full_bar = [&](int s) { return bar_u32 + 8u * s; };
  auto empty_bar = [&](int s) { return bar_u32 + 8u * (DSTAGES + s); };
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);

  if (tid == 0) {
    for (int s = 0; s < DSTAGES; ++s) {
      mbar_init(full_bar(s), DPROD * 32);  // one cp.async.mbarrier.arrive per producer thread
      mbar_init(empty_bar(s), DCONS);  // one arrive per consumer warp
    }
    tcx::fence_barrier_init();
  }
  __syncthreads();

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
  uint32_t it = 0;   // K blocks processed so far by this role (ring position = it % DSTAGES)

  if (warp < DPROD) {
    // ============================ producer warps ============================
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(DREG_PROD));
    const bool a_full = (m0 + DBM <= p.m), b_full = (n0 + DBN <= p.n);
    for (int64_t b = blockIdx.z; b < p.batch; b += gridDim.z) {
      const double* A = p.A + b * p.sa;
      const double* B = p.B + b * p.sb;
      for (int kb = 0; kb < nkb; ++kb, ++it) {
        const int s = (int)(it % DSTAGES);
        const uint32_t ph = (it / DSTAGES) & 1u;
        mbar_wait(empty_bar(s), ph ^ 1u);
        const int64_t k0 = (kb_beg + kb) * DBK;
        const bool k_full = (k0 + DBK <= p.k);
        TileCopy<AK, VEC>::run(sA_u32 + (uint32_t)(s * TILE_ELEMS * 8), A, p.lda, m0, k0, p.m, p.k, lane, warp,
                               a_full && k_full);
        TileCopy<BK_, VEC>::run(sB_u32 + (uint32_t)(s * TILE_ELEMS * 8), B, p.ldb, n0, k0, p.n, p.k, lane, warp,
                                b_full && k_full);
        cp_async_mbar_arrive(full_bar(s));
      }
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
    return;
  }

  // ============================ consumer warps ============================
  asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(DREG_CONS));
  const int cw = warp - DPROD;
  const int wm0 = (cw & 1) * 64, wn0 = (cw >> 1) * 32;
  const int lr = lane >> 2, lc = lane & 3;
  // per-thread fragment offsets inside a stage tile (doubles)
  const int a_off = AK ? (wm0 + lr) * LD_K + lc : lc * LD_MN + wm0 + lr;
  const int b_off = BK_ ? (wn0 + lr) * LD_K + lc : lc * LD_MN + wn0 + lr;
  constexpr int A_I = AK ? 8 * LD_K : 8;          // next 8 rows
  constexpr int B_J = BK_ ? 8 * LD_K : 8;
  constexpr int A_KK = AK ? 4 : 4 * LD_MN;        // next k4 step
  constexpr int B_KK = BK_ ? 4 : 4 * LD_MN;

  for (int64_t b = blockIdx.z; b < p.batch; b += gridDim.z) {
    double acc[8][4][2];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    for (int kb = 0; kb < nkb; ++kb, ++it) {
      const int s = (int)(it % DSTAGES);
      const uint32_t ph = (it / DSTAGES) & 1u;
      mbar_wait(full_bar(s), ph);
      const double* tA = sA + s * TILE_ELEMS + a_off;
      const double* tB = sB + s * TILE_ELEMS + b_off;
      double af[2][8], bf[2][4];
#pragma unroll
      for (int i = 0; i < 8; ++i) af[0][i] = tA[i * A_I];
#pragma unroll
      for (int j = 0; j < 4; ++j) bf[0][j] = tB[j * B_J];
#pragma unroll
      for (int q = 0; q < DBK / 4; ++q) {
        const int cur = q & 1, nxt = cur ^ 1;
        if (q + 1 < DBK / 4) {   // fragments of the next k4 step load under this step's DMMAs
#pragma unroll
          for (int i = 0; i < 8; ++i) af[nxt][i] = tA[(q + 1) * A_KK + i * A_I];
#pragma unroll
          for (int j = 0; j < 4; ++j) bf[nxt][j] = tB[(q + 1) * B_KK + j * B_J];
        }
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) dmma884(acc[i][j][0], acc[i][j][1], af[cur][i], bf[cur][j]);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(empty_bar(s));
    }

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
  static std::atomic<uint32_t> attr_set{0};   // bit per device ordinal: the attribute is sticky, set it once
  if (h->device >= 32 || !(attr_set.load(std::memory_order_acquire) & (1u << h->device))) {
    PBX_CUDA_CHECK(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, DMMA_SMEM_BYTES));
    if (h->device < 32) attr_set.fetch_or(1u << h->device, std::memory_order_release);
  }
  kern<<<grid, DTHREADS, DMMA_SMEM_BYTES, h->stream>>>(p);
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
