// gemm_tcgen05.cu -- the Blackwell tensor-core GEMM: TMA -> 128B-swizzled smem ring ->
// tcgen05.mma (accumulators in TMEM) -> tcgen05.ld epilogue with alpha/beta.
//
// Replaces the reference's NVIDIA production kernels:
//   Gemm<..., local, standard, full, ...>      src/operations/blas3/gemm_local.hpp:263-355,427-517,738-773
//   Gemm<..., joint_matrix ...> (mma.sync)      src/operations/blas3/gemm_local_joint_matrix.hpp:274-495,794-830
//   the batch-in-grid scheme                    gemm_local.hpp:273-279,511-516
// and their selection (src/interface/blas3/backend/nvidia_gpu.hpp:68-171).
//
// One persistent CTA per SM walks a static tile schedule (batch x K-slice x M-tile x N-tile).
// Warp roles: w0 TMA producer, w1 MMA issuer (one lane), w2 TMEM allocator, w4-7 epilogue
// (TMEM lane quarter = warp % 4), w8-11 (fp32 only) hi/lo splitters for 3xTF32.
// Accumulators are double-buffered in TMEM so the epilogue of tile i overlaps the
// mainloop of tile i+1.
//
// Operand layouts (column-major BLAS):
//   op(A)=A   : stored M x K, M contiguous -> "MN-major" UMMA operand
//   op(A)=A^T : stored K x M, K contiguous -> "K-major"
//   op(B)=B   : stored K x N, K contiguous -> "K-major"
//   op(B)=B^T : stored N x K, N contiguous -> "MN-major"
// D (128 x BN fp32) sits in TMEM with row m on lane m, so a warp's 32 lanes hold 32
// consecutive rows of one column: stores to column-major C are fully coalesced.
//
// fp32 (3xTF32): a = hi + lo with hi = tf32(a), lo = tf32(a - hi);
//   D += A_lo*B_hi + A_hi*B_lo + A_hi*B_hi   (fp32 accumulate, lo*lo dropped: ~2^-22 relative)
// The tensor core adds into its fp32 accumulator with truncation, a one-sided error that grows
// linearly with the length of the accumulation chain (measured: ~7e-9 relative per k).  To keep
// fp32 results within ~1e-5 the K loop is cut into chunks of kb_per_chunk blocks; each chunk
// starts a fresh TMEM accumulator and the epilogue warps fold finished chunks into a running sum
// (a third TMEM region) with round-to-nearest CUDA-core adds while the next chunk is in flight.
#include <stdio.h>

#include <mutex>

#include <cudaTypedefs.h>

#include "pbx_internal.cuh"
#include "tc_ptx.cuh"

namespace {

using namespace tcx;

struct TcParams {
  void* C;
  float* ws;
  int64_t M, N, K, ldc, sc;
  float alpha, beta;
  int batch, slices, m_tiles, n_tiles, group_m;
  int kb_total, kb_per_slice;
  int kb_per_chunk;  // K blocks accumulated inside the tensor core before an fp32 RN add (see below)
  int a_batched, b_batched;
  int64_t total_tiles;
};

template <typename T> struct OutCvt;
template <> struct OutCvt<float> {
  __device__ static float load(const float* p) { return *p; }
  __device__ static void store(float* p, float v) { *p = v; }
};
template <> struct OutCvt<__half> {
  __device__ static float load(const __half* p) { return __half2float(*p); }
  __device__ static void store(__half* p, float v) { *p = __float2half_rn(v); }
};
template <> struct OutCvt<__nv_bfloat16> {
  __device__ static float load(const __nv_bfloat16* p) { return __bfloat162float(*p); }
  __device__ static void store(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }
};

constexpr int BM = 128;
constexpr int ROW_BYTES = 128;  // one swizzle row

template <int ES, int BN, int STAGES>
struct TcCfg {
  static constexpr bool TF32X3 = (ES == 4);
  static constexpr int BK = ROW_BYTES / ES;        // 64 (16-bit) or 32 (fp32) elements
  static constexpr int UMMA_K = 32 / ES;           // 16 or 8
  static constexpr int A_BYTES = BM * ROW_BYTES;   // 16 KiB
  static constexpr int B_BYTES = BN * ROW_BYTES;
  static constexpr int RAW_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGE_BYTES = RAW_BYTES * (TF32X3 ? 2 : 1);  // + lo copies
  static constexpr int BAR_BYTES = 256;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + BAR_BYTES + 1024;  // + align slack
  // two MMA accumulator stages; fp32 adds a running-sum region (3*128 = 384 -> 512 columns)
  static constexpr int TMEM_COLS = TF32X3 ? 512 : 2 * BN;
  static constexpr int RSUM_COL = 2 * BN;          // running sum of chunk partials (fp32 only)
  static constexpr int NUM_THREADS = TF32X3 ? 384 : 256;
  static constexpr int NUM_SPLIT_THREADS = 128;
};

struct TileCoord {
  int b, slice, mt, nt;
};

__device__ __forceinline__ TileCoord decode_tile(const TcParams& p, int64_t tile) {
  const int per = p.m_tiles * p.n_tiles;
  const int t = (int)(tile % per);
  const int64_t bs = tile / per;
  TileCoord c;
  c.slice = (int)(bs % p.slices);
  c.b = (int)(bs / p.slices);
  const int tiles_per_group = p.group_m * p.n_tiles;
  const int g = t / tiles_per_group;
  const int first_m = g * p.group_m;
  const int gsize = min(p.m_tiles - first_m, p.group_m);
  c.mt = first_m + (t % tiles_per_group) % gsize;
  c.nt = (t % tiles_per_group) / gsize;
  return c;
}

template <typename TIn, typename TOut, int BN, int STAGES, bool A_MN, bool B_MN>
__global__ void __launch_bounds__(TcCfg<sizeof(TIn), BN, STAGES>::NUM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const TcParams p) {
  using Cfg = TcCfg<sizeof(TIn), BN, STAGES>;
  constexpr bool TF32X3 = Cfg::TF32X3;
  constexpr int BK = Cfg::BK;
  constexpr uint32_t FMT = TF32X3 ? 2u : (std::is_same<TIn, __nv_bfloat16>::value ? 1u : 0u);

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + STAGES * Cfg::STAGE_BYTES;
  // barrier map (8 B each): full[S] | empty[S] | split[S] | tmem_full[2] | tmem_empty[2] | tmem_ptr
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  auto split_bar = [&](int s) { return bar_base + 8u * (2 * STAGES + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (3 * STAGES + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (3 * STAGES + 2 + s); };
  const uint32_t tmem_ptr_smem = bar_base + 8u * (3 * STAGES + 4);
  volatile uint32_t* tmem_ptr_generic = reinterpret_cast<volatile uint32_t*>(
      smem_raw + (tmem_ptr_smem - smem_u32(smem_raw)));

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
      mbar_init(split_bar(s), Cfg::NUM_SPLIT_THREADS);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull_bar(s), 1);
      mbar_init(tempty_bar(s), 4);
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_ptr_smem, Cfg::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_generic;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int64_t tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        const TileCoord tc = decode_tile(p, tile);
        const int m0 = tc.mt * BM, n0 = tc.nt * BN;
        const int kb0 = tc.slice * p.kb_per_slice;
        const int kb1 = min(p.kb_total, kb0 + p.kb_per_slice);
        const int za = p.a_batched ? tc.b : 0, zb = p.b_batched ? tc.b : 0;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1u);
          const uint32_t sA = smem_base + stage * Cfg::STAGE_BYTES;
          const uint32_t sB = sA + Cfg::A_BYTES;
          mbar_expect_tx(full_bar(stage), Cfg::RAW_BYTES);
          if (A_MN) {
#pragma unroll
            for (int c = 0; c < BM / BK; ++c)
              tma_load_3d(sA + c * BK * ROW_BYTES, &tmA, full_bar(stage), m0 + c * BK, kb * BK, za);
          } else {
            tma_load_3d(sA, &tmA, full_bar(stage), kb * BK, m0, za);
          }
          if (B_MN) {
#pragma unroll
            for (int c = 0; c < BN / BK; ++c)
              tma_load_3d(sB + c * BK * ROW_BYTES, &tmB, full_bar(stage), n0 + c * BK, kb * BK, zb);
          } else {
            tma_load_3d(sB, &tmB, full_bar(stage), kb * BK, n0, zb);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(FMT, A_MN, B_MN, BN);
      // byte advance of the descriptor start address per UMMA_K step
      constexpr uint32_t A_KSTEP = A_MN ? Cfg::UMMA_K * ROW_BYTES : 32;
      constexpr uint32_t B_KSTEP = B_MN ? Cfg::UMMA_K * ROW_BYTES : 32;
      constexpr uint32_t A_LBO = A_MN ? BK * ROW_BYTES : 16;
      constexpr uint32_t B_LBO = B_MN ? BK * ROW_BYTES : 16;
      // fp32 MN-major operands must use the 32B-atom flavour of the 128B swizzle (4-row atoms)
      constexpr uint32_t A_LT = (TF32X3 && A_MN) ? 1u : 2u, B_LT = (TF32X3 && B_MN) ? 1u : 2u;
      constexpr uint32_t A_SBO = (TF32X3 && A_MN) ? 512u : 1024u, B_SBO = (TF32X3 && B_MN) ? 512u : 1024u;
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;  // accumulator hand-offs so far (one per K chunk)
      for (int64_t tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        const TileCoord tc = decode_tile(p, tile);
        const int kb0 = tc.slice * p.kb_per_slice;
        const int kb1 = min(p.kb_total, kb0 + p.kb_per_slice);
        for (int kc0 = kb0; kc0 < kb1; kc0 += p.kb_per_chunk, ++it) {
          const int kc1 = min(kb1, kc0 + p.kb_per_chunk);
          const int as = it & 1;
          const uint32_t aphase = (uint32_t)(it >> 1) & 1u;
          mbar_wait(tempty_bar(as), aphase ^ 1u);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + (uint32_t)(as * BN);
          for (int kb = kc0; kb < kc1; ++kb) {
            mbar_wait(TF32X3 ? split_bar(stage) : full_bar(stage), phase);
            tc_fence_after();
            const uint32_t sA = smem_base + stage * Cfg::STAGE_BYTES;
            const uint32_t sB = sA + Cfg::A_BYTES;
#pragma unroll
            for (int k = 0; k < BK / Cfg::UMMA_K; ++k) {
              const uint64_t adesc = make_smem_desc(sA + k * A_KSTEP, A_LBO, A_SBO, A_LT);
              const uint64_t bdesc = make_smem_desc(sB + k * B_KSTEP, B_LBO, B_SBO, B_LT);
              const uint32_t acc = (kb > kc0 || k > 0) ? 1u : 0u;
              if (TF32X3) {
                const uint64_t adesc_lo = make_smem_desc(sA + Cfg::RAW_BYTES + k * A_KSTEP, A_LBO, A_SBO, A_LT);
                const uint64_t bdesc_lo = make_smem_desc(sB + Cfg::RAW_BYTES + k * B_KSTEP, B_LBO, B_SBO, B_LT);
                tc_mma<true>(d_tmem, adesc_lo, bdesc, idesc, acc);
                tc_mma<true>(d_tmem, adesc, bdesc_lo, idesc, 1u);
                tc_mma<true>(d_tmem, adesc, bdesc, idesc, 1u);
              } else {
                tc_mma<false>(d_tmem, adesc, bdesc, idesc, acc);
              }
            }
            tc_commit(empty_bar(stage));  // smem slot reusable once these MMAs retire
            if (++stage == STAGES) { stage = 0; phase ^= 1u; }
          }
          tc_commit(tfull_bar(as));  // chunk accumulator complete
        }
      }
    }
    __syncwarp();
  } else if (warp >= 4 && warp < 8) {
    // ===================== epilogue =====================
    const int ew = warp & 3;  // TMEM lane quarter this warp may read
    const bool beta0 = (p.beta == 0.0f);
    int it = 0;
    for (int64_t tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      const TileCoord tc = decode_tile(p, tile);
      const int kb0 = tc.slice * p.kb_per_slice;
      const int kb1 = min(p.kb_total, kb0 + p.kb_per_slice);
      const int64_t m = (int64_t)tc.mt * BM + ew * 32 + lane;
      const int64_t n0 = (int64_t)tc.nt * BN;
      const bool m_ok = m < p.M;
      for (int kc0 = kb0; kc0 < kb1; kc0 += p.kb_per_chunk, ++it) {
        const bool first = (kc0 == kb0), last = (kc0 + p.kb_per_chunk >= kb1);
        const int as = it & 1;
        const uint32_t aphase = (uint32_t)(it >> 1) & 1u;
        mbar_wait(tfull_bar(as), aphase);
        tc_fence_after();
        const uint32_t t_lane = tmem_base + ((uint32_t)(ew * 32) << 16);
        const uint32_t t_row = t_lane + (uint32_t)(as * BN);
        const uint32_t t_sum = t_lane + (uint32_t)Cfg::RSUM_COL;
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 32) {
          if (n0 + c0 >= p.N) break;  // warp-uniform
          uint32_t v[32];
          tmem_ld_32x32(t_row + c0, v);
          if (TF32X3 && !first) {
            uint32_t r[32];
            tmem_ld_32x32(t_sum + c0, r);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(r[j]));
          } else {
            tmem_ld_wait();
          }
          if (TF32X3 && !last) {
            tmem_st_32x32(t_sum + c0, v);
            continue;
          }
          if (p.slices > 1) {
            float* ws = p.ws + (((int64_t)tc.b * p.slices + tc.slice) * p.N) * p.M;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const int64_t n = n0 + c0 + j;
              if (m_ok && n < p.N) ws[n * p.M + m] = __uint_as_float(v[j]);
            }
          } else {
            TOut* C = reinterpret_cast<TOut*>(p.C) + (int64_t)tc.b * p.sc;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const int64_t n = n0 + c0 + j;
              if (m_ok && n < p.N) {
                TOut* dst = C + m + n * p.ldc;
                float r = p.alpha * __uint_as_float(v[j]);
                if (!beta0) r += p.beta * OutCvt<TOut>::load(dst);
                OutCvt<TOut>::store(dst, r);
              }
            }
          }
        }
        if (TF32X3 && !last) tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tempty_bar(as));
      }
    }
  } else if (TF32X3 && warp >= 8) {
    // ===================== fp32 -> (hi, lo) tf32 splitters =====================
    const int st = threadIdx.x - 256;  // 0..127
    int stage = 0;
    uint32_t phase = 0;
    constexpr int VEC_PER_STAGE = Cfg::RAW_BYTES / 16;
    for (int64_t tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      const TileCoord tc = decode_tile(p, tile);
      const int kb0 = tc.slice * p.kb_per_slice;
      const int kb1 = min(p.kb_total, kb0 + p.kb_per_slice);
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(full_bar(stage), phase);
        // elementwise, so the swizzled placement is preserved: lo tile = raw tile + RAW_BYTES
        float4* raw = reinterpret_cast<float4*>(smem_raw + (smem_base - smem_u32(smem_raw)) +
                                                stage * Cfg::STAGE_BYTES);
        float4* lo = raw + VEC_PER_STAGE;
#pragma unroll 4
        for (int i = st; i < VEC_PER_STAGE; i += Cfg::NUM_SPLIT_THREADS) {
          const float4 a = raw[i];
          float4 h, l;
          auto split = [](float x, float& hi, float& lo_) {
            uint32_t hb;
            asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hb) : "f"(x));
            hi = __uint_as_float(hb);
            const float d = (fabsf(hi) == INFINITY) ? 0.0f : x - hi;
            uint32_t lb;
            asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(lb) : "f"(d));
            lo_ = __uint_as_float(lb);
          };
          split(a.x, h.x, l.x);
          split(a.y, h.y, l.y);
          split(a.z, h.z, l.z);
          split(a.w, h.w, l.w);
          raw[i] = h;
          lo[i] = l;
        }
        fence_proxy_async();
        mbar_arrive(split_bar(stage));
        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
}

// ---- host side ---------------------------------------------------------------------------
PFN_cuTensorMapEncodeTiled_v12000 get_encode_fn() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(f);
  });
  return fn;
}

// Tensor map of one operand.  kcontig: stored with K contiguous (K-major), else MN contiguous.
bool make_operand_map(CUtensorMap* out, int es, CUtensorMapDataType dt, const void* ptr, int64_t mn,
                      int64_t k, int64_t ld, int64_t batch, int64_t stride, bool kcontig, int box_mn) {
  auto fn = get_encode_fn();
  if (!fn) return false;
  const int bk = ROW_BYTES / es;
  const bool batched = batch > 1 && stride > 0;
  cuuint64_t dims[3] = {(cuuint64_t)(kcontig ? k : mn), (cuuint64_t)(kcontig ? mn : k),
                        (cuuint64_t)(batched ? batch : 1)};
  cuuint64_t strides[2] = {(cuuint64_t)ld * es, (cuuint64_t)(batched ? stride : ld) * es};
  cuuint32_t box[3] = {(cuuint32_t)bk, (cuuint32_t)(kcontig ? box_mn : bk), 1};
  cuuint32_t estr[3] = {1, 1, 1};
  // fp32 MN-major tiles: tcgen05 only accepts the 32B-atom 128B swizzle for them
  const CUtensorMapSwizzle sw = (es == 4 && !kcontig) ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B;
  CUresult r = fn(out, dt, 3, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

template <typename TIn, typename TOut, int BN, int STAGES, bool A_MN, bool B_MN>
int launch_inst(pbx_handle_t h, const CUtensorMap& tmA, const CUtensorMap& tmB, const TcParams& p) {
  using Cfg = TcCfg<sizeof(TIn), BN, STAGES>;
  auto kern = gemm_tc_kernel<TIn, TOut, BN, STAGES, A_MN, B_MN>;
  PBX_CUDA_CHECK(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
  const int64_t grid = p.total_tiles < h->sm_count ? p.total_tiles : h->sm_count;
  kern<<<(unsigned)grid, Cfg::NUM_THREADS, Cfg::SMEM_BYTES, h->stream>>>(tmA, tmB, p);
  h->launches++;
  PBX_CUDA_CHECK(h, cudaGetLastError());
  return PBX_OK;
}

template <typename TIn, typename TOut, int BN, int STAGES>
int launch_major(pbx_handle_t h, bool a_mn, bool b_mn, const CUtensorMap& tmA, const CUtensorMap& tmB,
                 const TcParams& p) {
  if (a_mn) {
    return b_mn ? launch_inst<TIn, TOut, BN, STAGES, true, true>(h, tmA, tmB, p)
                : launch_inst<TIn, TOut, BN, STAGES, true, false>(h, tmA, tmB, p);
  }
  return b_mn ? launch_inst<TIn, TOut, BN, STAGES, false, true>(h, tmA, tmB, p)
              : launch_inst<TIn, TOut, BN, STAGES, false, false>(h, tmA, tmB, p);
}

}  // namespace

bool pbx_tcgen05_eligible(pbx_handle_t h, const PbxGemmCall& c) {
  if (c.dtype == PBX_F64) return false;
  const int64_t es = (int64_t)pbx_in_size(c.dtype);
  auto ok = [&](const void* p, int64_t ld, int64_t stride) {
    return ((uintptr_t)p % 16 == 0) && ((ld * es) % 16 == 0) && ((stride * es) % 16 == 0) &&
           (ld * es < ((int64_t)1 << 40)) && (stride * es < ((int64_t)1 << 40));
  };
  if (!ok(c.A, c.lda, c.sa) || !ok(c.B, c.ldb, c.sb)) return false;
  if (c.m >= ((int64_t)1 << 31) || c.n >= ((int64_t)1 << 31) || c.k >= ((int64_t)1 << 31) ||
      c.batch >= ((int64_t)1 << 31))
    return false;
  return get_encode_fn() != nullptr;
}

int pbx_launch_tcgen05(pbx_handle_t h, const PbxGemmCall& c, int slices) {
  const int es = (int)pbx_in_size(c.dtype);
  const bool f32 = (c.dtype == PBX_F32);
  const int bk = ROW_BYTES / es;
  // tile width: 256 columns when the problem is wide enough to keep 148 SMs busy with it
  int bn = 128;
  if (!f32) {
    const int64_t tiles256 = ((c.m + BM - 1) / BM) * ((c.n + 255) / 256) * c.batch * slices;
    if (c.n > 128 && tiles256 >= h->sm_count) bn = 256;
  }
  const bool a_mn = !c.ta, b_mn = c.tb;
  CUtensorMapDataType dt = f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32
                               : ((c.dtype == PBX_F16 || c.dtype == PBX_F16_F32) ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16
                                                                                : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16);
  CUtensorMap tmA, tmB;
  if (!make_operand_map(&tmA, es, dt, c.A, c.m, c.k, c.lda, c.batch, c.sa, !a_mn, BM) ||
      !make_operand_map(&tmB, es, dt, c.B, c.n, c.k, c.ldb, c.batch, c.sb, !b_mn, bn)) {
    h->last_error = "cuTensorMapEncodeTiled failed";
    return PBX_ERR_CUDA;
  }
  TcParams p;
  p.C = c.C; p.ws = (float*)h->ws;
  p.M = c.m; p.N = c.n; p.K = c.k; p.ldc = c.ldc; p.sc = c.sc;
  p.alpha = (float)c.alpha; p.beta = (float)c.beta;
  p.batch = (int)c.batch; p.slices = slices;
  p.m_tiles = (int)((c.m + BM - 1) / BM);
  p.n_tiles = (int)((c.n + bn - 1) / bn);
  p.group_m = 8;
  p.kb_total = (int)((c.k + bk - 1) / bk);
  p.kb_per_slice = (p.kb_total + slices - 1) / slices;
  // fp32: 16 blocks x 32 = 512 k per tensor-core accumulation chain (~4e-6 relative bias per chunk)
  p.kb_per_chunk = f32 ? 16 : (1 << 30);
  p.a_batched = (c.batch > 1 && c.sa > 0) ? 1 : 0;
  p.b_batched = (c.batch > 1 && c.sb > 0) ? 1 : 0;
  p.total_tiles = (int64_t)p.m_tiles * p.n_tiles * c.batch * slices;

  switch (c.dtype) {
    case PBX_F32:
      return launch_major<float, float, 128, 3>(h, a_mn, b_mn, tmA, tmB, p);
    case PBX_F16:
      return bn == 256 ? launch_major<__half, __half, 256, 4>(h, a_mn, b_mn, tmA, tmB, p)
                       : launch_major<__half, __half, 128, 6>(h, a_mn, b_mn, tmA, tmB, p);
    case PBX_F16_F32:
      return bn == 256 ? launch_major<__half, float, 256, 4>(h, a_mn, b_mn, tmA, tmB, p)
                       : launch_major<__half, float, 128, 6>(h, a_mn, b_mn, tmA, tmB, p);
    case PBX_BF16:
      return bn == 256 ? launch_major<__nv_bfloat16, __nv_bfloat16, 256, 4>(h, a_mn, b_mn, tmA, tmB, p)
                       : launch_major<__nv_bfloat16, __nv_bfloat16, 128, 6>(h, a_mn, b_mn, tmA, tmB, p);
    case PBX_BF16_F32:
      return bn == 256 ? launch_major<__nv_bfloat16, float, 256, 4>(h, a_mn, b_mn, tmA, tmB, p)
                       : launch_major<__nv_bfloat16, float, 128, 6>(h, a_mn, b_mn, tmA, tmB, p);
  }
  return PBX_ERR_INVALID_ARG;
}
