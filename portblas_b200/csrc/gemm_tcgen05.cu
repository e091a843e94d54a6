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
// Skinny M (M <= 64 < N): the roles of the operands are swapped -- the kernel computes
// C^T = op(B)^T op(A)^T, so N rides on the 128 TMEM lanes and M on a 64-wide MMA N dimension (half the
// tensor-pipe time of a 128-row tile that is mostly zero padding).  Only the descriptors and the
// epilogue addressing change (TRANS_OUT): a thread then owns 32 CONSECUTIVE elements of a C column.
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
  int raw_hi;        // fp32: 1 = feed raw fp32 as the hi operand (hardware truncates to tf32)
  int a_batched, b_batched;
  int tma_store;     // 16-bit C through shared memory + TMA store (needs beta == 0, aligned C, no split-K)
  int c_vec;         // TRANS_OUT: C rows of 32 elements may be stored as 16-byte vectors
  int n_extra;       // multicast GEMM: the epilogue also stores the tile into these copies of C (peer GPUs' memory,
  void* Cx[7];       // mapped through CUDA IPC; same ldc / batch stride as C) -- the gather rides on the GEMM's stores
  int64_t total_tiles;
};

template <typename T> struct OutCvt;
template <> struct OutCvt<float> {
  __device__ static float load(const float* p) { return *p; }
  __device__ static void store(float* p, float v) { *p = v; }
  __device__ static uint32_t pack2(float lo, float) { return __float_as_uint(lo); }   // unused for 32-bit outputs
};
template <> struct OutCvt<__half> {
  __device__ static float load(const __half* p) { return __half2float(*p); }
  __device__ static void store(__half* p, float v) { *p = __float2half_rn(v); }
  __device__ static uint32_t pack2(float lo, float hi) {
    const __half2 h = __floats2half2_rn(lo, hi);
    return *reinterpret_cast<const uint32_t*>(&h);
  }
};
template <> struct OutCvt<__nv_bfloat16> {
  __device__ static float load(const __nv_bfloat16* p) { return __bfloat162float(*p); }
  __device__ static void store(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }
  __device__ static uint32_t pack2(float lo, float hi) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<const uint32_t*>(&h);
  }
};

constexpr int BM = 128;         // rows of D held by one CTA (TMEM lanes)
constexpr int ROW_BYTES = 128;  // one swizzle row

// CG = 1: one CTA computes a 128 x BN tile.  CG = 2: a CTA pair (2-CTA cluster, cta_group::2)
// computes a 256 x BN tile; each CTA stages its own 128 rows of A and BN/2 rows of B, so the
// shared-memory read rate per SM halves for the same MMA rate.
// PRE (fp32 only) selects how the 3xTF32 operands are produced:
//   0  in-kernel split: splitter warps derive the lo tiles from the staged raw tiles
//   1  pre-split: the lo halves of both operands were computed by a pre-pass into global memory and arrive by TMA
//      like the raw tiles, so the CTA has no splitter warps and its shared memory carries no split traffic
//   2  single tf32 product (no lo halves at all): the reduced-precision mode the reference reaches with
//      SB_ENABLE_JOINT_MATRIX=1 (float storage, 10-bit-mantissa fragments, fp32 accumulate,
//      src/interface/blas3/backend/nvidia_gpu.hpp:67-110); stages hold raw tiles only, so the ring is twice as deep
//   3  tf32 + 2 x bf16 (EXPERIMENTAL, PBX_F32_SPLIT16=1, not yet run on a GPU): A_hi*B_hi as one tf32 MMA on the raw
//      tiles, the two cross terms as kind::f16 MMAs on bf16 copies made by a pre-pass (bf16(a) and bf16(a - trunc_tf32(a))):
//      the lo halves are 2^-11 of the operand, so 8 bits of them keep the product inside the 1e-5 budget
//      (tools/split_emulation.py: <= 2e-7 of sum|a||b|), and bf16 MMAs run at twice the tf32 rate -- two tf32-MMA
//      times per k-step instead of three.  The bf16 tiles are 32 k wide like the fp32 ones (64-byte rows, 64B swizzle)
//      and take the place of the lo tiles in the stage: [A raw | B raw | A16 hi | A16 lo | B16 hi | B16 lo]
template <int ES, int BN, int STAGES, int CG, int OS = 4, int PRE = 0>
struct TcCfg {
  static constexpr bool TF32X3 = (ES == 4);
  static constexpr int BK = ROW_BYTES / ES;        // 64 (16-bit) or 32 (fp32) elements
  static constexpr int UMMA_K = 32 / ES;           // 16 or 8
  static constexpr int TILE_M = BM * CG;
  static constexpr int BN_CTA = BN / CG;           // rows of the B operand staged by this CTA
  static constexpr int A_BYTES = BM * ROW_BYTES;   // 16 KiB
  static constexpr int B_BYTES = BN_CTA * ROW_BYTES;
  static constexpr int RAW_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGE_BYTES = RAW_BYTES * ((TF32X3 && PRE != 2) ? 2 : 1);  // + lo copies
  static constexpr int BAR_BYTES = 256;
  static constexpr int EPI_WARPS = 8;
  // 16-bit outputs: every epilogue warp owns two 32x32 staging tiles (column-major, rows contiguous)
  // that it hands to TMA stores, so C leaves the SM as bulk writes instead of 2-byte stores
  static constexpr int EPI_TILE_BYTES = 32 * 32 * OS;
  static constexpr int EPI_BYTES = (OS == 2) ? EPI_WARPS * 2 * EPI_TILE_BYTES : 0;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EPI_BYTES + BAR_BYTES + 1024;  // + align slack
  // TMEM: ACC_STAGES accumulators of BN columns (+ for fp32 a BN-column running sum)
  static constexpr int ACC_STAGES = TF32X3 ? ((3 * BN <= 512) ? 2 : 1) : 2;
  static constexpr int RSUM_COL = ACC_STAGES * BN;
  static constexpr int TMEM_COLS = TF32X3 ? 512 : 2 * BN;
  static constexpr int SPLIT_WARPS = (TF32X3 && PRE == 0) ? 4 : 0;
  static constexpr int TMA_BYTES = RAW_BYTES * ((TF32X3 && (PRE == 1 || PRE == 3)) ? 2 : 1);   // bytes one CTA's producer lands per stage
  static constexpr int NUM_THREADS = 32 * (4 + EPI_WARPS + SPLIT_WARPS);
  static constexpr int NUM_SPLIT_THREADS = 32 * SPLIT_WARPS;
  static_assert(BN % (32 * 2) == 0 && BN <= 256 && (BN_CTA % 8) == 0, "tile width");
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

// M, N, ldc in TcParams are the KERNEL's view: with TRANS_OUT the caller passed (N, M) and the kernel's
// D(row r, column c) is C(c, r), i.e. element address c + r*ldc instead of r + c*ldc.
template <typename TIn, typename TOut, int BN, int STAGES, bool A_MN, bool B_MN, int CG, bool TRANS_OUT, int PRE>
__global__ void __launch_bounds__(TcCfg<sizeof(TIn), BN, STAGES, CG, sizeof(TOut), PRE>::NUM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmAlo,
               const __grid_constant__ CUtensorMap tmBlo, const TcParams p) {
  using Cfg = TcCfg<sizeof(TIn), BN, STAGES, CG, sizeof(TOut), PRE>;
  static_assert(PRE == 0 || Cfg::TF32X3, "pre-split / single-tf32 modes exist for fp32 only");
  constexpr bool TF32X3 = Cfg::TF32X3;
  constexpr int BK = Cfg::BK;
  constexpr int ACC_STAGES = Cfg::ACC_STAGES;
  constexpr uint32_t FMT = TF32X3 ? 2u : (std::is_same<TIn, __nv_bfloat16>::value ? 1u : 0u);

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t epi_base = smem_base + STAGES * Cfg::STAGE_BYTES;
  const uint32_t bar_base = epi_base + Cfg::EPI_BYTES;
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
  const uint32_t rank = (CG == 2) ? cluster_ctarank() : 0u;   // 0 = leader (issues the MMAs)
  const int64_t group = blockIdx.x / CG, num_groups = gridDim.x / CG;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (Cfg::EPI_BYTES > 0 && p.tma_store) tma_prefetch_desc(&tmC);
    if (PRE == 1 || PRE == 3) { tma_prefetch_desc(&tmAlo); tma_prefetch_desc(&tmBlo); }
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
      mbar_init(split_bar(s), Cfg::NUM_SPLIT_THREADS > 0 ? CG * Cfg::NUM_SPLIT_THREADS : 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull_bar(s), 1);
      mbar_init(tempty_bar(s), CG * Cfg::EPI_WARPS);
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    if (CG == 2) { tmem_alloc_2sm(tmem_ptr_smem, Cfg::TMEM_COLS); tmem_relinquish_2sm(); }
    else { tmem_alloc(tmem_ptr_smem, Cfg::TMEM_COLS); tmem_relinquish(); }
  }
  tc_fence_before();
  if (CG == 2) cluster_sync_all(); else __syncthreads();   // peer barriers must exist before any remote arrive
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_generic;

  if (warp == 0) {
    // ===================== TMA producer (every CTA loads its own A rows and B rows) =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      // pair mode without splitters (16-bit, pre-split fp32): both CTAs credit the leader's full barrier (the MMA
      // issuer waits there).  fp32 with in-kernel split: each CTA's splitter warps wait on their OWN full barrier.
      constexpr bool kLeaderFull = (CG == 2) && (!TF32X3 || PRE != 0);
      for (int64_t tile = group; tile < p.total_tiles; tile += num_groups) {
        const TileCoord tc = decode_tile(p, tile);
        const int m0 = tc.mt * Cfg::TILE_M + (int)rank * BM;
        const int n0 = tc.nt * BN + (int)rank * Cfg::BN_CTA;
        const int kb0 = tc.slice * p.kb_per_slice;
        const int kb1 = min(p.kb_total, kb0 + p.kb_per_slice);
        const int za = p.a_batched ? tc.b : 0, zb = p.b_batched ? tc.b : 0;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1u);
          const uint32_t sA = smem_base + stage * Cfg::STAGE_BYTES;
          const uint32_t sB = sA + Cfg::A_BYTES;
          const uint32_t fb = full_bar(stage);
          if (kLeaderFull) { if (rank == 0) mbar_expect_tx(fb, 2 * Cfg::TMA_BYTES); }
          else mbar_expect_tx(fb, Cfg::TMA_BYTES);
          auto load = [&](uint32_t dst, const CUtensorMap* tm, int x, int y, int z) {
            if (kLeaderFull) tma_load_3d_2sm(dst, tm, fb, x, y, z);
            else tma_load_3d(dst, tm, fb, x, y, z);
          };
          if (A_MN) {
#pragma unroll
            for (int c = 0; c < BM / BK; ++c) load(sA + c * BK * ROW_BYTES, &tmA, m0 + c * BK, kb * BK, za);
          } else {
            load(sA, &tmA, kb * BK, m0, za);
          }
          if (B_MN) {
#pragma unroll
            for (int c = 0; c < Cfg::BN_CTA / BK; ++c) load(sB + c * BK * ROW_BYTES, &tmB, n0 + c * BK, kb * BK, zb);
          } else {
            load(sB, &tmB, kb * BK, n0, zb);
          }
          if (PRE == 1) {   // lo tiles: same boxes of the pre-split copies, placed RAW_BYTES further
            const uint32_t sAl = sA + Cfg::RAW_BYTES, sBl = sB + Cfg::RAW_BYTES;
            if (A_MN) {
#pragma unroll
              for (int c = 0; c < BM / BK; ++c) load(sAl + c * BK * ROW_BYTES, &tmAlo, m0 + c * BK, kb * BK, za);
            } else {
              load(sAl, &tmAlo, kb * BK, m0, za);
            }
            if (B_MN) {
#pragma unroll
              for (int c = 0; c < Cfg::BN_CTA / BK; ++c)
                load(sBl + c * BK * ROW_BYTES, &tmBlo, n0 + c * BK, kb * BK, zb);
            } else {
              load(sBl, &tmBlo, kb * BK, n0, zb);
            }
          }
          if (PRE == 3) {
            // bf16 copies (tmAlo = A's, tmBlo = B's): z = which * copies + batch entry, which = 0 (hi) / 1 (lo);
            // K-major: one box of 32 k x rows; MN-major: boxes of 32 mn x 32 k (2 KiB each, 64-byte rows)
            const uint32_t s16 = sA + Cfg::RAW_BYTES;
            const int ca = p.a_batched ? p.batch : 1, cb = p.b_batched ? p.batch : 1;
#pragma unroll
            for (int w = 0; w < 2; ++w) {
              const uint32_t dA = s16 + w * (Cfg::A_BYTES / 2);
              const uint32_t dB = s16 + Cfg::A_BYTES + w * (Cfg::B_BYTES / 2);
              if (A_MN) {
#pragma unroll
                for (int c = 0; c < BM / 32; ++c) load(dA + c * 2048, &tmAlo, m0 + c * 32, kb * BK, w * ca + za);
              } else {
                load(dA, &tmAlo, kb * BK, m0, w * ca + za);
              }
              if (B_MN) {
#pragma unroll
                for (int c = 0; c < Cfg::BN_CTA / 32; ++c) load(dB + c * 2048, &tmBlo, n0 + c * 32, kb * BK, w * cb + zb);
              } else {
                load(dB, &tmBlo, kb * BK, n0, w * cb + zb);
              }
            }
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================== MMA issuer (one lane; leader CTA only in pair mode) =====================
    if (lane == 0 && rank == 0) {
      constexpr uint32_t idesc = make_idesc(FMT, A_MN, B_MN, BN, Cfg::TILE_M);
      // byte advance of the descriptor start address per UMMA_K step
      constexpr uint32_t A_KSTEP = A_MN ? Cfg::UMMA_K * ROW_BYTES : 32;
      constexpr uint32_t B_KSTEP = B_MN ? Cfg::UMMA_K * ROW_BYTES : 32;
      constexpr uint32_t A_LBO = A_MN ? BK * ROW_BYTES : 16;
      constexpr uint32_t B_LBO = B_MN ? BK * ROW_BYTES : 16;
      // fp32 MN-major operands must use the 32B-atom flavour of the 128B swizzle (4-row atoms)
      constexpr uint32_t A_LT = (TF32X3 && A_MN) ? 1u : 2u, B_LT = (TF32X3 && B_MN) ? 1u : 2u;
      constexpr uint32_t A_SBO = (TF32X3 && A_MN) ? 512u : 1024u, B_SBO = (TF32X3 && B_MN) ? 512u : 1024u;
      auto mma = [&](uint32_t d, uint64_t a, uint64_t b, uint32_t acc) {
        if (CG == 2) tc_mma_2sm<TF32X3>(d, a, b, idesc, acc); else tc_mma<TF32X3>(d, a, b, idesc, acc);
      };
      auto commit = [&](uint32_t bar) { if (CG == 2) tc_commit_2sm(bar); else tc_commit(bar); };
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;  // accumulator hand-offs so far (one per K chunk)
      for (int64_t tile = group; tile < p.total_tiles; tile += num_groups) {
        const TileCoord tc = decode_tile(p, tile);
        const int kb0 = tc.slice * p.kb_per_slice;
        const int kb1 = min(p.kb_total, kb0 + p.kb_per_slice);
        for (int kc0 = kb0; kc0 < kb1; kc0 += p.kb_per_chunk, ++it) {
          const int kc1 = min(kb1, kc0 + p.kb_per_chunk);
          const int as = it % ACC_STAGES;
          const uint32_t aphase = (uint32_t)(it / ACC_STAGES) & 1u;
          mbar_wait(tempty_bar(as), aphase ^ 1u);   // only TMEM state is handed over (tcgen05 fences order it)
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + (uint32_t)(as * BN);
          for (int kb = kc0; kb < kc1; ++kb) {
            constexpr bool kSplit = TF32X3 && PRE == 0;
            const uint32_t ready = kSplit ? split_bar(stage) : full_bar(stage);
            // fp32 pair mode with in-kernel split: the peer's splitter warps wrote shared memory with ordinary stores
            if (CG == 2 && kSplit) mbar_wait_cluster(ready, phase); else mbar_wait(ready, phase);
            tc_fence_after();
            const uint32_t sA = smem_base + stage * Cfg::STAGE_BYTES;
            const uint32_t sB = sA + Cfg::A_BYTES;
#pragma unroll
            for (int k = 0; k < BK / Cfg::UMMA_K; ++k) {
              const uint64_t adesc = make_smem_desc(sA + k * A_KSTEP, A_LBO, A_SBO, A_LT);
              const uint64_t bdesc = make_smem_desc(sB + k * B_KSTEP, B_LBO, B_SBO, B_LT);
              const uint32_t acc = (kb > kc0 || k > 0) ? 1u : 0u;
              if (PRE == 3) {
                mma(d_tmem, adesc, bdesc, acc);   // hi * hi (tf32 on the raw tiles); the cross terms follow per k-block
              } else if (TF32X3 && PRE != 2) {
                const uint64_t adesc_lo = make_smem_desc(sA + Cfg::RAW_BYTES + k * A_KSTEP, A_LBO, A_SBO, A_LT);
                const uint64_t bdesc_lo = make_smem_desc(sB + Cfg::RAW_BYTES + k * B_KSTEP, B_LBO, B_SBO, B_LT);
                mma(d_tmem, adesc_lo, bdesc, acc);
                mma(d_tmem, adesc, bdesc_lo, 1u);
                mma(d_tmem, adesc, bdesc, 1u);
              } else {
                mma(d_tmem, adesc, bdesc, acc);
              }
            }
            if (PRE == 3) {
              // cross terms on the bf16 tiles: 64-byte rows (K-major) / 64-byte mn chunks (MN-major), 64B swizzle
              // (layout type 4): SBO = 8 rows x 64 B; MN-major LBO = one 32 x 32 box; UMMA_K = 16
              constexpr uint32_t idesc16 = make_idesc(1u, A_MN, B_MN, BN, Cfg::TILE_M);
              constexpr uint32_t A16_KSTEP = A_MN ? 16 * 64 : 32, B16_KSTEP = B_MN ? 16 * 64 : 32;
              constexpr uint32_t A16_LBO = A_MN ? 2048 : 16, B16_LBO = B_MN ? 2048 : 16;
              const uint32_t sA16 = sA + Cfg::RAW_BYTES, sB16 = sA16 + Cfg::A_BYTES;
#pragma unroll
              for (int k = 0; k < BK / 16; ++k) {
                const uint64_t a_hi = make_smem_desc(sA16 + k * A16_KSTEP, A16_LBO, 512u, 4u);
                const uint64_t a_lo = make_smem_desc(sA16 + Cfg::A_BYTES / 2 + k * A16_KSTEP, A16_LBO, 512u, 4u);
                const uint64_t b_hi = make_smem_desc(sB16 + k * B16_KSTEP, B16_LBO, 512u, 4u);
                const uint64_t b_lo = make_smem_desc(sB16 + Cfg::B_BYTES / 2 + k * B16_KSTEP, B16_LBO, 512u, 4u);
                if (CG == 2) {
                  tc_mma_2sm<false>(d_tmem, a_lo, b_hi, idesc16, 1u);
                  tc_mma_2sm<false>(d_tmem, a_hi, b_lo, idesc16, 1u);
                } else {
                  tc_mma<false>(d_tmem, a_lo, b_hi, idesc16, 1u);
                  tc_mma<false>(d_tmem, a_hi, b_lo, idesc16, 1u);
                }
              }
            }
            commit(empty_bar(stage));  // smem slot (both CTAs) reusable once these MMAs retire
            if (++stage == STAGES) { stage = 0; phase ^= 1u; }
          }
          commit(tfull_bar(as));  // chunk accumulator complete (both CTAs' epilogues)
        }
      }
    }
    __syncwarp();
  } else if (warp >= 4 && warp < 4 + Cfg::EPI_WARPS) {
    // ===================== epilogue: 8 warps = 4 TMEM lane quarters x 2 column halves =====================
    const int ew = warp & 3;           // TMEM lane quarter this warp may read
    const int ch = (warp - 4) >> 2;    // column half
    constexpr int COLS_PER_WARP = BN / 2;
    constexpr bool OUT16 = (sizeof(TOut) == 2);
    const bool beta0 = (p.beta == 0.0f);
    const bool tma_store = OUT16 && p.tma_store;
    const uint32_t tempty_leader0 = (CG == 2) ? map_to_cta(tempty_bar(0), 0) : tempty_bar(0);
    // staging tiles of this warp (16-bit outputs): [buffer][column][32 rows]
    const uint32_t stage_u32 = epi_base + (uint32_t)((warp - 4) * 2 * Cfg::EPI_TILE_BYTES);
    TOut* stage_ptr = reinterpret_cast<TOut*>(smem_raw + (stage_u32 - smem_u32(smem_raw)));
    uint32_t store_blk = 0;
    int it = 0;
    for (int64_t tile = group; tile < p.total_tiles; tile += num_groups) {
      const TileCoord tc = decode_tile(p, tile);
      const int kb0 = tc.slice * p.kb_per_slice;
      const int kb1 = min(p.kb_total, kb0 + p.kb_per_slice);
      const int64_t m_warp = (int64_t)tc.mt * Cfg::TILE_M + rank * BM + ew * 32;   // first row of this warp
      const int64_t m = m_warp + lane;
      const int64_t n0 = (int64_t)tc.nt * BN + ch * COLS_PER_WARP;
      const bool m_ok = m < p.M;
      const bool rows_full = (m_warp + 32 <= p.M);   // warp-uniform
      for (int kc0 = kb0; kc0 < kb1; kc0 += p.kb_per_chunk, ++it) {
        const bool first = (kc0 == kb0), last = (kc0 + p.kb_per_chunk >= kb1);
        const int as = it % ACC_STAGES;
        const uint32_t aphase = (uint32_t)(it / ACC_STAGES) & 1u;
        mbar_wait(tfull_bar(as), aphase);
        tc_fence_after();
        const uint32_t t_lane = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(ch * COLS_PER_WARP);
        const uint32_t t_row = t_lane + (uint32_t)(as * BN);
        const uint32_t t_sum = t_lane + (uint32_t)Cfg::RSUM_COL;
#pragma unroll 1
        for (int c0 = 0; c0 < COLS_PER_WARP; c0 += 32) {
          if (n0 + c0 >= p.N || m_warp >= p.M) break;  // warp-uniform
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
          const bool full = rows_full && (n0 + c0 + 32 <= p.N);   // warp-uniform: no predicate per element
          if (TRANS_OUT) {
            // kernel row (lane) = column of C, kernel columns = 32 consecutive rows of C
            const int64_t cc = n0 + c0;
            if (p.slices > 1) {
              float* ws = p.ws + (((int64_t)tc.b * p.slices + tc.slice) * p.M + m) * p.N + cc;
              if (full && (p.N & 3) == 0) {
#pragma unroll
                for (int j = 0; j < 32; j += 4)
                  *reinterpret_cast<float4*>(ws + j) = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]),
                                                                   __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
              } else {
#pragma unroll
                for (int j = 0; j < 32; ++j)
                  if (m_ok && cc + j < p.N) ws[j] = __uint_as_float(v[j]);
              }
            } else if (OUT16 && tma_store) {
              const uint32_t buf = store_blk & 1u;
              ++store_blk;
              if (lane == 0) bulk_wait_read<1>();
              __syncwarp();
              TOut* st = stage_ptr + buf * (Cfg::EPI_TILE_BYTES / (int)sizeof(TOut)) + lane * 32;
#pragma unroll
              for (int j = 0; j < 32; ++j) OutCvt<TOut>::store(st + j, p.alpha * __uint_as_float(v[j]));
              fence_proxy_async();
              __syncwarp();
              if (lane == 0) {
                tma_store_3d(&tmC, stage_u32 + buf * Cfg::EPI_TILE_BYTES, (int)cc, (int)m_warp, tc.b);
                bulk_commit();
              }
            } else {
              TOut* dst = reinterpret_cast<TOut*>(p.C) + (int64_t)tc.b * p.sc + m * p.ldc + cc;
              if (full && p.c_vec && sizeof(TOut) == 4) {
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                  float4 o = make_float4(p.alpha * __uint_as_float(v[j]), p.alpha * __uint_as_float(v[j + 1]),
                                         p.alpha * __uint_as_float(v[j + 2]), p.alpha * __uint_as_float(v[j + 3]));
                  if (!beta0) {
                    const float4 ci = *reinterpret_cast<const float4*>(dst + j);
                    o.x += p.beta * ci.x; o.y += p.beta * ci.y; o.z += p.beta * ci.z; o.w += p.beta * ci.w;
                  }
                  *reinterpret_cast<float4*>(dst + j) = o;
                }
              } else {
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                  if (m_ok && cc + j < p.N) {
                    float r = p.alpha * __uint_as_float(v[j]);
                    if (!beta0) r += p.beta * OutCvt<TOut>::load(dst + j);
                    OutCvt<TOut>::store(dst + j, r);
                  }
                }
              }
            }
          } else if (p.slices > 1) {
            float* ws = p.ws + (((int64_t)tc.b * p.slices + tc.slice) * p.N + (n0 + c0)) * p.M + m;
            if (full) {
#pragma unroll
              for (int j = 0; j < 32; ++j) ws[(int64_t)j * p.M] = __uint_as_float(v[j]);
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (m_ok && n0 + c0 + j < p.N) ws[(int64_t)j * p.M] = __uint_as_float(v[j]);
            }
          } else if (OUT16 && tma_store) {
            // registers -> staging tile (a warp writes 64 contiguous bytes per column: conflict-free)
            // -> one TMA store of the 32x32 box; rows / columns outside C are clipped by the hardware
            const uint32_t buf = store_blk & 1u;
            ++store_blk;
            if (lane == 0) bulk_wait_read<1>();   // the store that last used this buffer has read it
            __syncwarp();
            // Two-byte stores put two lanes on every bank word (ncu: 40 % of the shared-memory wavefronts of a short-K
            // bf16 GEMM were conflicts).  Lanes 2p / 2p+1 swap one value per column pair instead, so the even lane
            // writes the packed word {row 2p, row 2p+1} of column j and the odd lane that of column j+1: 16
            // conflict-free 32-bit stores per lane instead of 32 two-byte ones.
            uint32_t* stw = reinterpret_cast<uint32_t*>(stage_ptr + buf * (Cfg::EPI_TILE_BYTES / (int)sizeof(TOut)));
            const int pr = lane >> 1;
            const bool odd = (lane & 1) != 0;
#pragma unroll
            for (int j = 0; j < 32; j += 2) {
              const float e = p.alpha * __uint_as_float(v[j]), o = p.alpha * __uint_as_float(v[j + 1]);
              const float recv = __shfl_xor_sync(0xffffffffu, odd ? e : o, 1);
              stw[(j + (odd ? 1 : 0)) * 16 + pr] = OutCvt<TOut>::pack2(odd ? recv : e, odd ? o : recv);
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
              tma_store_3d(&tmC, stage_u32 + buf * Cfg::EPI_TILE_BYTES, (int)m_warp, (int)(n0 + c0), tc.b);
              bulk_commit();
            }
          } else {
            const int64_t c_off = (int64_t)tc.b * p.sc + m + (n0 + c0) * p.ldc;
            TOut* dst = reinterpret_cast<TOut*>(p.C) + c_off;
            if (full) {
              if (beta0) {
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                  const float r = p.alpha * __uint_as_float(v[j]);
                  v[j] = __float_as_uint(r);
                  OutCvt<TOut>::store(dst + j * p.ldc, r);
                }
              } else {
#pragma unroll
                for (int j0 = 0; j0 < 32; j0 += 8) {   // 8 loads in flight, then 8 stores
                  float cin[8];
#pragma unroll
                  for (int j = 0; j < 8; ++j) cin[j] = OutCvt<TOut>::load(dst + (j0 + j) * p.ldc);
#pragma unroll
                  for (int j = 0; j < 8; ++j) {
                    const float r = p.alpha * __uint_as_float(v[j0 + j]) + p.beta * cin[j];
                    v[j0 + j] = __float_as_uint(r);
                    OutCvt<TOut>::store(dst + (j0 + j) * p.ldc, r);
                  }
                }
              }
              // multicast: the same 32 x 32 block goes to every other copy of C (peer memory over NVLink); the stores
              // are posted, so the transfer of this tile overlaps the mainloop of the next one
              for (int x = 0; x < p.n_extra; ++x) {
                TOut* dx = reinterpret_cast<TOut*>(p.Cx[x]) + c_off;
#pragma unroll
                for (int j = 0; j < 32; ++j) OutCvt<TOut>::store(dx + j * p.ldc, __uint_as_float(v[j]));
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                if (m_ok && n0 + c0 + j < p.N) {
                  float r = p.alpha * __uint_as_float(v[j]);
                  if (!beta0) r += p.beta * OutCvt<TOut>::load(dst + j * p.ldc);
                  OutCvt<TOut>::store(dst + j * p.ldc, r);
                  for (int x = 0; x < p.n_extra; ++x)
                    OutCvt<TOut>::store(reinterpret_cast<TOut*>(p.Cx[x]) + c_off + j * p.ldc, r);
                }
              }
            }
          }
        }
        if (TF32X3 && !last) tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (CG == 2) mbar_arrive_remote(tempty_leader0 + 8u * as); else mbar_arrive(tempty_bar(as));
        }
      }
    }
    if (OUT16 && tma_store && lane == 0) bulk_wait_read<0>();   // staging tiles must outlive their stores
    __syncwarp();
  } else if (TF32X3 && PRE == 0 && warp >= 4 + Cfg::EPI_WARPS) {
    // ===================== fp32 -> (hi, lo) tf32 splitters (each CTA splits what it staged) ==============
    const int st = threadIdx.x - 32 * (4 + Cfg::EPI_WARPS);
    int stage = 0;
    uint32_t phase = 0;
    constexpr int VEC_PER_STAGE = Cfg::RAW_BYTES / 16;
    const uint32_t split_leader0 = (CG == 2) ? map_to_cta(split_bar(0), 0) : split_bar(0);
    const bool raw_hi = p.raw_hi != 0;
    for (int64_t tile = group; tile < p.total_tiles; tile += num_groups) {
      const TileCoord tc = decode_tile(p, tile);
      const int kb0 = tc.slice * p.kb_per_slice;
      const int kb1 = min(p.kb_total, kb0 + p.kb_per_slice);
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(full_bar(stage), phase);
        // elementwise, so the swizzled placement is preserved: lo tile = raw tile + RAW_BYTES
        float4* raw = reinterpret_cast<float4*>(smem_raw + (smem_base - smem_u32(smem_raw)) +
                                                stage * Cfg::STAGE_BYTES);
        float4* lo = raw + VEC_PER_STAGE;
        if (raw_hi) {
          // hi operand = the raw fp32 tile (the tensor core reads only the tf32 bits, i.e. truncates);
          // lo = rn_tf32(a - trunc_tf32(a)).  Halves the shared-memory writes of the splitter.
          auto lo_of = [](float x) {
            const float d = x - __uint_as_float(__float_as_uint(x) & 0xffffe000u);
            uint32_t lb;
            asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(lb) : "f"(d == d ? d : 0.0f));  // inf - inf -> 0
            return __uint_as_float(lb);
          };
#pragma unroll 8
          for (int i = st; i < VEC_PER_STAGE; i += Cfg::NUM_SPLIT_THREADS) {
            const float4 a = raw[i];
            lo[i] = make_float4(lo_of(a.x), lo_of(a.y), lo_of(a.z), lo_of(a.w));
          }
        } else {
          auto split = [](float x, float& hi, float& lo_) {
            uint32_t hb;
            asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hb) : "f"(x));
            hi = __uint_as_float(hb);
            const float d = x - hi;
            uint32_t lb;
            asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(lb) : "f"(d == d ? d : 0.0f));
            lo_ = __uint_as_float(lb);
          };
#pragma unroll 4
          for (int i = st; i < VEC_PER_STAGE; i += Cfg::NUM_SPLIT_THREADS) {
            const float4 a = raw[i];
            float4 h, l;
            split(a.x, h.x, l.x);
            split(a.y, h.y, l.y);
            split(a.z, h.z, l.z);
            split(a.w, h.w, l.w);
            raw[i] = h;
            lo[i] = l;
          }
        }
        fence_proxy_async();
        if (CG == 2) mbar_arrive_cluster(split_leader0 + 8u * stage); else mbar_arrive(split_bar(stage));
        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
      }
    }
  }

  tc_fence_before();
  if (CG == 2) cluster_sync_all(); else __syncthreads();
  if (warp == 2) {
    if (CG == 2) tmem_dealloc_2sm(tmem_base, Cfg::TMEM_COLS); else tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
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

// Tensor map of the bf16 copies of one fp32 operand (PRE == 3): the hi copies of all batch entries, then the lo copies,
// along z (z = which * copies + b; zstride elements apart); 32-element (64-byte) rows, 64B swizzle.
bool make_split16_map(CUtensorMap* out, const void* ptr, int64_t mn, int64_t k, int64_t ld16, int64_t copies,
                      int64_t zstride, bool kcontig, int box_mn) {
  auto fn = get_encode_fn();
  if (!fn) return false;
  cuuint64_t dims[3] = {(cuuint64_t)(kcontig ? k : mn), (cuuint64_t)(kcontig ? mn : k), (cuuint64_t)(2 * copies)};
  cuuint64_t strides[2] = {(cuuint64_t)ld16 * 2, (cuuint64_t)zstride * 2};
  cuuint32_t box[3] = {32, (cuuint32_t)(kcontig ? box_mn : 32), 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

// Tensor map of C for the TMA-store epilogue: 32 x 32 boxes of the column-major output, no swizzle.
bool make_c_map(CUtensorMap* out, int es, CUtensorMapDataType dt, void* ptr, int64_t m, int64_t n, int64_t ld,
                int64_t batch, int64_t stride) {
  auto fn = get_encode_fn();
  if (!fn) return false;
  const bool batched = batch > 1;
  cuuint64_t dims[3] = {(cuuint64_t)m, (cuuint64_t)n, (cuuint64_t)(batched ? batch : 1)};
  cuuint64_t strides[2] = {(cuuint64_t)ld * es, (cuuint64_t)(batched ? stride : ld * n) * es};
  cuuint32_t box[3] = {32, 32, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(out, dt, 3, ptr, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

struct TcMaps {
  CUtensorMap a, b, c, alo, blo;
};

template <typename TIn, typename TOut, int BN, int STAGES, bool A_MN, bool B_MN, int CG, bool TRANS_OUT, int PRE>
int launch_inst(pbx_handle_t h, const TcMaps& tm, const TcParams& p) {
  using Cfg = TcCfg<sizeof(TIn), BN, STAGES, CG, sizeof(TOut), PRE>;
  static_assert(Cfg::SMEM_BYTES <= 227 * 1024, "shared memory budget");
  auto kern = gemm_tc_kernel<TIn, TOut, BN, STAGES, A_MN, B_MN, CG, TRANS_OUT, PRE>;
  static bool attr_set[16] = {};   // per device ordinal: the attribute is sticky, set it once
  if (h->device >= 16 || !attr_set[h->device]) {
    PBX_CUDA_CHECK(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    if (h->device < 16) attr_set[h->device] = true;
  }
  const int64_t units = h->sm_count / CG;   // persistent: one CTA (or CTA pair) per SM (pair)
  const int64_t groups = p.total_tiles < units ? p.total_tiles : units;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(groups * CG));
  cfg.blockDim = dim3(Cfg::NUM_THREADS);
  cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
  cfg.stream = h->stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  PBX_CUDA_CHECK(h, cudaLaunchKernelEx(&cfg, kern, tm.a, tm.b, tm.c, tm.alo, tm.blo, p));
  h->launches++;
  return PBX_OK;
}

template <typename TIn, typename TOut, int BN, int STAGES, int CG, bool TRANS_OUT, int PRE>
int launch_major(pbx_handle_t h, bool a_mn, bool b_mn, const TcMaps& tm, const TcParams& p) {
  if (a_mn) {
    return b_mn ? launch_inst<TIn, TOut, BN, STAGES, true, true, CG, TRANS_OUT, PRE>(h, tm, p)
                : launch_inst<TIn, TOut, BN, STAGES, true, false, CG, TRANS_OUT, PRE>(h, tm, p);
  }
  return b_mn ? launch_inst<TIn, TOut, BN, STAGES, false, true, CG, TRANS_OUT, PRE>(h, tm, p)
              : launch_inst<TIn, TOut, BN, STAGES, false, false, CG, TRANS_OUT, PRE>(h, tm, p);
}

// tile configurations, most efficient first: CTA pair 256x256, CTA pair 256x128, single CTA 128x128;
// bn == 64 is the skinny-M configuration (operands swapped, 128 columns of C x 64 rows per tile)
template <typename TIn, typename TOut, int PRE>
int launch_cfg_pre(pbx_handle_t h, int cg, int bn, bool a_mn, bool b_mn, const TcMaps& tm, const TcParams& p) {
  // fp32 stages hold raw + lo tiles (half as many stages as 16-bit) unless there are no lo tiles at all (PRE == 2)
  constexpr bool F32 = sizeof(TIn) == 4 && PRE != 2;
  if (bn == 64) return launch_major<TIn, TOut, 64, F32 ? 4 : 8, 1, true, PRE>(h, a_mn, b_mn, tm, p);  // swapped operands
  if (cg == 2 && bn == 256) return launch_major<TIn, TOut, 256, F32 ? 3 : 6, 2, false, PRE>(h, a_mn, b_mn, tm, p);
  if (cg == 2) return launch_major<TIn, TOut, 128, F32 ? 4 : 8, 2, false, PRE>(h, a_mn, b_mn, tm, p);
  return launch_major<TIn, TOut, 128, F32 ? 3 : 6, 1, false, PRE>(h, a_mn, b_mn, tm, p);
}

template <typename TIn, typename TOut>
int launch_cfg(pbx_handle_t h, int cg, int bn, bool a_mn, bool b_mn, int pre, const TcMaps& tm, const TcParams& p) {
  if constexpr (sizeof(TIn) == 4) {
    if (pre == 1) return launch_cfg_pre<TIn, TOut, 1>(h, cg, bn, a_mn, b_mn, tm, p);
    if (pre == 2) return launch_cfg_pre<TIn, TOut, 2>(h, cg, bn, a_mn, b_mn, tm, p);
    if (pre == 3) return launch_cfg_pre<TIn, TOut, 3>(h, cg, bn, a_mn, b_mn, tm, p);
  }
  return launch_cfg_pre<TIn, TOut, 0>(h, cg, bn, a_mn, b_mn, tm, p);
}

struct TcPlan {
  int cg, bn, slices;
  bool swap;   // compute C^T = op(B)^T op(A)^T (skinny M)
};

// Shape -> {cta_group, tile width, K slices}.  Replaces the reference's NVIDIA heuristics
// (src/interface/blas3/backend/nvidia_gpu.hpp:116-171: tile by M,N thresholds) with the quantity
// that matters on a 148-SM persistent kernel: how many tiles there are per SM (pair).
TcPlan make_plan(pbx_handle_t h, const PbxGemmCall& c) {
  const int64_t k_block = ROW_BYTES / (int64_t)pbx_in_size(c.dtype);
  const int64_t kb = (c.k + k_block - 1) / k_block;
  struct Cand { int cg, bn; };
  const Cand cands[3] = {{2, 256}, {2, 128}, {1, 128}};
  auto tiles_of = [&](const Cand& cd) {
    return ((c.m + 128 * cd.cg - 1) / (128 * cd.cg)) * ((c.n + cd.bn - 1) / cd.bn) * c.batch;
  };
  auto usable = [&](const Cand& cd) { return !(cd.cg == 2 && c.m <= 128) && !(cd.bn == 256 && c.n <= 128); };
  TcPlan plan = {1, 128, 1, false};
  const char* swap_env = getenv("PBX_TC_SWAP");   // "0" disables the skinny-M operand swap (testing)
  const bool want_swap = c.m <= 64 && c.n > c.m && !(swap_env && atoi(swap_env) == 0) && c.n_extra == 0;
  const char* force = getenv("PBX_TC_CONFIG");  // "cg,bn" (testing)
  int fcg = 0, fbn = 0;
  if (force && sscanf(force, "%d,%d", &fcg, &fbn) == 2 && (fcg == 1 || fcg == 2) && (fbn == 128 || (fbn == 256 && fcg == 2))) {
    plan.cg = fcg; plan.bn = fbn;
  } else if (want_swap) {
    plan.cg = 1; plan.bn = 64; plan.swap = true;
  } else {
    bool found = false;
    for (const Cand& cd : cands) {
      if (!usable(cd)) continue;
      const int64_t units = h->sm_count / cd.cg;
      if (tiles_of(cd) * 10 >= units * 6) { plan.cg = cd.cg; plan.bn = cd.bn; found = true; break; }
    }
    if (!found) {
      // too few tiles for any config: take the biggest usable tile if K is deep enough to be split
      // across the machine, else the smallest tile (most CTAs)
      const bool deep = c.k >= 2 * 2048;
      for (const Cand& cd : cands) {
        if (!usable(cd)) continue;
        plan.cg = cd.cg; plan.bn = cd.bn;
        if (deep) break;
      }
    }
  }
  // HBM-bound 16-bit shapes (BASELINE cfg4: 4096 x 256^3, 85 flop/B): measured on B200, 148 independent 128x128
  // tiles keep DRAM busier than 74 CTA pairs on 256x256 tiles (529 vs 518 TFLOP/s) -- the tensor pipe is a third
  // loaded either way, what counts is how many independent load streams are in flight.
  if (!(force && fcg) && !plan.swap && pbx_in_size(c.dtype) == 2) {
    const double flops = 2.0 * (double)c.m * (double)c.n * (double)c.k;
    const double byts = 2.0 * ((double)c.m * c.k + (double)c.k * c.n) + (double)pbx_out_size(c.dtype) * c.m * c.n;
    const Cand small = {1, 128};
    if (flops / byts < 100.0 && tiles_of(small) >= 2 * (int64_t)h->sm_count) { plan.cg = 1; plan.bn = 128; }
  }
  // K slices.  The reference splits by depth = ceil(4*CUs / tiles) when K > 2048
  // (gemm_partial_local.hpp:191-199, portblas_handle.hpp:323).  Here: when the output tiles cannot fill
  // half of the SM (pairs), spread the K loop over up to two waves of them, but keep at least 4 K blocks
  // (128 fp32 / 256 16-bit k) per slice and split only loops of >= 16 blocks: a lone CTA walks one
  // K block per ~1 us (TMA -> MMA latency chain), the reduce epilogue costs one extra short launch.
  const Cand chosen = {plan.cg, plan.bn};
  const int64_t tiles = plan.swap ? ((c.n + 127) / 128) * c.batch : tiles_of(chosen), units = h->sm_count / plan.cg;
  int64_t slices = 1;
  if (c.n_extra > 0) slices = 1;   // multicast epilogue: the tile is stored straight from TMEM, no split-K partials
  else if (h->forced_split_k > 1) slices = h->forced_split_k;
  else if (h->forced_split_k == 0 && tiles * 2 <= units && kb >= 16) {
    slices = (2 * units) / tiles;
    if (slices > kb / 4) slices = kb / 4;
  }
  if (slices > kb) slices = kb;
  if (slices < 1) slices = 1;
  const int64_t kbps = (kb + slices - 1) / slices;
  plan.slices = (int)((kb + kbps - 1) / kbps);  // drop empty trailing slices
  return plan;
}

}  // namespace

bool pbx_tma_operand_ok(int dtype, const void* p, int64_t ld, int64_t stride) {
  const int64_t es = (int64_t)pbx_in_size(dtype);
  return ((uintptr_t)p % 16 == 0) && ((ld * es) % 16 == 0) && ((stride * es) % 16 == 0) &&
         (ld * es < ((int64_t)1 << 40)) && (stride * es < ((int64_t)1 << 40));
}

bool pbx_tcgen05_shape_ok(pbx_handle_t h, const PbxGemmCall& c) {
  if (c.dtype == PBX_F64) return false;
  if (c.m >= ((int64_t)1 << 31) || c.n >= ((int64_t)1 << 31) || c.k >= ((int64_t)1 << 31) ||
      c.batch >= ((int64_t)1 << 31))
    return false;
  return get_encode_fn() != nullptr;
}

bool pbx_tcgen05_eligible(pbx_handle_t h, const PbxGemmCall& c) {
  return pbx_tcgen05_shape_ok(h, c) && pbx_tma_operand_ok(c.dtype, c.A, c.lda, c.sa) &&
         pbx_tma_operand_ok(c.dtype, c.B, c.ldb, c.sb);
}

int pbx_tcgen05_slices(pbx_handle_t h, const PbxGemmCall& c) { return make_plan(h, c).slices; }

// The tile plan as a pure function of the shape (no device needed): lets the CPU test-suite pin the selector.
extern "C" int pbx_plan_query(int sm_count, int dtype, int64_t m, int64_t n, int64_t k, int64_t batch, int* cta_group,
                              int* tile_n, int* k_slices, int* swapped) {
  if (sm_count <= 0 || dtype < PBX_F32 || dtype > PBX_BF16_F32 || dtype == PBX_F64 || m <= 0 || n <= 0 || k <= 0 ||
      batch <= 0 || !cta_group || !tile_n || !k_slices || !swapped)
    return PBX_ERR_INVALID_ARG;
  pbx_handle_s fake;
  fake.sm_count = sm_count;
  PbxGemmCall c;
  c.dtype = dtype; c.ta = c.tb = false;
  c.m = m; c.n = n; c.k = k; c.alpha = 1.0; c.beta = 0.0;
  c.A = c.B = nullptr; c.C = nullptr;
  c.lda = m; c.ldb = k; c.ldc = m; c.sa = m * k; c.sb = k * n; c.sc = m * n; c.batch = batch;
  const TcPlan p = make_plan(&fake, c);
  *cta_group = p.cg; *tile_n = p.bn; *k_slices = p.slices; *swapped = p.swap ? 1 : 0;
  return PBX_OK;
}

int pbx_launch_tcgen05(pbx_handle_t h, const PbxGemmCall& c, int slices) {
  const int es = (int)pbx_in_size(c.dtype);
  const bool f32 = (c.dtype == PBX_F32);
  const int bk = ROW_BYTES / es;
  TcPlan plan = make_plan(h, c);
  plan.slices = slices;
  const int bn = plan.bn, cg = plan.cg;
  CUtensorMapDataType dt = f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32
                               : ((c.dtype == PBX_F16 || c.dtype == PBX_F16_F32) ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16
                                                                                : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16);
  // The kernel's operands.  Normal: X = op(A) (M x K), Y = op(B) (K x N).  Swapped (skinny M): X = op(B)^T
  // (N x K), Y = op(A)^T (K x M), output transposed.  "mn-major" = the non-K index is the contiguous one.
  struct Opnd { const void* p; int64_t mn, ld, st; bool mn_major; };
  Opnd X = {c.A, c.m, c.lda, c.sa, !c.ta}, Y = {c.B, c.n, c.ldb, c.sb, c.tb};
  if (plan.swap) { Opnd t = X; X = Y; Y = t; }
  const bool a_mn = X.mn_major, b_mn = Y.mn_major;
  // fp32, compute-bound shapes: split the lo halves of both operands ONCE into pooled global buffers (one
  // HBM-bound pass each) instead of once per tile in shared memory.  The in-kernel splitters read and rewrite
  // every staged tile, which together with the three tf32 MMAs' operand reads oversubscribes the 128 B/clk of
  // shared memory (measured: tensor pipe 68 % active at 16384^3); with the lo tiles arriving by TMA the
  // mainloop's shared-memory traffic drops by a quarter and the CTA needs no splitter warps.  Memory-bound
  // shapes (arithmetic intensity < 256 flop/B) keep the in-kernel split: the pre-pass would triple their traffic.
  bool pre = false, split16 = false;
  int64_t s16_ld[2] = {0, 0}, s16_z[2] = {0, 0}, s16_copies[2] = {1, 1};
  const void* lo_ptr[2] = {nullptr, nullptr};
  // SB_ENABLE_JOINT_MATRIX=1 is the reference's switch (read per call, nvidia_gpu.hpp:68-69) from the fp32 kernels to
  // its tensor-core kernels with reduced-precision fragments; here it selects the single-tf32 product.
  const char* jm_env = getenv("SB_ENABLE_JOINT_MATRIX");
  const bool tf32x1 = f32 && jm_env != nullptr && jm_env[0] == '1';
  if (f32 && !tf32x1) {
    const int pre_env = getenv("PBX_TF32_PRESPLIT") ? atoi(getenv("PBX_TF32_PRESPLIT")) : -1;
    const double flops = 2.0 * (double)c.m * (double)c.n * (double)c.k * (double)c.batch;
    const double byts = 4.0 * ((double)c.m * c.k + (double)c.k * c.n + (double)c.m * c.n) * (double)c.batch;
    pre = (pre_env >= 0) ? (pre_env != 0) : (flops >= 5e8 && flops / byts >= 256.0);
    // EXPERIMENTAL, off unless PBX_F32_SPLIT16=1 (written without a GPU at hand; see TcCfg's PRE == 3): tf32 + 2 x bf16.
    // Takes the place of the fp32 lo pre-split on the shapes that would get it; same pooled buffers (2 x 2 bytes per
    // element instead of 4), the copies laid out [hi of every batch entry | lo of every batch entry].
    const char* s16_env = getenv("PBX_F32_SPLIT16");
    if (pre && s16_env != nullptr && s16_env[0] == '1' && c.n_extra == 0) {
      split16 = true;
      const Opnd* ops[2] = {&X, &Y};
      for (int i = 0; i < 2 && split16; ++i) {
        const Opnd& o = *ops[i];
        const int64_t rows = o.mn_major ? o.mn : c.k, cols = o.mn_major ? c.k : o.mn;
        const int64_t copies = (c.batch > 1 && o.st > 0) ? c.batch : 1;
        const int64_t ld16 = (rows + 7) / 8 * 8, st16 = ld16 * cols;   // multiples of 8 elements = 16 bytes
        if (2 * copies >= ((int64_t)1 << 31) || pbx_ensure_lo(h, i, 2 * copies * st16 * 2) != PBX_OK ||
            pbx_launch_split16(h, (const float*)o.p, h->lo[i], (char*)h->lo[i] + copies * st16 * 2, rows, cols, o.ld, o.st,
                               ld16, st16, copies) != PBX_OK)
          split16 = false;   // no room: the fp32 lo pre-split below takes over
        s16_ld[i] = ld16; s16_z[i] = st16; s16_copies[i] = copies;
      }
    }
    if (pre && !split16) {
      const Opnd* ops[2] = {&X, &Y};
      for (int i = 0; i < 2 && pre; ++i) {
        const Opnd& o = *ops[i];
        const int64_t rows = o.mn_major ? o.mn : c.k, cols = o.mn_major ? c.k : o.mn;
        const int64_t copies = (c.batch > 1 && o.st > 0) ? c.batch : 1;
        const int64_t elems = (copies - 1) * o.st + o.ld * cols;
        if (pbx_ensure_lo(h, i, elems * 4) != PBX_OK ||
            pbx_launch_tf32_lo(h, (const float*)o.p, (float*)h->lo[i], rows, cols, o.ld, o.st, copies) != PBX_OK) {
          pre = false;   // no room for the copies: fall back to the in-kernel split
          break;
        }
        lo_ptr[i] = h->lo[i];
      }
    }
  }
  TcMaps tm;
  if (!make_operand_map(&tm.a, es, dt, X.p, X.mn, c.k, X.ld, c.batch, X.st, !a_mn, BM) ||
      !make_operand_map(&tm.b, es, dt, Y.p, Y.mn, c.k, Y.ld, c.batch, Y.st, !b_mn, bn / cg)) {
    h->last_error = "cuTensorMapEncodeTiled failed";
    return PBX_ERR_CUDA;
  }
  tm.alo = tm.a; tm.blo = tm.b;   // placeholders when unused (never dereferenced)
  if (split16) {
    if (!make_split16_map(&tm.alo, h->lo[0], X.mn, c.k, s16_ld[0], s16_copies[0], s16_z[0], !a_mn, BM) ||
        !make_split16_map(&tm.blo, h->lo[1], Y.mn, c.k, s16_ld[1], s16_copies[1], s16_z[1], !b_mn, bn / cg)) {
      h->last_error = "cuTensorMapEncodeTiled failed (bf16 split copies)";
      return PBX_ERR_CUDA;
    }
  } else if (pre && (!make_operand_map(&tm.alo, es, dt, lo_ptr[0], X.mn, c.k, X.ld, c.batch, X.st, !a_mn, BM) ||
                     !make_operand_map(&tm.blo, es, dt, lo_ptr[1], Y.mn, c.k, Y.ld, c.batch, Y.st, !b_mn, bn / cg))) {
    h->last_error = "cuTensorMapEncodeTiled failed";
    return PBX_ERR_CUDA;
  }
  const int pre_mode = tf32x1 ? 2 : (split16 ? 3 : (pre ? 1 : 0));
  h->last_presplit = pre_mode;
  TcParams p;
  p.C = c.C; p.ws = (float*)h->ws;
  p.M = X.mn; p.N = Y.mn; p.K = c.k; p.ldc = c.ldc; p.sc = c.sc;
  p.alpha = (float)c.alpha; p.beta = (float)c.beta;
  p.batch = (int)c.batch; p.slices = slices;
  p.m_tiles = (int)((p.M + BM * cg - 1) / (BM * cg));
  p.n_tiles = (int)((p.N + bn - 1) / bn);
  p.group_m = cg == 2 ? 8 : 16;
  p.kb_total = (int)((c.k + bk - 1) / bk);
  p.kb_per_slice = (p.kb_total + slices - 1) / slices;
  // fp32: the tensor core truncates when it adds into its fp32 accumulator, so an accumulation
  // chain is limited to kb_per_chunk blocks of 32 (default 16 -> 512 k: ~4e-6 relative bias)
  const int chunk_env = getenv("PBX_TF32_CHUNK_KB") ? atoi(getenv("PBX_TF32_CHUNK_KB")) : 0;
  p.kb_per_chunk = f32 ? (chunk_env > 0 ? chunk_env : 16) : (1 << 30);
  // default: raw fp32 tile as the hi operand (verified on B200: kind::tf32 ignores the low 13 mantissa bits)
  const int raw_hi_env = getenv("PBX_TF32_RAW_HI") ? atoi(getenv("PBX_TF32_RAW_HI")) : 1;
  p.raw_hi = raw_hi_env;
  p.a_batched = (c.batch > 1 && X.st > 0) ? 1 : 0;
  p.b_batched = (c.batch > 1 && Y.st > 0) ? 1 : 0;
  const int64_t eo = (int64_t)pbx_out_size(c.dtype);
  p.c_vec = (((uintptr_t)c.C % 16 == 0) && (c.ldc * eo) % 16 == 0 && (c.sc * eo) % 16 == 0) ? 1 : 0;
  p.total_tiles = (int64_t)p.m_tiles * p.n_tiles * c.batch * slices;
  p.n_extra = c.n_extra;
  for (int x = 0; x < 7; ++x) p.Cx[x] = x < c.n_extra ? c.c_extra[x] : nullptr;

  // 16-bit C with beta == 0 leaves through shared memory + TMA stores when C is TMA-legal
  tm.c = tm.a;  // placeholder when unused (never dereferenced)
  p.tma_store = 0;
  const bool out16 = (c.dtype == PBX_F16 || c.dtype == PBX_BF16);
  const char* ts_env = getenv("PBX_TMA_STORE");
  if (out16 && c.beta == 0.0 && slices == 1 && c.n_extra == 0 && !(ts_env && atoi(ts_env) == 0) && ((uintptr_t)c.C % 16 == 0) &&
      (c.ldc * 2) % 16 == 0 && (c.batch == 1 || (c.sc * 2) % 16 == 0) && c.ldc * 2 < ((int64_t)1 << 40) &&
      c.sc * 2 < ((int64_t)1 << 40)) {
    if (make_c_map(&tm.c, 2, dt, c.C, c.m, c.n, c.ldc, c.batch, c.sc)) p.tma_store = 1;
  }

  switch (c.dtype) {
    case PBX_F32: return launch_cfg<float, float>(h, cg, bn, a_mn, b_mn, pre_mode, tm, p);
    case PBX_F16: return launch_cfg<__half, __half>(h, cg, bn, a_mn, b_mn, tf32x1 ? 2 : (pre ? 1 : 0), tm, p);
    case PBX_F16_F32: return launch_cfg<__half, float>(h, cg, bn, a_mn, b_mn, tf32x1 ? 2 : (pre ? 1 : 0), tm, p);
    case PBX_BF16: return launch_cfg<__nv_bfloat16, __nv_bfloat16>(h, cg, bn, a_mn, b_mn, tf32x1 ? 2 : (pre ? 1 : 0), tm, p);
    case PBX_BF16_F32: return launch_cfg<__nv_bfloat16, float>(h, cg, bn, a_mn, b_mn, tf32x1 ? 2 : (pre ? 1 : 0), tm, p);
  }
  return PBX_ERR_INVALID_ARG;
}
