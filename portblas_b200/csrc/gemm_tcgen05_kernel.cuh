#pragma once
// gemm_tcgen05_kernel.cuh -- the tcgen05 GEMM kernel template and its launcher (included by gemm_tcgen05.cu, which holds
// the tile plan and the tensor maps, and by the gemm_tc_inst_*.cu translation units, each of which instantiates one
// (element types, fp32 split mode) family: 16 kernels per unit, compiled in parallel).
//
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

#include <atomic>
#include <type_traits>
#include <mutex>

#include <cudaTypedefs.h>

#include "pbx_internal.cuh"
#include "tc_ptx.cuh"

// shared with gemm_tcgen05.cu and every instantiation unit (external linkage: they appear in the units' entry points)
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
  // dynamic tile schedule: sched[0] = tiles handed out beyond each group's first one, sched[1] = groups that are done
  // (the last one re-arms both for the next launch); dynamic == 0 -> static round-robin over the groups
  unsigned int* sched;
  int dynamic;
  // multicast GEMM, asynchronous form: the peer copies of a finished tile are pushed by TMA (see the pusher warp) instead
  // of being stored by the epilogue warps; push_pace: spread a tile's peer stores over the next tile's duration
  int push;
  int push_pace;
  unsigned int wait_hint_ns;   // suspend-time hint of the long mbarrier waits (0 = the instruction's default)
};

// tensor maps of the multicast GEMM's asynchronous peer copies: the local C (read back, 128 x 32 boxes) and the peers'
// C (fp32 / fp64-sized outputs: 128 x 32 boxes written by the pusher warp; 16-bit outputs: 32 x 32 boxes written
// straight from the TMA-store epilogue's staging tiles)
struct TcPushMaps {
  CUtensorMap local;
  CUtensorMap peer[7];
};

struct TcMaps {
  CUtensorMap a, b, c, alo, blo;
  TcPushMaps push;
};

namespace {

using namespace tcx;


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
//   4  tf32 + 2 x bf16 with the bf16 tiles made IN the kernel: only the raw tiles arrive by TMA, splitter warps convert
//      every staged tile into the same four bf16 tiles (re-laid out from the 128-byte-swizzled fp32 rows to 64-byte-
//      swizzled bf16 rows).  For shapes whose arithmetic intensity does not pay for a pre-pass (tall-skinny split-K,
//      mid-size problems): two tf32-MMA times per k-step instead of PRE == 0's three, and a third less operand traffic
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
  static constexpr int BAR_BYTES = 512;
  // Warp budget after the four role warps: 8 epilogue warps (4 TMEM lane quarters x 2 column halves) and, in the
  // in-kernel split modes, 4 splitter warps.  (4 epilogue + 8 splitter warps were measured in round 2: no gain -- SGEMM
  // 512 x 512 x 2^20 302.8 vs 300.7 TFLOP/s, 384 x 5408 x 3456 173 vs 181 -- the split modes are bound by shared-memory
  // bandwidth, ~160 KB per K block and SM against 128 B/clk, not by the splitters' issue rate.)
  static constexpr bool SPLIT_MODE = (ES == 4) && (PRE == 0 || PRE == 4);
  static constexpr int EPI_WARPS = 8;
  // 16-bit outputs: every epilogue warp owns two 32x32 staging tiles (column-major, rows contiguous)
  // that it hands to TMA stores, so C leaves the SM as bulk writes instead of 2-byte stores
  static constexpr int EPI_TILE_BYTES = 32 * 32 * OS;
  static constexpr int EPI_BYTES = (OS == 2) ? EPI_WARPS * 2 * EPI_TILE_BYTES : 0;
  // 32-bit outputs: two 128 x 32 staging boxes of the pusher warp (multicast GEMM): a finished tile is read back from
  // the local C (L2) and sent to the peers by TMA while the next tile is on the tensor cores
  static constexpr int PUSH_BOX_BYTES = BM * 32 * OS;
  static constexpr int PUSH_BYTES = (OS == 4) ? 2 * PUSH_BOX_BYTES : 0;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EPI_BYTES + PUSH_BYTES + BAR_BYTES + 1024;  // + align slack
  // In-kernel split modes (PRE 0 / 4) run TWO rings over the same STAGES * STAGE_BYTES of shared memory: a deep one of
  // raw tiles (TMA -> splitter -> MMA: its depth has to cover the HBM latency) and a shallow one of derived tiles (the
  // lo halves / bf16 tiles: splitter -> MMA only, an on-chip hand-over).  With one ring of 3 coupled stages every
  // in-kernel-split shape ran at one K block per ~1.4 us whatever its size (3 stages in flight over a ~4 us
  // TMA + split + MMA chain: SGEMM 512 x 512 x 2^20 sat at 1.7 TB/s with the tensor pipe a third busy).
  static constexpr bool SPLIT_RINGS = TF32X3 && (PRE == 0 || PRE == 4);
  static constexpr int DER_STAGES = SPLIT_RINGS ? 2 : 0;
  static constexpr int RAW_STAGES = SPLIT_RINGS ? 2 * STAGES - 2 : STAGES;
  static constexpr int RAW_STRIDE = SPLIT_RINGS ? RAW_BYTES : STAGE_BYTES;   // bytes between consecutive raw stages
  static constexpr int DER_BASE = RAW_STAGES * RAW_STRIDE;                   // derived stage d at DER_BASE + d * RAW_BYTES
  static_assert(!SPLIT_RINGS || DER_BASE + DER_STAGES * RAW_BYTES == STAGES * STAGE_BYTES, "ring split");
  // TMEM: ACC_STAGES accumulators of BN columns (+ for fp32 a BN-column running sum)
  static constexpr int ACC_STAGES = TF32X3 ? ((3 * BN <= 512) ? 2 : 1) : 2;
  static constexpr int RSUM_COL = ACC_STAGES * BN;
  static constexpr int TMEM_COLS = TF32X3 ? 512 : 2 * BN;
  static constexpr int SPLIT_WARPS = SPLIT_MODE ? 4 : 0;
  static constexpr int TMA_BYTES = RAW_BYTES * ((TF32X3 && (PRE == 1 || PRE == 3)) ? 2 : 1);   // bytes one CTA's producer lands per stage
  static constexpr int NUM_THREADS = 32 * (4 + EPI_WARPS + SPLIT_WARPS);
  static constexpr int NUM_SPLIT_THREADS = 32 * SPLIT_WARPS;
  // tile scheduler ring: the leader's producer lane publishes the next tile index, every other role reads it
  static constexpr int SCHED_SLOTS = 4;
  // MMA lane (leader) / producer lane (peer) + pusher lane + epilogue and splitter warps
  static constexpr int SCHED_CONSUMERS = 2 + EPI_WARPS + SPLIT_WARPS;
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
               const __grid_constant__ CUtensorMap tmBlo, const __grid_constant__ TcPushMaps tmPush, const TcParams p) {
  using Cfg = TcCfg<sizeof(TIn), BN, STAGES, CG, sizeof(TOut), PRE>;
  static_assert(PRE == 0 || Cfg::TF32X3, "pre-split / single-tf32 modes exist for fp32 only");
  constexpr bool TF32X3 = Cfg::TF32X3;
  constexpr int BK = Cfg::BK;
  constexpr int ACC_STAGES = Cfg::ACC_STAGES;
  constexpr uint32_t FMT = TF32X3 ? 2u : (std::is_same<TIn, __nv_bfloat16>::value ? 1u : 0u);

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t epi_base = smem_base + STAGES * Cfg::STAGE_BYTES;
  const uint32_t push_base = epi_base + Cfg::EPI_BYTES;
  const uint32_t bar_base = push_base + Cfg::PUSH_BYTES;
  // barrier map (8 B each): full[S] | empty[S] | split[S] | tmem_full[2] | tmem_empty[2] | tmem_ptr
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  constexpr int RS = Cfg::RAW_STAGES;   // ring of TMA-written stages (== STAGES unless the split rings are in use)
  constexpr int DS = Cfg::DER_STAGES;   // ring of splitter-written stages
  auto empty_bar = [&](int s) { return bar_base + 8u * (RS + s); };
  auto split_bar = [&](int s) { return bar_base + 8u * (2 * RS + s); };   // [DS] with the split rings, else unused
  auto tfull_bar = [&](int s) { return bar_base + 8u * (3 * RS + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (3 * RS + 2 + s); };
  const uint32_t tmem_ptr_smem = bar_base + 8u * (3 * RS + 4);
  // scheduler ring: sfull[4] | sempty[4] (the LEADER's are the ones waited on) | tile index [4]
  const uint32_t sched_base = bar_base + 8u * (3 * RS + 5);
  auto sfull_bar = [&](int s) { return sched_base + 8u * s; };
  auto sempty_bar = [&](int s) { return sched_base + 8u * (Cfg::SCHED_SLOTS + s); };
  auto stile = [&](int s) { return sched_base + 16u * Cfg::SCHED_SLOTS + 4u * s; };
  // pusher: load barriers of its two staging boxes | count of epilogue warps that finished storing a tile
  const uint32_t pload_base = sched_base + 20u * Cfg::SCHED_SLOTS;
  auto pload_bar = [&](int b) { return pload_base + 8u * b; };
  const uint32_t push_count = pload_base + 16u;
  auto dempty_bar = [&](int s) { return pload_base + 24u + 8u * s; };   // derived stage s is free again (MMA -> splitters)
  static_assert(8 * (3 * RS + 5) + 20 * Cfg::SCHED_SLOTS + 24 + 8 * 2 <= Cfg::BAR_BYTES, "barrier area");
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
    for (int s = 0; s < RS; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int s = 0; s < DS; ++s) {
      mbar_init(split_bar(s), CG);   // one arrival per CTA: its splitter warps meet on a named barrier first
      mbar_init(dempty_bar(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull_bar(s), 1);
      mbar_init(tempty_bar(s), CG * Cfg::EPI_WARPS);
    }
    for (int s = 0; s < Cfg::SCHED_SLOTS; ++s) {
      mbar_init(sfull_bar(s), 1);
      mbar_init(sempty_bar(s), CG * Cfg::SCHED_CONSUMERS);
    }
    mbar_init(pload_bar(0), 1);
    mbar_init(pload_bar(1), 1);
    st_shared_u32(push_count, 0u);
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
  // Programmatic dependent launch: everything above (barriers, TMEM, descriptor prefetch) may overlap the tail of the
  // previous kernel in the stream; from here on global memory is touched.
  grid_dependency_wait();
  grid_launch_dependents();

  // ---- tile schedule ------------------------------------------------------------------------------------------
  // A group's first tile is its own index; every further one is claimed from a global counter (p.sched[0]) by the
  // leader's producer lane one tile ahead and handed to the other roles (and to the peer CTA) through a 4-slot ring in
  // shared memory.  HBM-bound shapes need this: with a static round-robin the SMs near the memory partitions finish up
  // to 20 % earlier than the slowest ones (ncu, batch 4096 of 256^3: SM active cycles 77 % .. 95 % of the kernel).
  const uint32_t sempty_leader0 = (CG == 2) ? map_to_cta(sempty_bar(0), 0) : sempty_bar(0);
  // n-th hand-over as seen by a consumer warp (all lanes call; returns the next tile or total_tiles at the end)
  auto sched_consume = [&](uint32_t n) -> int64_t {
    const int slot = (int)(n % Cfg::SCHED_SLOTS);
    const uint32_t ph = (n / Cfg::SCHED_SLOTS) & 1u;
    if (CG == 2) mbar_wait_cluster(sfull_bar(slot), ph); else mbar_wait(sfull_bar(slot), ph);
    const uint32_t v = ld_shared_u32(stile(slot));
    __syncwarp();
    if (lane == 0) {
      uint32_t dep;   // the arrive's address depends on the loaded value: the slot has been read before it is released
      asm volatile("and.b32 %0, %1, 0;" : "=r"(dep) : "r"(v));
      if (CG == 2) mbar_arrive_remote(sempty_leader0 + 8u * slot + dep); else mbar_arrive(sempty_bar(slot) + dep);
    }
    return (v == 0xFFFFFFFFu) ? p.total_tiles : (int64_t)v;
  };
  auto next_tile = [&](int64_t tile, uint32_t& n) -> int64_t {
    if (!p.dynamic) return tile + num_groups;
    return sched_consume(n++);
  };

  if (warp == 0) {
    // ===================== TMA producer (every CTA loads its own A rows and B rows) =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      // pair mode without splitters (16-bit, pre-split fp32): both CTAs credit the leader's full barrier (the MMA
      // issuer waits there).  fp32 with in-kernel split: each CTA's splitter warps wait on their OWN full barrier.
      constexpr bool kLeaderFull = (CG == 2) && (!TF32X3 || (PRE != 0 && PRE != 4));
      uint32_t handed = 0;   // hand-overs published (leader) / consumed (peer) so far
      const uint32_t sfull_peer0 = (CG == 2) ? map_to_cta(sfull_bar(0), 1) : 0u;
      const uint32_t stile_peer0 = (CG == 2) ? map_to_cta(stile(0), 1) : 0u;
      for (int64_t tile = group; tile < p.total_tiles;) {
        unsigned int claim = 0;
        if (p.dynamic && rank == 0) claim = atomicAdd(p.sched, 1u);   // next tile, needed only after this one's loads
        const TileCoord tc = decode_tile(p, tile);
        const int m0 = tc.mt * Cfg::TILE_M + (int)rank * BM;
        const int n0 = tc.nt * BN + (int)rank * Cfg::BN_CTA;
        const int kb0 = tc.slice * p.kb_per_slice;
        const int kb1 = min(p.kb_total, kb0 + p.kb_per_slice);
        const int za = p.a_batched ? tc.b : 0, zb = p.b_batched ? tc.b : 0;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1u, p.wait_hint_ns);
          const uint32_t sA = smem_base + stage * Cfg::RAW_STRIDE;
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
          if (++stage == RS) { stage = 0; phase ^= 1u; }
        }
        if (!p.dynamic) {
          tile += num_groups;
        } else if (rank == 0) {
          const int64_t nxt = num_groups + (int64_t)claim;
          const uint32_t v = nxt < p.total_tiles ? (uint32_t)nxt : 0xFFFFFFFFu;
          const int slot = (int)(handed % Cfg::SCHED_SLOTS);
          const uint32_t ph = (handed / Cfg::SCHED_SLOTS) & 1u;
          if (CG == 2) mbar_wait_cluster(sempty_bar(slot), ph ^ 1u); else mbar_wait(sempty_bar(slot), ph ^ 1u);
          st_shared_u32(stile(slot), v);
          mbar_arrive(sfull_bar(slot));
          if (CG == 2) {
            st_shared_cluster_u32(stile_peer0 + 4u * slot, v);
            mbar_arrive_cluster(sfull_peer0 + 8u * slot);   // release.cluster: the slot is written before the arrive lands
          }
          ++handed;
          tile = nxt < p.total_tiles ? nxt : p.total_tiles;
        } else {
          // the peer's producer lane is a consumer of the leader's schedule (single lane: no __syncwarp needed)
          const int slot = (int)(handed % Cfg::SCHED_SLOTS);
          const uint32_t ph = (handed / Cfg::SCHED_SLOTS) & 1u;
          mbar_wait_cluster(sfull_bar(slot), ph);
          const uint32_t v = ld_shared_u32(stile(slot));
          uint32_t dep;
          asm volatile("and.b32 %0, %1, 0;" : "=r"(dep) : "r"(v));
          mbar_arrive_remote(sempty_leader0 + 8u * slot + dep);
          ++handed;
          tile = (v == 0xFFFFFFFFu) ? p.total_tiles : (int64_t)v;
        }
      }
      // re-arm the counters for the next launch: the last group to finish resets them (every group made its final
      // claim before it got here)
      if (p.dynamic && rank == 0) {
        __threadfence();
        if (atomicAdd(p.sched + 1, 1u) == (unsigned int)(num_groups - 1)) {
          p.sched[0] = 0u;
          p.sched[1] = 0u;
          __threadfence();
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
      int stage = 0, dstage = 0;
      uint32_t phase = 0, dphase = 0;
      int it = 0;  // accumulator hand-offs so far (one per K chunk)
      uint32_t handed = 0;
      for (int64_t tile = group; tile < p.total_tiles;) {
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
            constexpr bool kSplit = Cfg::SPLIT_RINGS;
            // in-kernel split: the derived stage is complete once every splitter thread has arrived (they waited for
            // the raw stage's TMA first); in pair mode the peer's splitter warps wrote shared memory with ordinary stores
            if (kSplit) { if (CG == 2) mbar_wait_cluster(split_bar(dstage), dphase); else mbar_wait(split_bar(dstage), dphase); }
            else mbar_wait(full_bar(stage), phase, p.wait_hint_ns);
            tc_fence_after();
            const uint32_t sA = smem_base + stage * Cfg::RAW_STRIDE;
            const uint32_t sB = sA + Cfg::A_BYTES;
            // derived half of the k-block: [A lo | B lo] (3xTF32) or [A16 hi | A16 lo | B16 hi | B16 lo] (bf16 split)
            const uint32_t sX = kSplit ? smem_base + Cfg::DER_BASE + dstage * Cfg::RAW_BYTES : sA + Cfg::RAW_BYTES;
#pragma unroll
            for (int k = 0; k < BK / Cfg::UMMA_K; ++k) {
              const uint64_t adesc = make_smem_desc(sA + k * A_KSTEP, A_LBO, A_SBO, A_LT);
              const uint64_t bdesc = make_smem_desc(sB + k * B_KSTEP, B_LBO, B_SBO, B_LT);
              const uint32_t acc = (kb > kc0 || k > 0) ? 1u : 0u;
              if (PRE == 3 || PRE == 4) {
                mma(d_tmem, adesc, bdesc, acc);   // hi * hi (tf32 on the raw tiles); the cross terms follow per k-block
              } else if (TF32X3 && PRE != 2) {
                const uint64_t adesc_lo = make_smem_desc(sX + k * A_KSTEP, A_LBO, A_SBO, A_LT);
                const uint64_t bdesc_lo = make_smem_desc(sX + Cfg::A_BYTES + k * B_KSTEP, B_LBO, B_SBO, B_LT);
                mma(d_tmem, adesc_lo, bdesc, acc);
                mma(d_tmem, adesc, bdesc_lo, 1u);
                mma(d_tmem, adesc, bdesc, 1u);
              } else {
                mma(d_tmem, adesc, bdesc, acc);
              }
            }
            if (PRE == 3 || PRE == 4) {
              // cross terms on the bf16 tiles: 64-byte rows (K-major) / 64-byte mn chunks (MN-major), 64B swizzle
              // (layout type 4): SBO = 8 rows x 64 B; MN-major LBO = one 32 x 32 box; UMMA_K = 16
              constexpr uint32_t idesc16 = make_idesc(1u, A_MN, B_MN, BN, Cfg::TILE_M);
              constexpr uint32_t A16_KSTEP = A_MN ? 16 * 64 : 32, B16_KSTEP = B_MN ? 16 * 64 : 32;
              constexpr uint32_t A16_LBO = A_MN ? 2048 : 16, B16_LBO = B_MN ? 2048 : 16;
              const uint32_t sA16 = sX, sB16 = sA16 + Cfg::A_BYTES;
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
            if (kSplit) {
              commit(dempty_bar(dstage));
              if (++dstage == DS) { dstage = 0; dphase ^= 1u; }
            }
            if (++stage == RS) { stage = 0; phase ^= 1u; }
          }
          commit(tfull_bar(as));  // chunk accumulator complete (both CTAs' epilogues)
        }
        if (!p.dynamic) {
          tile += num_groups;
        } else {   // single lane: the consumer protocol without the warp-wide parts
          const int slot = (int)(handed % Cfg::SCHED_SLOTS);
          const uint32_t ph = (handed / Cfg::SCHED_SLOTS) & 1u;
          mbar_wait(sfull_bar(slot), ph);
          const uint32_t v = ld_shared_u32(stile(slot));
          uint32_t dep;
          asm volatile("and.b32 %0, %1, 0;" : "=r"(dep) : "r"(v));
          mbar_arrive(sempty_bar(slot) + dep);
          ++handed;
          tile = (v == 0xFFFFFFFFu) ? p.total_tiles : (int64_t)v;
        }
      }
    }
    __syncwarp();
  } else if (warp == 3) {
    // ===================== pusher (one lane): asynchronous peer copies of the multicast GEMM =====================
    // The epilogue warps store a finished tile into the LOCAL C only and count themselves in; this lane then reads the
    // tile back (L2 hits) in 128 x 32 boxes and sends every box to all peers with bulk tensor stores.  Nothing waits for
    // NVLink except this lane: the accumulator is released as soon as the local stores are issued, so the transfer of
    // tile i runs under the mainloop of tile i+1 (round 1 stored to the peers from the epilogue warps: each tile's
    // 8 x 128 KiB burst had to drain through NVLink before the next MMA could start: 5.28 ms against 4.32 ms at 8 GPUs).
    // Pacing: the per-SM TMA engine serves the producer's operand loads and these stores in order, so a burst of peer
    // stores that NVLink cannot absorb at once (148 SMs x 7 peers x 128 KiB at the end of every round of tiles) holds
    // up the loads behind it -- measured at 8 GPUs: +0.95 ms on 3.56 ms, i.e. the whole transfer time exposed although
    // no warp waited for it.  The lane therefore spreads a tile's stores over about three quarters of the time the
    // previous tile took (clock64 between tile completions); only a group's last tile goes out at full speed.
    if (lane == 0) {
      uint32_t handed = 0, ntile = 0, chunk = 0;
      long long t_prev = clock64();
      for (int64_t tile = group; tile < p.total_tiles;) {
        // the next tile first (it was published while this one was being loaded): is this the group's last one?
        int64_t next;
        if (!p.dynamic) {
          next = tile + num_groups;
        } else {   // single lane: the consumer protocol without the warp-wide parts
          const int slot = (int)(handed % Cfg::SCHED_SLOTS);
          const uint32_t ph = (handed / Cfg::SCHED_SLOTS) & 1u;
          if (CG == 2) mbar_wait_cluster(sfull_bar(slot), ph); else mbar_wait(sfull_bar(slot), ph);
          const uint32_t v = ld_shared_u32(stile(slot));
          uint32_t dep;
          asm volatile("and.b32 %0, %1, 0;" : "=r"(dep) : "r"(v));
          if (CG == 2) mbar_arrive_remote(sempty_leader0 + 8u * slot + dep); else mbar_arrive(sempty_bar(slot) + dep);
          ++handed;
          next = (v == 0xFFFFFFFFu) ? p.total_tiles : (int64_t)v;
        }
        if (Cfg::PUSH_BYTES > 0 && p.push) {
          const TileCoord tc = decode_tile(p, tile);
          const int m0 = tc.mt * Cfg::TILE_M + (int)rank * BM;
          const int n0 = tc.nt * BN;
          const uint32_t target = (uint32_t)Cfg::EPI_WARPS * (ntile + 1u);
          while (ld_acquire_shared_u32(push_count) < target) { }
          const long long t_done = clock64();
          const long long period = t_done - t_prev;   // first tile: since the kernel started
          t_prev = t_done;
          const bool last = next >= p.total_tiles;
          const long long gap = (last || p.push_pace == 0) ? 0 : (period * 3 / 4) / ((BN / 32) * max(p.n_extra, 1));
          for (int j = 0; j < BN / 32; ++j) {
            if (n0 + 32 * j >= p.N || m0 >= p.M) break;
            const uint32_t buf = chunk & 1u;
            bulk_wait_read<1>();   // the stores that last used this box have read it
            const uint32_t box = push_base + buf * Cfg::PUSH_BOX_BYTES;
            mbar_expect_tx(pload_bar(buf), Cfg::PUSH_BOX_BYTES);
            tma_load_3d(box, &tmPush.local, pload_bar(buf), m0, n0 + 32 * j, tc.b);
            mbar_wait(pload_bar(buf), (chunk >> 1) & 1u);
            for (int x = 0; x < p.n_extra; ++x) {
              const long long t0 = clock64();
              tma_store_3d(&tmPush.peer[x], box, m0, n0 + 32 * j, tc.b);
              if (x + 1 == p.n_extra) bulk_commit();
              while (clock64() - t0 < gap) __nanosleep(256);
            }
            ++chunk;
          }
          ++ntile;
        }
        tile = next;
      }
      if (Cfg::PUSH_BYTES > 0 && p.push) bulk_wait_all();   // every peer copy has left before the CTA retires
    }
    __syncwarp();
  } else if (warp >= 4 && warp < 4 + Cfg::EPI_WARPS) {
    // ===================== epilogue: 8 warps = 4 TMEM lane quarters x 2 column halves =====================
    const int ew = warp & 3;           // TMEM lane quarter this warp may read
    const int ch = (warp - 4) >> 2;    // column half
    constexpr int COLS_PER_WARP = BN / (Cfg::EPI_WARPS / 4);   // column groups = epilogue warps per lane quarter
    constexpr bool OUT16 = (sizeof(TOut) == 2);
    const bool beta0 = (p.beta == 0.0f);
    const bool tma_store = OUT16 && p.tma_store;
    const uint32_t tempty_leader0 = (CG == 2) ? map_to_cta(tempty_bar(0), 0) : tempty_bar(0);
    // staging tiles of this warp (16-bit outputs): [buffer][column][32 rows]
    const uint32_t stage_u32 = epi_base + (uint32_t)((warp - 4) * 2 * Cfg::EPI_TILE_BYTES);
    TOut* stage_ptr = reinterpret_cast<TOut*>(smem_raw + (stage_u32 - smem_u32(smem_raw)));
    uint32_t store_blk = 0;
    int it = 0;
    uint32_t handed = 0;
    for (int64_t tile = group; tile < p.total_tiles; tile = next_tile(tile, handed)) {
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
        mbar_wait(tfull_bar(as), aphase, p.wait_hint_ns);
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
              // multicast GEMM: the same staging tile goes to every peer's C (bulk stores over NVLink, asynchronous)
              for (int x = 0; x < p.n_extra; ++x)
                tma_store_3d(&tmPush.peer[x], stage_u32 + buf * Cfg::EPI_TILE_BYTES, (int)m_warp, (int)(n0 + c0), tc.b);
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
              for (int x = 0; x < (p.push ? 0 : p.n_extra); ++x) {
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
                  for (int x = 0; x < (p.push ? 0 : p.n_extra); ++x)
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
      if (Cfg::PUSH_BYTES > 0 && p.push) {
        // this warp's part of the tile is in the local C: order the stores before the pusher's TMA reads of them
        // (generic proxy -> async proxy), then count the warp in
        fence_proxy_async_global();
        __syncwarp();
        if (lane == 0) red_release_shared_add(push_count, 1u);
      }
    }
    if (OUT16 && tma_store && lane == 0) bulk_wait_read<0>();   // staging tiles must outlive their stores
    __syncwarp();
  } else if (TF32X3 && PRE == 4 && warp >= 4 + Cfg::EPI_WARPS) {
    // ===================== fp32 -> bf16 (hi, lo) tiles, made in shared memory (each CTA converts what it staged) ========
    // source: the raw fp32 tile, 128-byte rows in the 128B swizzle (K-major: row = mn index, 16-byte chunk c of the
    // row sits at c ^ (row % 8)) or, MN-major, 4 KiB boxes of 32 k-rows x 32 mn in the 32B-atom flavour (32-byte unit
    // u of row r sits at u ^ (r % 4)).  destination: bf16 tiles with 64-byte rows in the 64B swizzle (16-byte chunk c
    // of row r sits at c ^ ((r / 2) % 4)), the layout the PRE == 3 tensor maps deliver: [A16 hi | A16 lo | B16 hi | B16 lo].
    const int st = threadIdx.x - 32 * (4 + Cfg::EPI_WARPS);
    int stage = 0, dstage = 0;
    uint32_t phase = 0, dphase = 0;
    const uint32_t split_leader0 = (CG == 2) ? map_to_cta(split_bar(0), 0) : split_bar(0);
    auto lo_of = [](float x) {
      const float d = x - __uint_as_float(__float_as_uint(x) & 0xffffe000u);
      return d == d ? d : 0.0f;   // inf - inf -> 0
    };
    auto cvt8 = [&](const float4 x0, const float4 x1, uint8_t* hi_dst, uint8_t* lo_dst) {
      __nv_bfloat162 h[4], l[4];
      h[0] = __floats2bfloat162_rn(x0.x, x0.y); h[1] = __floats2bfloat162_rn(x0.z, x0.w);
      h[2] = __floats2bfloat162_rn(x1.x, x1.y); h[3] = __floats2bfloat162_rn(x1.z, x1.w);
      l[0] = __floats2bfloat162_rn(lo_of(x0.x), lo_of(x0.y)); l[1] = __floats2bfloat162_rn(lo_of(x0.z), lo_of(x0.w));
      l[2] = __floats2bfloat162_rn(lo_of(x1.x), lo_of(x1.y)); l[3] = __floats2bfloat162_rn(lo_of(x1.z), lo_of(x1.w));
      *reinterpret_cast<uint4*>(hi_dst) = *reinterpret_cast<const uint4*>(h);
      *reinterpret_cast<uint4*>(lo_dst) = *reinterpret_cast<const uint4*>(l);
    };
    // one operand tile of ROWS mn-indices x 32 k.  A thread owns ROWS*4/128 items of 8 values; all of its loads are
    // issued before the first conversion (a lone warp per scheduler has nothing else to hide the shared-memory latency)
    auto convert = [&](const uint8_t* raw, uint8_t* hi16, uint8_t* lo16, auto rows_tag, auto mn_tag) {
      constexpr int ROWS = decltype(rows_tag)::value;
      constexpr bool MN = decltype(mn_tag)::value;
      constexpr int NSPLIT = Cfg::NUM_SPLIT_THREADS > 0 ? Cfg::NUM_SPLIT_THREADS : 128;   // (instantiated for every PRE)
      constexpr int IT = (ROWS * 4 + NSPLIT - 1) / NSPLIT;
      float4 x0[IT], x1[IT];
      int off[IT];
#pragma unroll
      for (int t = 0; t < IT; ++t) {
        const int i = st + t * NSPLIT;
        const int c = i & 3;
        if (ROWS * 4 % NSPLIT != 0 && i >= ROWS * 4) { off[t] = -1; continue; }
        if (!MN) {
          const int r = i >> 2;   // mn index; k = 8c .. 8c+7 = fp32 chunks 2c, 2c+1
          const uint8_t* src = raw + r * 128;
          x0[t] = *reinterpret_cast<const float4*>(src + (((2 * c) ^ (r & 7)) << 4));
          x1[t] = *reinterpret_cast<const float4*>(src + (((2 * c + 1) ^ (r & 7)) << 4));
          off[t] = r * 64 + ((c ^ ((r >> 1) & 3)) << 4);
        } else {
          const int r = (i >> 2) & 31, box = i >> 7;   // k row of the box; mn = 32*box + 8c .. 8c+7 = 32-byte unit c
          const uint8_t* src = raw + box * 4096 + r * 128 + ((c ^ (r & 3)) << 5);
          // odd rows read their second half first: the two rows of a quarter-warp then never meet on a bank
          const int h0 = (r & 1) << 4;
          const float4 xa = *reinterpret_cast<const float4*>(src + h0);
          const float4 xb = *reinterpret_cast<const float4*>(src + (h0 ^ 16));
          x0[t] = (r & 1) ? xb : xa;
          x1[t] = (r & 1) ? xa : xb;
          off[t] = box * 2048 + r * 64 + ((c ^ ((r >> 1) & 3)) << 4);
        }
      }
#pragma unroll
      for (int t = 0; t < IT; ++t)
        if (ROWS * 4 % NSPLIT == 0 || off[t] >= 0) cvt8(x0[t], x1[t], hi16 + off[t], lo16 + off[t]);
    };
    uint32_t handed = 0;
    for (int64_t tile = group; tile < p.total_tiles; tile = next_tile(tile, handed)) {
      const TileCoord tc = decode_tile(p, tile);
      const int kb0 = tc.slice * p.kb_per_slice;
      const int kb1 = min(p.kb_total, kb0 + p.kb_per_slice);
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(full_bar(stage), phase);            // raw tiles have landed
        mbar_wait(dempty_bar(dstage), dphase ^ 1u);   // the MMAs that read this derived stage last have retired
        uint8_t* sm0 = smem_raw + (smem_base - smem_u32(smem_raw));
        uint8_t* sA = sm0 + stage * Cfg::RAW_STRIDE;
        uint8_t* sB = sA + Cfg::A_BYTES;
        uint8_t* s16 = sm0 + Cfg::DER_BASE + dstage * Cfg::RAW_BYTES;
        convert(sA, s16, s16 + Cfg::A_BYTES / 2, std::integral_constant<int, BM>{}, std::integral_constant<bool, A_MN>{});
        convert(sB, s16 + Cfg::A_BYTES, s16 + Cfg::A_BYTES + Cfg::B_BYTES / 2, std::integral_constant<int, Cfg::BN_CTA>{},
                std::integral_constant<bool, B_MN>{});
        // every thread makes its tile writes visible to the async proxy (the tensor core reads them through it), the
        // splitter warps meet on a named barrier and ONE thread signals the MMA issuer.  In pair mode that signal is a
        // remote arrive; it carries no fence of its own -- a cluster-scope release compiles to MEMBAR.ALL.GPU, and 128
        // of those per stage and CTA (round 1) held the pair tiles of the in-kernel split modes at 1.6 us per K block
        // against 0.54 us of MMA time (ncu: the MMA lane waiting on this barrier, the splitters on the MMA's commit)
        fence_proxy_async();
        named_bar_sync(1, Cfg::NUM_SPLIT_THREADS);
        if (st == 0) { if (CG == 2) mbar_arrive_relaxed_cluster(split_leader0 + 8u * dstage); else mbar_arrive(split_bar(dstage)); }
        if (++dstage == DS) { dstage = 0; dphase ^= 1u; }
        if (++stage == RS) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (TF32X3 && PRE == 0 && warp >= 4 + Cfg::EPI_WARPS) {
    // ===================== fp32 -> (hi, lo) tf32 splitters (each CTA splits what it staged) ==============
    const int st = threadIdx.x - 32 * (4 + Cfg::EPI_WARPS);
    int stage = 0, dstage = 0;
    uint32_t phase = 0, dphase = 0;
    constexpr int VEC_PER_STAGE = Cfg::RAW_BYTES / 16;
    const uint32_t split_leader0 = (CG == 2) ? map_to_cta(split_bar(0), 0) : split_bar(0);
    const bool raw_hi = p.raw_hi != 0;
    uint32_t handed = 0;
    for (int64_t tile = group; tile < p.total_tiles; tile = next_tile(tile, handed)) {
      const TileCoord tc = decode_tile(p, tile);
      const int kb0 = tc.slice * p.kb_per_slice;
      const int kb1 = min(p.kb_total, kb0 + p.kb_per_slice);
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(full_bar(stage), phase);            // raw tiles have landed
        mbar_wait(dempty_bar(dstage), dphase ^ 1u);   // the MMAs that read this derived stage last have retired
        // elementwise, so the swizzled placement is preserved: the lo tiles mirror the raw tiles' layout
        uint8_t* sm0 = smem_raw + (smem_base - smem_u32(smem_raw));
        float4* raw = reinterpret_cast<float4*>(sm0 + stage * Cfg::RAW_STRIDE);
        float4* lo = reinterpret_cast<float4*>(sm0 + Cfg::DER_BASE + dstage * Cfg::RAW_BYTES);
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
        // every thread makes its tile writes visible to the async proxy (the tensor core reads them through it), the
        // splitter warps meet on a named barrier and ONE thread signals the MMA issuer.  In pair mode that signal is a
        // remote arrive; it carries no fence of its own -- a cluster-scope release compiles to MEMBAR.ALL.GPU, and 128
        // of those per stage and CTA (round 1) held the pair tiles of the in-kernel split modes at 1.6 us per K block
        // against 0.54 us of MMA time (ncu: the MMA lane waiting on this barrier, the splitters on the MMA's commit)
        fence_proxy_async();
        named_bar_sync(1, Cfg::NUM_SPLIT_THREADS);
        if (st == 0) { if (CG == 2) mbar_arrive_relaxed_cluster(split_leader0 + 8u * dstage); else mbar_arrive(split_bar(dstage)); }
        if (++dstage == DS) { dstage = 0; dphase ^= 1u; }
        if (++stage == RS) { stage = 0; phase ^= 1u; }
      }
    }
  }

  tc_fence_before();
  if (CG == 2) cluster_sync_all(); else __syncthreads();
  if (warp == 2) {
    if (CG == 2) tmem_dealloc_2sm(tmem_base, Cfg::TMEM_COLS); else tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}


template <typename TIn, typename TOut, int BN, int STAGES, bool A_MN, bool B_MN, int CG, bool TRANS_OUT, int PRE>
int launch_inst(pbx_handle_t h, const TcMaps& tm, const TcParams& p_in) {
  using Cfg = TcCfg<sizeof(TIn), BN, STAGES, CG, sizeof(TOut), PRE>;
  static_assert(Cfg::SMEM_BYTES <= 227 * 1024, "shared memory budget");
  auto kern = gemm_tc_kernel<TIn, TOut, BN, STAGES, A_MN, B_MN, CG, TRANS_OUT, PRE>;
  static std::atomic<uint32_t> attr_set{0};   // bit per device ordinal: the attribute is sticky, set it once
  if (h->device >= 32 || !(attr_set.load(std::memory_order_acquire) & (1u << h->device))) {
    PBX_CUDA_CHECK(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    if (h->device < 32) attr_set.fetch_or(1u << h->device, std::memory_order_release);
  }
  const int64_t units = h->sm_count / CG;   // persistent: one CTA (or CTA pair) per SM (pair)
  const int64_t groups = p_in.total_tiles < units ? p_in.total_tiles : units;
  TcParams p = p_in;
  // more tiles than groups: hand them out dynamically (the counters live in the handle, re-armed by the kernel itself)
  p.dynamic = (h->tile_sched != nullptr && h->dynamic_sched && p.total_tiles > groups &&
               p.total_tiles < ((int64_t)1 << 31)) ? 1 : 0;
  p.sched = h->tile_sched;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(groups * CG));
  cfg.blockDim = dim3(Cfg::NUM_THREADS);
  cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
  cfg.stream = h->stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  // programmatic dependent launch: this grid may be scheduled while the previous kernel of the stream drains; the
  // kernel's griddepcontrol.wait sits after its prologue (barrier init, TMEM allocation, descriptor prefetch)
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = h->pdl ? 2 : 1;
  PBX_CUDA_CHECK(h, cudaLaunchKernelEx(&cfg, kern, tm.a, tm.b, tm.c, tm.alo, tm.blo, tm.push, p));
  h->last_grid_ctas = (int)(groups * CG);
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

}  // namespace

// one entry point per instantiation unit (gemm_tc_inst_*.cu)
#define PBX_TC_INST_DECL(name) \
  int name(pbx_handle_t h, int cg, int bn, bool a_mn, bool b_mn, const TcMaps& tm, const TcParams& p)
#define PBX_TC_INST_DEFINE(name, TIn, TOut, PRE) \
  PBX_TC_INST_DECL(name) { return launch_cfg_pre<TIn, TOut, PRE>(h, cg, bn, a_mn, b_mn, tm, p); }
