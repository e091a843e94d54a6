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
#include "gemm_tcgen05_kernel.cuh"

// instantiation units (16 kernels each: 4 tile configurations x 4 operand-major combinations)
PBX_TC_INST_DECL(pbx_tc_inst_f32_pre0);
PBX_TC_INST_DECL(pbx_tc_inst_f32_pre1);
PBX_TC_INST_DECL(pbx_tc_inst_f32_pre2);
PBX_TC_INST_DECL(pbx_tc_inst_f32_pre3);
PBX_TC_INST_DECL(pbx_tc_inst_f32_pre4);
PBX_TC_INST_DECL(pbx_tc_inst_f16);
PBX_TC_INST_DECL(pbx_tc_inst_f16f32);
PBX_TC_INST_DECL(pbx_tc_inst_bf16);
PBX_TC_INST_DECL(pbx_tc_inst_bf16f32);

namespace {

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
// direct-mapped cache of encoded maps in the handle: `key` holds every argument of the encode call
bool tmap_cache_get(pbx_handle_t h, const uint64_t (&key)[10], CUtensorMap* out, PbxTmapCacheEntry** slot) {
  uint64_t hsh = 1469598103934665603ull;
  for (int i = 0; i < 10; ++i) { hsh ^= key[i]; hsh *= 1099511628211ull; }
  PbxTmapCacheEntry& e = h->tmap_cache[(hsh >> 20) % 64];
  *slot = &e;
  if (!e.valid) return false;
  for (int i = 0; i < 10; ++i)
    if (e.key[i] != key[i]) return false;
  *out = e.map;
  return true;
}
void tmap_cache_put(PbxTmapCacheEntry* slot, const uint64_t (&key)[10], const CUtensorMap& map) {
  for (int i = 0; i < 10; ++i) slot->key[i] = key[i];
  slot->map = map;
  slot->valid = true;
}

bool make_operand_map(pbx_handle_t h, CUtensorMap* out, int es, CUtensorMapDataType dt, const void* ptr, int64_t mn,
                      int64_t k, int64_t ld, int64_t batch, int64_t stride, bool kcontig, int box_mn) {
  auto fn = get_encode_fn();
  if (!fn) return false;
  const uint64_t key[10] = {1, (uint64_t)(uintptr_t)ptr, ((uint64_t)es << 32) | (uint64_t)dt, (uint64_t)mn, (uint64_t)k,
                            (uint64_t)ld, (uint64_t)batch, (uint64_t)stride, (uint64_t)kcontig, (uint64_t)box_mn};
  PbxTmapCacheEntry* slot = nullptr;
  if (tmap_cache_get(h, key, out, &slot)) return true;
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
  if (r == CUDA_SUCCESS) tmap_cache_put(slot, key, *out);
  return r == CUDA_SUCCESS;
}

// Tensor map of the bf16 copies of one fp32 operand (PRE == 3): the hi copies of all batch entries, then the lo copies,
// along z (z = which * copies + b; zstride elements apart); 32-element (64-byte) rows, 64B swizzle.
bool make_split16_map(pbx_handle_t h, CUtensorMap* out, const void* ptr, int64_t mn, int64_t k, int64_t ld16, int64_t copies,
                      int64_t zstride, bool kcontig, int box_mn) {
  auto fn = get_encode_fn();
  if (!fn) return false;
  const uint64_t key[10] = {2, (uint64_t)(uintptr_t)ptr, (uint64_t)mn, (uint64_t)k, (uint64_t)ld16, (uint64_t)copies,
                            (uint64_t)zstride, (uint64_t)kcontig, (uint64_t)box_mn, 0};
  PbxTmapCacheEntry* slot = nullptr;
  if (tmap_cache_get(h, key, out, &slot)) return true;
  cuuint64_t dims[3] = {(cuuint64_t)(kcontig ? k : mn), (cuuint64_t)(kcontig ? mn : k), (cuuint64_t)(2 * copies)};
  cuuint64_t strides[2] = {(cuuint64_t)ld16 * 2, (cuuint64_t)zstride * 2};
  cuuint32_t box[3] = {32, (cuuint32_t)(kcontig ? box_mn : 32), 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r == CUDA_SUCCESS) tmap_cache_put(slot, key, *out);
  return r == CUDA_SUCCESS;
}

// Tensor map of C for the TMA-store epilogue: 32 x 32 boxes of the column-major output, no swizzle.
bool make_c_map(pbx_handle_t h, CUtensorMap* out, int es, CUtensorMapDataType dt, void* ptr, int64_t m, int64_t n, int64_t ld,
                int64_t batch, int64_t stride, int box_m = 32) {
  auto fn = get_encode_fn();
  if (!fn) return false;
  const uint64_t key[10] = {3, (uint64_t)(uintptr_t)ptr, ((uint64_t)es << 32) | (uint64_t)dt, (uint64_t)m, (uint64_t)n,
                            (uint64_t)ld, (uint64_t)batch, (uint64_t)stride, (uint64_t)box_m, 0};
  PbxTmapCacheEntry* slot = nullptr;
  if (tmap_cache_get(h, key, out, &slot)) return true;
  const bool batched = batch > 1;
  cuuint64_t dims[3] = {(cuuint64_t)m, (cuuint64_t)n, (cuuint64_t)(batched ? batch : 1)};
  cuuint64_t strides[2] = {(cuuint64_t)ld * es, (cuuint64_t)(batched ? stride : ld * n) * es};
  cuuint32_t box[3] = {(cuuint32_t)box_m, 32, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(out, dt, 3, ptr, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r == CUDA_SUCCESS) tmap_cache_put(slot, key, *out);
  return r == CUDA_SUCCESS;
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
  const PbxKnobs& kn = h->knobs;
  const bool want_swap = c.m <= 64 && c.n > c.m && kn.tc_swap != 0 && c.n_extra == 0;
  const int fcg = kn.tc_cg, fbn = kn.tc_bn;   // PBX_TC_CONFIG (testing)
  const bool force = (fcg == 1 || fcg == 2) && (fbn == 128 || (fbn == 256 && fcg == 2));
  if (force) {
    plan.cg = fcg; plan.bn = fbn;
  } else if (want_swap) {
    plan.cg = 1; plan.bn = 64; plan.swap = true;
  } else if (kn.plan_model != 0) {
    // Cost model (round 2; replaces the ">= 0.6 of a wave" threshold, which left 1.16-wave schedules and rejected
    // 0.59-wave ones: 384 x 5408 x 3456 ran at 53 TFLOP/s on 86 half-tiles in two rounds).  For every tile
    // configuration and K-slice count the duration is estimated in units of one K block of a CTA pair on a 256x256 tile:
    //     rounds(tiles * slices / units) * (K blocks per slice + fill/drain) * cost per K block  +  reduce pass
    // The constants are fitted to tools/plan_probe.py runs on a B200 (profiles/r02/plan_probe_*.jsonl: twelve mid-size
    // shapes under every configuration x slice count; the model's pick is within 0.5 % (fp32) / 1.7 % (bf16) of the best
    // measured choice in the geometric mean).  cost per K block per unit: fp32 1.0 / 0.79 / 0.94 (0.80 / 0.63 / 0.75 us
    // measured: both fp32 split forms run near 0.8 us per K block of a pair tile), 16-bit 1.0 / 0.9 / 0.8 (0.25 us for
    // the pair tile); `fill` is the cost of an extra tile in a running pipeline (the per-launch fixed cost, ~12 us of
    // launch + prologue + first loads + last epilogue, is the same for every choice and left out).
    const bool f32 = pbx_in_size(c.dtype) == 4;
    const double t_kb = f32 ? 0.80e-6 : 0.25e-6;     // seconds per K block of a 256x256 pair tile
    const double fill = f32 ? 0.5 : 1.0;
    const double cost_f32[3] = {1.0, 0.79, 0.94}, cost_16[3] = {1.0, 0.90, 0.80};
    const double* cost_kb = f32 ? cost_f32 : cost_16;
    double best = 1e300;
    for (int ci = 0; ci < 3; ++ci) {
      const Cand& cd = cands[ci];
      if (!usable(cd)) continue;
      const int64_t units = h->sm_count / cd.cg, tiles = tiles_of(cd);
      int64_t smax = 1;
      if (c.n_extra == 0 && h->forced_split_k == 0 && kb >= 16) {
        smax = (2 * units) / tiles;
        if (smax > kb / 4) smax = kb / 4;
        if (smax < 1) smax = 1;
      }
      for (int64_t sl = 1; sl <= smax; ++sl) {
        const int64_t kbps = (kb + sl - 1) / sl, rounds = (tiles * sl + units - 1) / units;
        double t = (double)rounds * ((double)kbps + fill) * cost_kb[ci];
        if (sl > 1)   // partial sums: written once, read once, plus the reduce launch
          t += (((double)(sl + 1) * 4.0 * (double)c.m * (double)c.n * (double)c.batch) / 5.0e12 + 1.5e-6) / t_kb;
        if (t < best * 0.97) {   // ties go to the earlier (larger-tile, fewer-slice) choice
          best = t; plan.cg = cd.cg; plan.bn = cd.bn; plan.slices = (int)sl;
        }
      }
    }
    if (h->forced_split_k == 0 && c.n_extra == 0) {
      const int64_t kbps = (kb + plan.slices - 1) / plan.slices;
      plan.slices = (int)((kb + kbps - 1) / kbps);   // drop empty trailing slices
      return plan;
    }
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
  // HBM-bound 16-bit shapes (BASELINE cfg4: 4096 x 256^3, 85 flop/B).  With the STATIC tile schedule 148 independent
  // 128x128 tiles kept DRAM busier than 74 CTA pairs on 256x256 tiles (round 1: 529 vs 518 TFLOP/s).  With the dynamic
  // schedule the pairs win (round 2, same box, burst: 548 vs 490 TFLOP/s = 6.42 vs 5.74 TB/s; cuBLAS 571): a 256x256
  // tile loads every operand byte exactly once, while the four 128x128 tiles of a batch entry load each panel twice
  // and lean on L2 to merge the copies (ncu: L2 hit rate 47 %, i.e. half of the load traffic is duplicate).
  if (!force && !plan.swap && pbx_in_size(c.dtype) == 2 && !h->dynamic_sched) {
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
    // one round of work items when the cost model is on (a second, partly filled round of very short tiles costs a
    // whole pipeline fill: 64 x 147 x 13225 on 2 x 103 items 35 us, on one round 27 us); round 1's rule otherwise
    slices = ((h->knobs.plan_model != 0 ? 1 : 2) * units) / tiles;
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
    const int pre_env = h->knobs.tf32_presplit;
    const double flops = 2.0 * (double)c.m * (double)c.n * (double)c.k * (double)c.batch;
    const double byts = 4.0 * ((double)c.m * c.k + (double)c.k * c.n + (double)c.m * c.n) * (double)c.batch;
    pre = (pre_env >= 0) ? (pre_env != 0) : (flops >= 5e8 && flops / byts >= 256.0);
    // tf32 + 2 x bf16 (TcCfg's PRE == 3; PBX_F32_SPLIT16=0 falls back to the fp32 lo pre-split): takes the place of the
    // fp32 lo pre-split on the shapes that would get it; same pooled buffers (2 x 2 bytes per element instead of 4),
    // the copies laid out [hi of every batch entry | lo of every batch entry].  Measured on B200 (round 2): SGEMM
    // 8192^3 270 -> 329 TFLOP/s, 16384^3 (power-capped) 197 -> 257, same <= 1e-5 error bound.
    if (pre && h->knobs.f32_split16 != 0) {
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
  if (!make_operand_map(h, &tm.a, es, dt, X.p, X.mn, c.k, X.ld, c.batch, X.st, !a_mn, BM) ||
      !make_operand_map(h, &tm.b, es, dt, Y.p, Y.mn, c.k, Y.ld, c.batch, Y.st, !b_mn, bn / cg)) {
    h->last_error = "cuTensorMapEncodeTiled failed";
    return PBX_ERR_CUDA;
  }
  tm.alo = tm.a; tm.blo = tm.b;   // placeholders when unused (never dereferenced)
  if (split16) {
    if (!make_split16_map(h, &tm.alo, h->lo[0], X.mn, c.k, s16_ld[0], s16_copies[0], s16_z[0], !a_mn, BM) ||
        !make_split16_map(h, &tm.blo, h->lo[1], Y.mn, c.k, s16_ld[1], s16_copies[1], s16_z[1], !b_mn, bn / cg)) {
      h->last_error = "cuTensorMapEncodeTiled failed (bf16 split copies)";
      return PBX_ERR_CUDA;
    }
  } else if (pre && (!make_operand_map(h, &tm.alo, es, dt, lo_ptr[0], X.mn, c.k, X.ld, c.batch, X.st, !a_mn, BM) ||
                     !make_operand_map(h, &tm.blo, es, dt, lo_ptr[1], Y.mn, c.k, Y.ld, c.batch, Y.st, !b_mn, bn / cg))) {
    h->last_error = "cuTensorMapEncodeTiled failed";
    return PBX_ERR_CUDA;
  }
  // no pre-pass: the splitter warps of the kernel make the lo halves -- as bf16 tiles (mode 4: tf32 + 2 x bf16, two
  // tf32-MMA times per k-step) unless PBX_F32_SPLIT16=0 asks for the 3xTF32 form (mode 0)
  const bool inkernel16 = f32 && !tf32x1 && !pre && h->knobs.f32_split16 != 0;
  const int pre_mode = tf32x1 ? 2 : (split16 ? 3 : (pre ? 1 : (inkernel16 ? 4 : 0)));
  h->last_presplit = pre_mode;
  TcParams p;
  p.C = c.C; p.ws = (float*)h->ws;
  p.M = X.mn; p.N = Y.mn; p.K = c.k; p.ldc = c.ldc; p.sc = c.sc;
  p.alpha = (float)c.alpha; p.beta = (float)c.beta;
  p.batch = (int)c.batch; p.slices = slices;
  p.m_tiles = (int)((p.M + BM * cg - 1) / (BM * cg));
  p.n_tiles = (int)((p.N + bn - 1) / bn);
  p.group_m = h->knobs.group_m > 0 ? h->knobs.group_m : (cg == 2 ? 8 : 16);
  p.kb_total = (int)((c.k + bk - 1) / bk);
  p.kb_per_slice = (p.kb_total + slices - 1) / slices;
  // fp32: the tensor core truncates when it adds into its fp32 accumulator, so an accumulation
  // chain is limited to kb_per_chunk blocks of 32 (default 16 -> 512 k: ~4e-6 relative bias)
  const int chunk_env = h->knobs.tf32_chunk_kb;
  p.kb_per_chunk = f32 ? (chunk_env > 0 ? chunk_env : 16) : (1 << 30);
  // default: raw fp32 tile as the hi operand (verified on B200: kind::tf32 ignores the low 13 mantissa bits)
  p.raw_hi = h->knobs.tf32_raw_hi;
  p.a_batched = (c.batch > 1 && X.st > 0) ? 1 : 0;
  p.b_batched = (c.batch > 1 && Y.st > 0) ? 1 : 0;
  const int64_t eo = (int64_t)pbx_out_size(c.dtype);
  p.c_vec = (((uintptr_t)c.C % 16 == 0) && (c.ldc * eo) % 16 == 0 && (c.sc * eo) % 16 == 0) ? 1 : 0;
  p.total_tiles = (int64_t)p.m_tiles * p.n_tiles * c.batch * slices;
  p.n_extra = c.n_extra;
  for (int x = 0; x < 7; ++x) p.Cx[x] = x < c.n_extra ? c.c_extra[x] : nullptr;

  // 16-bit C with beta == 0 leaves through shared memory + TMA stores when C is TMA-legal
  tm.c = tm.a;  // placeholder when unused (never dereferenced)
  tm.push.local = tm.a;
  for (int x = 0; x < 7; ++x) tm.push.peer[x] = tm.a;
  p.tma_store = 0;
  p.push = 0;
  p.push_pace = h->knobs.multicast_pace;
  p.wait_hint_ns = h->knobs.wait_hint_ns;
  const bool out16 = (c.dtype == PBX_F16 || c.dtype == PBX_BF16);
  auto c_legal = [&](const void* ptr) {
    return ((uintptr_t)ptr % 16 == 0) && (c.ldc * eo) % 16 == 0 && (c.batch == 1 || (c.sc * eo) % 16 == 0) &&
           c.ldc * eo < ((int64_t)1 << 40) && c.sc * eo < ((int64_t)1 << 40);
  };
  bool peers_legal = true;
  for (int x = 0; x < c.n_extra; ++x) peers_legal = peers_legal && c_legal(c.c_extra[x]);
  if (out16 && c.beta == 0.0 && slices == 1 && h->knobs.tma_store != 0 && c_legal(c.C) &&
      peers_legal) {
    bool ok = make_c_map(h, &tm.c, 2, dt, c.C, c.m, c.n, c.ldc, c.batch, c.sc);
    // multicast GEMM: the staging tiles of the TMA-store epilogue also go to every peer's C
    for (int x = 0; x < c.n_extra && ok; ++x) ok = make_c_map(h, &tm.push.peer[x], 2, dt, c.c_extra[x], c.m, c.n, c.ldc, c.batch, c.sc);
    if (ok) p.tma_store = 1;
  }
  // multicast GEMM with 32-bit outputs: asynchronous peer copies by the pusher warp (128 x 32 boxes read back from the
  // local C); PBX_MULTICAST_PUSH=0 keeps the round-1 form (the epilogue warps store to every copy themselves)
  if (c.n_extra > 0 && eo == 4 && slices == 1 && !plan.swap && h->knobs.multicast_push != 0 && c_legal(c.C) && peers_legal) {
    bool ok = make_c_map(h, &tm.push.local, 4, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, c.C, c.m, c.n, c.ldc, c.batch, c.sc, BM);
    for (int x = 0; x < c.n_extra && ok; ++x)
      ok = make_c_map(h, &tm.push.peer[x], 4, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, c.c_extra[x], c.m, c.n, c.ldc, c.batch, c.sc, BM);
    if (ok) p.push = 1;
  }

  switch (c.dtype) {
    case PBX_F32:
      if (pre_mode == 1) return pbx_tc_inst_f32_pre1(h, cg, bn, a_mn, b_mn, tm, p);
      if (pre_mode == 2) return pbx_tc_inst_f32_pre2(h, cg, bn, a_mn, b_mn, tm, p);
      if (pre_mode == 3) return pbx_tc_inst_f32_pre3(h, cg, bn, a_mn, b_mn, tm, p);
      if (pre_mode == 4) return pbx_tc_inst_f32_pre4(h, cg, bn, a_mn, b_mn, tm, p);
      return pbx_tc_inst_f32_pre0(h, cg, bn, a_mn, b_mn, tm, p);
    case PBX_F16: return pbx_tc_inst_f16(h, cg, bn, a_mn, b_mn, tm, p);
    case PBX_F16_F32: return pbx_tc_inst_f16f32(h, cg, bn, a_mn, b_mn, tm, p);
    case PBX_BF16: return pbx_tc_inst_bf16(h, cg, bn, a_mn, b_mn, tm, p);
    case PBX_BF16_F32: return pbx_tc_inst_bf16f32(h, cg, bn, a_mn, b_mn, tm, p);
  }
  return PBX_ERR_INVALID_ARG;
}
