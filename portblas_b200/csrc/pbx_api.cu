// pbx_api.cu -- C-ABI entry points, front-end rules and the kernel selector.
//
// Front-end restated from reference src/interface/gemm_interface.hpp:105-185
// (alpha==0 shortcut first, then trans/stride validation, then beta==0
// specialisation).  The selector replaces
// src/interface/blas3/backend/nvidia_gpu.hpp:40-260.
#include <ctype.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "pbx_internal.cuh"

extern "C" {

const char* pbx_status_string(int status) {
  switch (status) {
    case PBX_OK: return "ok";
    case PBX_ERR_INVALID_TRANSA: return "invalid _TransA";
    case PBX_ERR_INVALID_TRANSB: return "invalid _TransB";
    case PBX_ERR_INVALID_STRIDEC: return "invalid _stridec";
    case PBX_ERR_INVALID_STRIDEA: return "invalid _stridea";
    case PBX_ERR_INVALID_STRIDEB: return "invalid _strideb";
    case PBX_ERR_INVALID_ARG: return "invalid argument";
    case PBX_ERR_CUDA: return "CUDA error";
    case PBX_ERR_NO_DEVICE: return "no sm_100 CUDA device (this library has no CPU fallback)";
    case PBX_ERR_WORKSPACE: return "workspace allocation failed";
    case PBX_ERR_INVALID_UPLO: return "invalid _uplo";
    case PBX_ERR_INVALID_SIDE: return "invalid _side";
    case PBX_ERR_TRSM_SIZE: return "invalid matrix size argument";
    case PBX_ERR_TRSM_SIDE: return "invalid Side argument";
    case PBX_ERR_TRSM_UPLO: return "invalid Triangle argument";
    case PBX_ERR_TRSM_TRANS: return "invalid Transpose argument";
    case PBX_ERR_TRSM_DIAG: return "invalid Diagonal argument";
  }
  return "unknown status";
}

int pbx_create(pbx_handle_t* out, int device_ordinal, void* cuda_stream) {
  if (!out) return PBX_ERR_INVALID_ARG;
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || device_ordinal >= ndev ||
      device_ordinal < 0) {
    fprintf(stderr, "[pbx_gemm] %s\n", pbx_status_string(PBX_ERR_NO_DEVICE));
    return PBX_ERR_NO_DEVICE;
  }
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device_ordinal) != cudaSuccess) return PBX_ERR_CUDA;
  if (prop.major != 10) {
    fprintf(stderr, "[pbx_gemm] device %d is sm_%d%d; this library is built for sm_100a only\n",
            device_ordinal, prop.major, prop.minor);
    return PBX_ERR_NO_DEVICE;
  }
  pbx_handle_s* h = new pbx_handle_s();
  h->device = device_ordinal;
  h->stream = (cudaStream_t)cuda_stream;
  h->sm_count = prop.multiProcessorCount;
  h->cc_major = prop.major;
  h->cc_minor = prop.minor;
  const char* fk = getenv("PBX_FORCE_KERNEL");
  if (fk) h->forced_kernel = atoi(fk);
  pbx_reload_env(h);
  {
    PbxDeviceGuard guard(device_ordinal);
    if (!guard.ok() || cudaMalloc(&h->tile_sched, 256) != cudaSuccess ||
        cudaMemset(h->tile_sched, 0, 256) != cudaSuccess) {
      delete h;
      return PBX_ERR_CUDA;
    }
  }
  *out = h;
  return PBX_OK;
}

// (Re-)read the PBX_* testing / tuning switches into the handle.  pbx_create calls it; a caller that changes the
// environment afterwards (the test-suite does, per case) calls it again.
int pbx_reload_env(pbx_handle_t h) {
  if (!h) return PBX_ERR_INVALID_ARG;
  auto geti = [](const char* name, int dflt) { const char* v = getenv(name); return v ? atoi(v) : dflt; };
  PbxKnobs k;
  k.tc_swap = geti("PBX_TC_SWAP", 1);
  const char* cfg = getenv("PBX_TC_CONFIG");
  if (!(cfg && sscanf(cfg, "%d,%d", &k.tc_cg, &k.tc_bn) == 2)) { k.tc_cg = 0; k.tc_bn = 0; }
  k.plan_model = geti("PBX_PLAN_MODEL", 1);
  k.tf32_presplit = geti("PBX_TF32_PRESPLIT", -1);
  k.f32_split16 = geti("PBX_F32_SPLIT16", 1);
  k.tf32_chunk_kb = geti("PBX_TF32_CHUNK_KB", 0);
  k.tf32_raw_hi = geti("PBX_TF32_RAW_HI", 1);
  k.multicast_pace = geti("PBX_MULTICAST_PACE", 1);
  k.multicast_push = geti("PBX_MULTICAST_PUSH", 1);
  k.wait_hint_ns = (unsigned)geti("PBX_WAIT_HINT_NS", 0);
  k.tma_store = geti("PBX_TMA_STORE", 1);
  k.ilv_via_strided = geti("PBX_ILV_VIA_STRIDED", -1);
  k.group_m = geti("PBX_GROUP_M", 0);
  h->knobs = k;
  h->dynamic_sched = geti("PBX_DYNAMIC_SCHED", 1) != 0;
  h->pdl = geti("PBX_PDL", 1) != 0;
  h->pdl_reduce = geti("PBX_PDL_REDUCE", 1) != 0;
  return PBX_OK;
}

int pbx_destroy(pbx_handle_t h) {
  if (!h) return PBX_ERR_INVALID_ARG;
  PbxDeviceGuard guard(h);
  cudaStreamSynchronize(h->stream);
  if (h->ws) cudaFree(h->ws);
  if (h->tile_sched) cudaFree(h->tile_sched);
  for (int i = 0; i < 3; ++i)
    if (h->stage[i]) cudaFree(h->stage[i]);
  for (int i = 0; i < 2; ++i)
    if (h->pack[i]) cudaFree(h->pack[i]);
  for (int i = 0; i < 4; ++i)
    if (h->aux[i]) cudaFree(h->aux[i]);
  for (int i = 0; i < 2; ++i)
    if (h->lo[i]) cudaFree(h->lo[i]);
  for (auto& kv : h->ipc_open) cudaIpcCloseMemHandle(kv.second);
  for (cudaEvent_t e : h->events) cudaEventDestroy(e);
  if (h->s_in) cudaStreamDestroy(h->s_in);
  if (h->s_out) cudaStreamDestroy(h->s_out);
  delete h;
  return PBX_OK;
}

int pbx_set_stream(pbx_handle_t h, void* s) { if (!h) return PBX_ERR_INVALID_ARG; h->stream = (cudaStream_t)s; return PBX_OK; }
void* pbx_get_stream(pbx_handle_t h) { return h ? (void*)h->stream : nullptr; }
int pbx_get_num_compute_units(pbx_handle_t h) { return h ? h->sm_count : 0; }
int pbx_get_device(pbx_handle_t h) { return h ? h->device : -1; }
int pbx_synchronize(pbx_handle_t h) {
  if (!h) return PBX_ERR_INVALID_ARG;
  PBX_CUDA_CHECK(h, cudaStreamSynchronize(h->stream));
  return PBX_OK;
}
const char* pbx_last_error(pbx_handle_t h) { return h ? h->last_error.c_str() : "null handle"; }
int pbx_set_forced_kernel(pbx_handle_t h, int k) { if (!h) return PBX_ERR_INVALID_ARG; h->forced_kernel = k; return PBX_OK; }
int pbx_set_split_k(pbx_handle_t h, int s) { if (!h || s < 0) return PBX_ERR_INVALID_ARG; h->forced_split_k = s; return PBX_OK; }
int pbx_last_kernel(pbx_handle_t h) { return h ? h->last_kernel : PBX_KERNEL_NONE; }
int pbx_last_split_k(pbx_handle_t h) { return h ? h->last_split_k : 0; }
int pbx_last_repack(pbx_handle_t h) { return h ? h->last_repack : 0; }
int pbx_last_presplit(pbx_handle_t h) { return h ? h->last_presplit : 0; }
int64_t pbx_launch_count(pbx_handle_t h) { return h ? h->launches : 0; }
int64_t pbx_workspace_bytes(pbx_handle_t h) { return h ? h->ws_bytes : 0; }

}  // extern "C"

// Pooled temporaries grow through the stream-ordered allocator: the old buffer is released and the new one obtained IN
// stream order (cudaFreeAsync / cudaMallocAsync), so work already queued on the stream keeps its buffer and the call
// stays asynchronous (round 1 synchronised the stream here, a hidden wait inside an "asynchronous" API).
static int pbx_grow(pbx_handle_t h, void** buf, int64_t* cap, int64_t bytes, const char* what) {
  if (bytes <= *cap) return PBX_OK;
  if (*buf) {
    if (cudaFreeAsync(*buf, h->stream) != cudaSuccess) {
      cudaGetLastError();
      h->last_error = std::string(what) + " free failed";
      return PBX_ERR_WORKSPACE;
    }
    *buf = nullptr; *cap = 0;
  }
  const int64_t rounded = ((bytes + (1 << 20) - 1) >> 20) << 20;
  if (cudaMallocAsync(buf, (size_t)rounded, h->stream) != cudaSuccess) {
    cudaGetLastError();
    *buf = nullptr;
    h->last_error = std::string(what) + " allocation failed";
    return PBX_ERR_WORKSPACE;
  }
  *cap = rounded;
  return PBX_OK;
}

int pbx_ensure_workspace(pbx_handle_t h, int64_t bytes) { return pbx_grow(h, &h->ws, &h->ws_bytes, bytes, "workspace"); }

int pbx_ensure_aux(pbx_handle_t h, int i, int64_t bytes) {
  if (i < 0 || i >= 4) return PBX_ERR_INVALID_ARG;
  return pbx_grow(h, &h->aux[i], &h->aux_bytes[i], bytes, "temporary");
}

int pbx_ensure_lo(pbx_handle_t h, int i, int64_t bytes) {
  if (i < 0 || i >= 2) return PBX_ERR_INVALID_ARG;
  return pbx_grow(h, &h->lo[i], &h->lo_bytes[i], bytes, "fp32 split copy");
}

// ---- split-K policy -------------------------------------------------------------
// The reference rule is depth = ceil(4*CUs / (ceil(M/tm)*ceil(N/tn))) and "no split when
// depth == 1 || K <= 2048" (gemm_partial_local.hpp:191-199, portblas_handle.hpp:323).  On
// B200 the quantity that matters is wave quantisation of 148 SMs: split only when the
// output tiles cannot fill the machine and every slice still has a deep K loop.
static int choose_split_k(pbx_handle_t h, const PbxGemmCall& c, int tile_m, int tile_n,
                          int64_t k_block, int64_t min_k_per_slice) {
  if (h->forced_split_k == 1) return 1;
  const int64_t kb = (c.k + k_block - 1) / k_block;
  int slices = 1;
  if (h->forced_split_k > 1) {
    slices = h->forced_split_k;
  } else {
    const int64_t tiles = ((c.m + tile_m - 1) / tile_m) * ((c.n + tile_n - 1) / tile_n) * c.batch;
    if (tiles * 2 > h->sm_count || c.k < 2 * min_k_per_slice) return 1;
    // as many whole waves of CTAs as the K depth allows, at most 2 waves
    int64_t want = (2 * (int64_t)h->sm_count) / tiles;
    int64_t maxs = c.k / min_k_per_slice;
    slices = (int)(want < maxs ? want : maxs);
    if (slices < 1) slices = 1;
  }
  if (slices > kb) slices = (int)kb;
  if (slices < 1) slices = 1;
  // drop empty trailing slices
  const int64_t kbps = (kb + slices - 1) / slices;
  slices = (int)((kb + kbps - 1) / kbps);
  return slices;
}

// An operand the TMA cannot address (odd leading dimension, base or batch stride off 16 bytes -- the
// reference's offset / odd-ld test grids and many rows of its benchmark sweeps) is first copied to a
// 16-byte-legal layout in a pooled buffer; the copy is one HBM-bound pass over that operand, after which
// the call runs on the tensor cores instead of the CUDA-core kernel.
static int ensure_pack(pbx_handle_t h, int i, int64_t bytes) {
  return pbx_grow(h, &h->pack[i], &h->pack_bytes[i], bytes, "packed operand");
}

static bool repack_for_tma(pbx_handle_t h, PbxGemmCall& c) {
  const int64_t es = (int64_t)pbx_in_size(c.dtype);
  const int64_t quantum = 16 / es;
  struct Op { const void** p; int64_t* ld; int64_t* st; int64_t rows, cols; };
  Op ops[2] = {{&c.A, &c.lda, &c.sa, c.ta ? c.k : c.m, c.ta ? c.m : c.k},
               {&c.B, &c.ldb, &c.sb, c.tb ? c.n : c.k, c.tb ? c.k : c.n}};
  int mask = 0;
  for (int i = 0; i < 2; ++i) {
    Op& o = ops[i];
    if (pbx_tma_operand_ok(c.dtype, *o.p, *o.ld, *o.st)) continue;
    const int64_t ld2 = (o.rows + quantum - 1) / quantum * quantum;
    const int64_t st2 = (*o.st > 0) ? ld2 * o.cols : 0;
    const int64_t copies = (*o.st > 0) ? c.batch : 1;
    const int64_t bytes = ld2 * o.cols * copies * es;
    if (bytes > ((int64_t)16 << 30) || ensure_pack(h, i, bytes) != PBX_OK) return false;
    if (pbx_launch_repack(h, (int)es, *o.p, h->pack[i], o.rows, o.cols, *o.ld, ld2, *o.st, st2, copies) != PBX_OK)
      return false;
    *o.p = h->pack[i]; *o.ld = ld2; *o.st = st2;
    mask |= 1 << i;
  }
  h->last_repack = mask;
  return true;
}

static int run_gemm(pbx_handle_t h, const PbxGemmCall& c_in, int batch_type);

// Interleaved batches through the tensor cores: one HBM-bound pass re-lays A and B (and C when beta != 0) out as
// strided batches with 16-byte-legal leading dimensions in pooled buffers, the ordinary strided path runs, one more
// pass writes C back interleaved.  The reference runs a dedicated CUDA-core kernel (gemm_interleaved.hpp:219-312), and
// so does gemm_interleaved_kernel here; that one is FMA / issue bound at 0.06-0.09 of the HBM roof, while the three
// extra passes cost about twice the algorithmic bytes.  Taken when the batch is large enough to fill the transposes'
// 32-entry tiles and the matrices are not tiny; PBX_ILV_VIA_STRIDED=0 keeps the dedicated kernel.
static bool interleaved_via_strided(pbx_handle_t h, const PbxGemmCall& c, int* status) {
  const int env = h->knobs.ilv_via_strided;   // PBX_ILV_VIA_STRIDED: 0 never, 1 always (testing)
  if (env == 0) return false;
  if (h->forced_kernel == PBX_KERNEL_INTERLEAVED) return false;
  const double flops = 2.0 * (double)c.m * (double)c.n * (double)c.k * (double)c.batch;
  if (env != 1 && (c.batch < 16 || flops < 2e8 || c.m * c.n < 1024)) return false;
  const int64_t es = (int64_t)pbx_in_size(c.dtype), eo = (int64_t)pbx_out_size(c.dtype);
  const int64_t a_rows = c.ta ? c.k : c.m, a_cols = c.ta ? c.m : c.k;
  const int64_t b_rows = c.tb ? c.n : c.k, b_cols = c.tb ? c.k : c.n;
  if (a_cols > 65535 || b_cols > 65535 || c.n > 65535 || (c.batch + 31) / 32 > 65535) return false;
  auto up = [](int64_t v, int64_t q) { return (v + q - 1) / q * q; };
  const int64_t lda2 = up(a_rows, 16 / es), ldb2 = up(b_rows, 16 / es), ldc2 = up(c.m, 16 / eo);
  const int64_t sa2 = lda2 * a_cols, sb2 = ldb2 * b_cols, sc2 = ldc2 * c.n;
  if (pbx_ensure_aux(h, 0, sa2 * c.batch * es) != PBX_OK || pbx_ensure_aux(h, 1, sb2 * c.batch * es) != PBX_OK ||
      pbx_ensure_aux(h, 2, sc2 * c.batch * eo) != PBX_OK)
    return false;   // no room for the copies: the dedicated kernel works in place
  int st = pbx_launch_ilv_relayout(h, (int)es, c.A, h->aux[0], a_rows, a_cols, c.lda, lda2, sa2, c.batch, true);
  if (st == PBX_OK) st = pbx_launch_ilv_relayout(h, (int)es, c.B, h->aux[1], b_rows, b_cols, c.ldb, ldb2, sb2, c.batch, true);
  if (st == PBX_OK && c.beta != 0.0)
    st = pbx_launch_ilv_relayout(h, (int)eo, c.C, h->aux[2], c.m, c.n, c.ldc, ldc2, sc2, c.batch, true);
  if (st == PBX_OK) {
    PbxGemmCall s = c;
    s.A = h->aux[0]; s.B = h->aux[1]; s.C = h->aux[2];
    s.lda = lda2; s.ldb = ldb2; s.ldc = ldc2; s.sa = sa2; s.sb = sb2; s.sc = sc2;
    st = run_gemm(h, s, 0);
  }
  if (st == PBX_OK) st = pbx_launch_ilv_relayout(h, (int)eo, h->aux[2], c.C, c.m, c.n, c.ldc, ldc2, sc2, c.batch, false);
  *status = st;
  return true;
}

static int run_gemm(pbx_handle_t h, const PbxGemmCall& c_in, int batch_type) {
  PbxGemmCall c = c_in;
  int kernel = h->forced_kernel;
  h->last_repack = 0;
  // tiny problems are launch-latency bound either way; the 128-row MMA tile wastes most of its lanes
  // below ~32 rows, keep those on the CUDA-core kernel.
  const double flops = 2.0 * (double)c.m * (double)c.n * (double)c.k * (double)c.batch;
  const bool tiny = flops < 4e6 || (c.m * c.n < 64 * 64 && c.k < 256);
  // re-laying out an operand costs one pass over it: worth it once the contraction has real work
  // (and a K loop deep enough to fill an MMA K block: below that the CUDA-core kernel reading in place wins)
  const bool heavy = c.k >= 64;
  if (batch_type == 1 && c.batch > 1) {
    int st_via = PBX_OK;
    if (interleaved_via_strided(h, c, &st_via)) return st_via;   // last_kernel = the strided path's kernel
    kernel = PBX_KERNEL_INTERLEAVED;
  } else if (kernel == PBX_KERNEL_AUTO || kernel == PBX_KERNEL_INTERLEAVED) {
    if (c.dtype == PBX_F64) {
      kernel = PBX_KERNEL_DMMA;
    } else if (pbx_tcgen05_eligible(h, c)) {
      kernel = tiny ? PBX_KERNEL_SIMT : PBX_KERNEL_TCGEN05;
    } else if (!tiny && heavy && pbx_tcgen05_shape_ok(h, c) && repack_for_tma(h, c)) {
      kernel = PBX_KERNEL_TCGEN05;
    } else {
      kernel = PBX_KERNEL_SIMT;
    }
  }
  if (kernel == PBX_KERNEL_TCGEN05 && c.dtype != PBX_F64 && !pbx_tcgen05_eligible(h, c) &&
      pbx_tcgen05_shape_ok(h, c))
    repack_for_tma(h, c);   // forced tensor-core path on an unaligned operand
  if (kernel == PBX_KERNEL_TCGEN05 && (c.dtype == PBX_F64 || !pbx_tcgen05_eligible(h, c)))
    kernel = (c.dtype == PBX_F64) ? PBX_KERNEL_DMMA : PBX_KERNEL_SIMT;
  if (kernel == PBX_KERNEL_DMMA && c.dtype != PBX_F64) kernel = PBX_KERNEL_SIMT;

  h->last_kernel = kernel;
  h->last_split_k = 1;
  h->last_presplit = 0;
  int st = PBX_OK;
  if (kernel == PBX_KERNEL_INTERLEAVED) return pbx_launch_interleaved(h, c);

  int slices = 1;
  const size_t acc_size = (c.dtype == PBX_F64) ? 8 : 4;
  if (kernel == PBX_KERNEL_TCGEN05) slices = pbx_tcgen05_slices(h, c);
  else if (kernel == PBX_KERNEL_DMMA) slices = choose_split_k(h, c, 128, 128, 16, 1024);
  else slices = choose_split_k(h, c, 64, 64, 16, 256);
  if (slices > 1) {
    st = pbx_ensure_workspace(h, (int64_t)acc_size * c.m * c.n * c.batch * slices);
    if (st != PBX_OK) return st;
  }
  h->last_split_k = slices;
  if (kernel == PBX_KERNEL_TCGEN05) st = pbx_launch_tcgen05(h, c, slices);
  else if (kernel == PBX_KERNEL_DMMA) st = pbx_launch_dmma(h, c, slices);
  else st = pbx_launch_simt(h, c, slices);
  if (st != PBX_OK) return st;
  if (slices > 1) st = pbx_launch_splitk_reduce(h, c, slices);
  return st;
}

static bool valid_dtype(int d) { return d >= PBX_F32 && d <= PBX_BF16_F32; }

static double read_scalar(int dtype, const void* p) {
  return dtype == PBX_F64 ? *reinterpret_cast<const double*>(p)
                          : (double)*reinterpret_cast<const float*>(p);
}

extern "C" {

int pbx_scal_matrix(pbx_handle_t h, int dtype, int64_t m, int64_t n, const void* beta, void* C,
                    int64_t ldc, int64_t stridec, int64_t batch) {
  if (!h || !valid_dtype(dtype) || !beta || m < 0 || n < 0 || batch < 0) return PBX_ERR_INVALID_ARG;
  PBX_DEVICE_GUARD(h);
  const double b = read_scalar(dtype, beta);
  h->last_split_k = 1;
  if (b == 1.0 || m == 0 || n == 0 || batch == 0) {  // blas1_interface.hpp:498-499
    h->last_kernel = PBX_KERNEL_NONE;
    return PBX_OK;
  }
  h->last_kernel = PBX_KERNEL_SCAL;
  return pbx_launch_scal(h, dtype, m, n, b, C, ldc, stridec, batch, 0);
}

int pbx_gemm(pbx_handle_t h, int dtype, char transa, char transb, int64_t m, int64_t n, int64_t k,
             const void* alpha, const void* A, int64_t lda, int64_t stridea, const void* B,
             int64_t ldb, int64_t strideb, const void* beta, void* C, int64_t ldc, int64_t stridec,
             int64_t batch, int batch_type) {
  if (!h) return PBX_ERR_INVALID_ARG;
  if (!valid_dtype(dtype) || !alpha || !beta || m < 0 || n < 0 || k < 0 || batch < 0 ||
      (batch_type != 0 && batch_type != 1)) {
    h->last_error = "pbx_gemm: invalid argument";
    return PBX_ERR_INVALID_ARG;
  }
  PBX_DEVICE_GUARD(h);
  const double al = read_scalar(dtype, alpha);
  const double be = read_scalar(dtype, beta);
  h->last_split_k = 1;

  // (1) alpha == 0 comes first, before any validation (gemm_interface.hpp:112-139).
  if (al == 0.0) {
    if (m == 0 || n == 0 || batch == 0 || be == 1.0) { h->last_kernel = PBX_KERNEL_NONE; return PBX_OK; }
    h->last_kernel = PBX_KERNEL_SCAL;
    return pbx_launch_scal(h, dtype, m, n, be, C, ldc, stridec, batch, batch_type == 1 && batch > 1);
  }
  // (2) trans validation (gemm_interface.hpp:141-148); 'c' == 't' for real types (:150-151)
  const int ta_c = tolower((unsigned char)transa), tb_c = tolower((unsigned char)transb);
  if (ta_c != 'n' && ta_c != 't' && ta_c != 'c') return PBX_ERR_INVALID_TRANSA;
  if (tb_c != 'n' && tb_c != 't' && tb_c != 'c') return PBX_ERR_INVALID_TRANSB;
  // (3) stride validation, only for strided batches (gemm_interface.hpp:153-166)
  if (batch > 1 && batch_type == 0) {
    if (stridec < ldc * n || stridec < 0) return PBX_ERR_INVALID_STRIDEC;
    if (stridea < 0) return PBX_ERR_INVALID_STRIDEA;
    if (strideb < 0) return PBX_ERR_INVALID_STRIDEB;
  }
  if (m == 0 || n == 0 || batch == 0) { h->last_kernel = PBX_KERNEL_NONE; return PBX_OK; }
  if (k == 0) {  // BLAS: C <- beta*C
    if (be == 1.0) { h->last_kernel = PBX_KERNEL_NONE; return PBX_OK; }
    h->last_kernel = PBX_KERNEL_SCAL;
    return pbx_launch_scal(h, dtype, m, n, be, C, ldc, stridec, batch, batch_type == 1 && batch > 1);
  }
  if (!A || !B || !C) { h->last_error = "pbx_gemm: null matrix pointer"; return PBX_ERR_INVALID_ARG; }

  PbxGemmCall c;
  c.dtype = dtype; c.ta = (ta_c != 'n'); c.tb = (tb_c != 'n');
  c.m = m; c.n = n; c.k = k; c.alpha = al; c.beta = be;
  c.A = A; c.B = B; c.C = C; c.lda = lda; c.ldb = ldb; c.ldc = ldc;
  c.sa = (batch > 1) ? stridea : 0; c.sb = (batch > 1) ? strideb : 0; c.sc = (batch > 1) ? stridec : 0;
  c.batch = batch;
  return run_gemm(h, c, batch_type);
}

int pbx_sgemm(pbx_handle_t h, char ta, char tb, int64_t m, int64_t n, int64_t k, const float* alpha,
              const float* A, int64_t lda, int64_t sa, const float* B, int64_t ldb, int64_t sb,
              const float* beta, float* C, int64_t ldc, int64_t sc, int64_t batch, int bt) {
  return pbx_gemm(h, PBX_F32, ta, tb, m, n, k, alpha, A, lda, sa, B, ldb, sb, beta, C, ldc, sc, batch, bt);
}
int pbx_dgemm(pbx_handle_t h, char ta, char tb, int64_t m, int64_t n, int64_t k, const double* alpha,
              const double* A, int64_t lda, int64_t sa, const double* B, int64_t ldb, int64_t sb,
              const double* beta, double* C, int64_t ldc, int64_t sc, int64_t batch, int bt) {
  return pbx_gemm(h, PBX_F64, ta, tb, m, n, k, alpha, A, lda, sa, B, ldb, sb, beta, C, ldc, sc, batch, bt);
}
int pbx_hgemm(pbx_handle_t h, char ta, char tb, int64_t m, int64_t n, int64_t k, const float* alpha,
              const void* A, int64_t lda, int64_t sa, const void* B, int64_t ldb, int64_t sb,
              const float* beta, void* C, int64_t ldc, int64_t sc, int64_t batch, int bt) {
  return pbx_gemm(h, PBX_F16, ta, tb, m, n, k, alpha, A, lda, sa, B, ldb, sb, beta, C, ldc, sc, batch, bt);
}
int pbx_hsgemm(pbx_handle_t h, char ta, char tb, int64_t m, int64_t n, int64_t k, const float* alpha,
               const void* A, int64_t lda, int64_t sa, const void* B, int64_t ldb, int64_t sb,
               const float* beta, float* C, int64_t ldc, int64_t sc, int64_t batch, int bt) {
  return pbx_gemm(h, PBX_F16_F32, ta, tb, m, n, k, alpha, A, lda, sa, B, ldb, sb, beta, C, ldc, sc, batch, bt);
}
int pbx_bf16gemm(pbx_handle_t h, char ta, char tb, int64_t m, int64_t n, int64_t k, const float* alpha,
                 const void* A, int64_t lda, int64_t sa, const void* B, int64_t ldb, int64_t sb,
                 const float* beta, void* C, int64_t ldc, int64_t sc, int64_t batch, int bt) {
  return pbx_gemm(h, PBX_BF16, ta, tb, m, n, k, alpha, A, lda, sa, B, ldb, sb, beta, C, ldc, sc, batch, bt);
}

// ---- multicast GEMM: C <- alpha*op(A)*op(B) + beta*C written into n_dst copies of C at once ---------------------
// C_list[0] is this GPU's C (read when beta != 0); C_list[1..] are further copies with the same ldc, typically the
// same row block inside the full C of every peer GPU (pointers from pbx_ipc_import).  On the tcgen05 path the
// epilogue stores every finished tile to all copies, so the gather of an M-block sharded GEMM costs no extra pass
// and no collective: NVLink carries tile i while tile i+1 is on the tensor cores.  Other kernels (fp64, tiny or
// unaligned problems) compute locally and then push the window to the peers with 2-D device-to-device copies.
int pbx_gemm_multicast(pbx_handle_t h, int dtype, char transa, char transb, int64_t m, int64_t n, int64_t k,
                       const void* alpha, const void* A, int64_t lda, const void* B, int64_t ldb, const void* beta,
                       void* const* C_list, int n_dst, int64_t ldc) {
  if (!h) return PBX_ERR_INVALID_ARG;
  if (!C_list || n_dst < 1 || n_dst > 8 || !valid_dtype(dtype) || !alpha || !beta || m < 0 || n < 0 || k < 0) {
    h->last_error = "pbx_gemm_multicast: invalid argument";
    return PBX_ERR_INVALID_ARG;
  }
  if (n_dst == 1)
    return pbx_gemm(h, dtype, transa, transb, m, n, k, alpha, A, lda, 0, B, ldb, 0, beta, C_list[0], ldc, 0, 1, 0);
  PBX_DEVICE_GUARD(h);
  const double al = read_scalar(dtype, alpha), be = read_scalar(dtype, beta);
  const int ta_c = tolower((unsigned char)transa), tb_c = tolower((unsigned char)transb);
  const bool plain = (al != 0.0) && (ta_c == 'n' || ta_c == 't' || ta_c == 'c') &&
                     (tb_c == 'n' || tb_c == 't' || tb_c == 'c') && m > 0 && n > 0 && k > 0 && A && B && C_list[0];
  int st;
  if (plain) {
    PbxGemmCall c;
    c.dtype = dtype; c.ta = (ta_c != 'n'); c.tb = (tb_c != 'n');
    c.m = m; c.n = n; c.k = k; c.alpha = al; c.beta = be;
    c.A = A; c.B = B; c.C = C_list[0]; c.lda = lda; c.ldb = ldb; c.ldc = ldc;
    c.sa = c.sb = c.sc = 0; c.batch = 1;
    c.n_extra = n_dst - 1;
    for (int x = 1; x < n_dst; ++x) c.c_extra[x - 1] = C_list[x];
    h->last_split_k = 1;
    st = run_gemm(h, c, 0);
    if (st != PBX_OK) return st;
    if (h->last_kernel == PBX_KERNEL_TCGEN05) return PBX_OK;   // the epilogue already wrote every copy
  } else {   // shortcuts and errors: exactly pbx_gemm's behaviour on the local copy
    st = pbx_gemm(h, dtype, transa, transb, m, n, k, alpha, A, lda, 0, B, ldb, 0, beta, C_list[0], ldc, 0, 1, 0);
    if (st != PBX_OK || m == 0 || n == 0) return st;
  }
  const size_t eo = pbx_out_size(dtype);
  for (int x = 1; x < n_dst; ++x)
    PBX_CUDA_CHECK(h, cudaMemcpy2DAsync(C_list[x], (size_t)ldc * eo, C_list[0], (size_t)ldc * eo, (size_t)m * eo,
                                        (size_t)n, cudaMemcpyDefault, h->stream));
  return PBX_OK;
}

// ---- CUDA IPC: one process per GPU, so a peer's C is reached through an exported allocation handle ---------------
int pbx_ipc_export(pbx_handle_t h, const void* dptr, void* handle_out, int64_t* offset_out) {
  if (!h || !dptr || !handle_out || !offset_out) return PBX_ERR_INVALID_ARG;
  PBX_DEVICE_GUARD(h);
  // the handle names the whole allocation: find its base (allocator blocks are sub-ranges of one cudaMalloc)
  typedef CUresult (*range_fn_t)(CUdeviceptr*, size_t*, CUdeviceptr);
  void* f = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &f, cudaEnableDefault, &q) != cudaSuccess ||
      q != cudaDriverEntryPointSuccess || !f) {
    h->last_error = "cuMemGetAddressRange unavailable";
    return PBX_ERR_CUDA;
  }
  CUdeviceptr base = 0;
  size_t size = 0;
  if (reinterpret_cast<range_fn_t>(f)(&base, &size, (CUdeviceptr)dptr) != CUDA_SUCCESS) {
    h->last_error = "cuMemGetAddressRange failed";
    return PBX_ERR_CUDA;
  }
  PBX_CUDA_CHECK(h, cudaIpcGetMemHandle(reinterpret_cast<cudaIpcMemHandle_t*>(handle_out), (void*)base));
  *offset_out = (int64_t)((CUdeviceptr)dptr - base);
  return PBX_OK;
}

int pbx_ipc_import(pbx_handle_t h, const void* handle, int64_t offset, void** dptr_out) {
  if (!h || !handle || !dptr_out || offset < 0) return PBX_ERR_INVALID_ARG;
  PBX_DEVICE_GUARD(h);
  const std::string key(reinterpret_cast<const char*>(handle), sizeof(cudaIpcMemHandle_t));
  void* base = nullptr;
  for (auto& kv : h->ipc_open)
    if (kv.first == key) base = kv.second;
  if (!base) {   // an allocation can be opened once per process: keep it mapped until the handle is destroyed
    cudaIpcMemHandle_t hd;
    memcpy(&hd, handle, sizeof(hd));
    PBX_CUDA_CHECK(h, cudaIpcOpenMemHandle(&base, hd, cudaIpcMemLazyEnablePeerAccess));
    h->ipc_open.emplace_back(key, base);
  }
  *dptr_out = (char*)base + offset;
  return PBX_OK;
}

// ---- memory helpers ------------------------------------------------------------------
int pbx_malloc(pbx_handle_t h, void** dptr, int64_t bytes) {
  if (!h || !dptr || bytes < 0) return PBX_ERR_INVALID_ARG;
  PBX_DEVICE_GUARD(h);
  PBX_CUDA_CHECK(h, cudaMalloc(dptr, (size_t)(bytes > 0 ? bytes : 1)));
  return PBX_OK;
}
int pbx_free(pbx_handle_t h, void* dptr) {
  if (!h) return PBX_ERR_INVALID_ARG;
  PBX_DEVICE_GUARD(h);
  PBX_CUDA_CHECK(h, cudaStreamSynchronize(h->stream));
  PBX_CUDA_CHECK(h, cudaFree(dptr));
  return PBX_OK;
}
int pbx_copy_to_device(pbx_handle_t h, const void* src, void* dst, int64_t bytes) {
  if (!h || bytes < 0) return PBX_ERR_INVALID_ARG;
  PBX_DEVICE_GUARD(h);
  PBX_CUDA_CHECK(h, cudaMemcpyAsync(dst, src, (size_t)bytes, cudaMemcpyHostToDevice, h->stream));
  return PBX_OK;
}
int pbx_copy_to_host(pbx_handle_t h, const void* src, void* dst, int64_t bytes) {
  if (!h || bytes < 0) return PBX_ERR_INVALID_ARG;
  PBX_DEVICE_GUARD(h);
  PBX_CUDA_CHECK(h, cudaMemcpyAsync(dst, src, (size_t)bytes, cudaMemcpyDeviceToHost, h->stream));
  return PBX_OK;
}
// rows x cols window of a column-major matrix (elem_bytes per element) between host and device, each side with its own
// leading dimension: the sub-matrix form of copy_to_device / copy_to_host (an M-block of A or C is such a window)
int pbx_copy2d_to_device(pbx_handle_t h, const void* host_src, int64_t ld_src, void* dev_dst, int64_t ld_dst, int64_t rows,
                         int64_t cols, int elem_bytes) {
  if (!h || rows < 0 || cols < 0 || elem_bytes <= 0 || ld_src < rows || ld_dst < rows) return PBX_ERR_INVALID_ARG;
  if (rows == 0 || cols == 0) return PBX_OK;
  PBX_DEVICE_GUARD(h);
  PBX_CUDA_CHECK(h, cudaMemcpy2DAsync(dev_dst, (size_t)ld_dst * elem_bytes, host_src, (size_t)ld_src * elem_bytes,
                                      (size_t)rows * elem_bytes, (size_t)cols, cudaMemcpyHostToDevice, h->stream));
  return PBX_OK;
}
int pbx_copy2d_to_host(pbx_handle_t h, const void* dev_src, int64_t ld_src, void* host_dst, int64_t ld_dst, int64_t rows,
                       int64_t cols, int elem_bytes) {
  if (!h || rows < 0 || cols < 0 || elem_bytes <= 0 || ld_src < rows || ld_dst < rows) return PBX_ERR_INVALID_ARG;
  if (rows == 0 || cols == 0) return PBX_OK;
  PBX_DEVICE_GUARD(h);
  PBX_CUDA_CHECK(h, cudaMemcpy2DAsync(host_dst, (size_t)ld_dst * elem_bytes, dev_src, (size_t)ld_src * elem_bytes,
                                      (size_t)rows * elem_bytes, (size_t)cols, cudaMemcpyDeviceToHost, h->stream));
  return PBX_OK;
}
int pbx_fill_bytes(pbx_handle_t h, void* dst, int value, int64_t bytes) {
  if (!h || bytes < 0) return PBX_ERR_INVALID_ARG;
  PBX_DEVICE_GUARD(h);
  PBX_CUDA_CHECK(h, cudaMemsetAsync(dst, value, (size_t)bytes, h->stream));
  return PBX_OK;
}

int pbx_copy_device_to_device(pbx_handle_t h, const void* src, void* dst, int64_t bytes) {
  if (!h || bytes < 0) return PBX_ERR_INVALID_ARG;
  PBX_DEVICE_GUARD(h);
  PBX_CUDA_CHECK(h, cudaMemcpyAsync(dst, src, (size_t)bytes, cudaMemcpyDeviceToDevice, h->stream));
  return PBX_OK;
}

}  // extern "C"
template <typename T>
__global__ void pbx_fill_kernel(T* dst, T v, int64_t count) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x)
    dst[i] = v;
}
extern "C" {

int pbx_fill(pbx_handle_t h, void* dst, const void* value, int elem_bytes, int64_t count) {
  if (!h || !value || count < 0 || (elem_bytes != 2 && elem_bytes != 4 && elem_bytes != 8)) return PBX_ERR_INVALID_ARG;
  if (count == 0) return PBX_OK;
  PBX_DEVICE_GUARD(h);
  int64_t blocks = (count + 255) / 256;
  if (blocks > (int64_t)h->sm_count * 16) blocks = (int64_t)h->sm_count * 16;
  if (elem_bytes == 2) pbx_fill_kernel<uint16_t><<<(unsigned)blocks, 256, 0, h->stream>>>((uint16_t*)dst, *(const uint16_t*)value, count);
  else if (elem_bytes == 4) pbx_fill_kernel<uint32_t><<<(unsigned)blocks, 256, 0, h->stream>>>((uint32_t*)dst, *(const uint32_t*)value, count);
  else pbx_fill_kernel<uint64_t><<<(unsigned)blocks, 256, 0, h->stream>>>((uint64_t*)dst, *(const uint64_t*)value, count);
  h->launches++;
  PBX_CUDA_CHECK(h, cudaGetLastError());
  return PBX_OK;
}

// ---- events ---------------------------------------------------------------------------------
int pbx_event_create(pbx_handle_t h, void** ev) {
  if (!h || !ev) return PBX_ERR_INVALID_ARG;
  PBX_DEVICE_GUARD(h);
  cudaEvent_t e;
  PBX_CUDA_CHECK(h, cudaEventCreate(&e));
  *ev = (void*)e;
  return PBX_OK;
}
int pbx_event_record(pbx_handle_t h, void* ev) {
  if (!h || !ev) return PBX_ERR_INVALID_ARG;
  PBX_CUDA_CHECK(h, cudaEventRecord((cudaEvent_t)ev, h->stream));
  return PBX_OK;
}
int pbx_event_synchronize(pbx_handle_t h, void* ev) {
  if (!h || !ev) return PBX_ERR_INVALID_ARG;
  PBX_CUDA_CHECK(h, cudaEventSynchronize((cudaEvent_t)ev));
  return PBX_OK;
}
int pbx_event_elapsed_ms(pbx_handle_t h, void* a, void* b, float* ms) {
  if (!h || !a || !b || !ms) return PBX_ERR_INVALID_ARG;
  PBX_CUDA_CHECK(h, cudaEventElapsedTime(ms, (cudaEvent_t)a, (cudaEvent_t)b));
  return PBX_OK;
}
int pbx_event_destroy(pbx_handle_t h, void* ev) {
  if (!h || !ev) return PBX_ERR_INVALID_ARG;
  PBX_CUDA_CHECK(h, cudaEventDestroy((cudaEvent_t)ev));
  return PBX_OK;
}
int pbx_stream_wait_event(pbx_handle_t h, void* ev) {
  if (!h || !ev) return PBX_ERR_INVALID_ARG;
  PBX_CUDA_CHECK(h, cudaStreamWaitEvent(h->stream, (cudaEvent_t)ev, 0));
  return PBX_OK;
}
int pbx_device_name(pbx_handle_t h, char* buf, int len) {
  if (!h || !buf || len <= 0) return PBX_ERR_INVALID_ARG;
  cudaDeviceProp prop;
  PBX_CUDA_CHECK(h, cudaGetDeviceProperties(&prop, h->device));
  snprintf(buf, (size_t)len, "%s", prop.name);
  return PBX_OK;
}

}  // extern "C"
