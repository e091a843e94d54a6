// pbx_internal.cuh -- shared host-side types of libpbx_gemm (not part of the ABI).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/pbx_gemm.h"

// One GEMM call after front-end normalisation: trans chars resolved, scalars
// widened.  Layout convention: column-major, element (r,c) of a stored matrix
// X at X[r + c*ldx].  op(A) is m x k, op(B) is k x n  (reference
// src/interface/gemm_launcher.hpp:54-56 builds its views the same way).
struct PbxGemmCall {
  int dtype;
  bool ta, tb;
  int64_t m, n, k;
  double alpha, beta;  // exact for float and double scalars
  const void* A;
  const void* B;
  void* C;
  int64_t lda, ldb, ldc;
  int64_t sa, sb, sc;  // batch strides in elements (strided)
  int64_t batch;
  // multicast GEMM (pbx_gemm_multicast): further copies of C that receive every tile (peer GPUs' memory)
  int n_extra = 0;
  void* c_extra[7] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
};

// cuTensorMapEncodeTiled is a pure function of its arguments and costs a microsecond or two per map; a GEMM call
// needs two to thirteen maps, and callers repeat calls on the same buffers.  The handle keeps the last encodings.
struct PbxTmapCacheEntry {
  uint64_t key[10];
  CUtensorMap map;
  bool valid = false;
};

// Testing / tuning switches (PBX_* environment variables).  They are read when the handle is created and again by
// pbx_reload_env(): a GEMM call itself reads one variable only, SB_ENABLE_JOINT_MATRIX -- the one the reference reads
// per call (src/interface/blas3/backend/nvidia_gpu.hpp:68-69).
struct PbxKnobs {
  int tc_swap = 1;            // PBX_TC_SWAP=0: no skinny-M operand swap
  int tc_cg = 0, tc_bn = 0;   // PBX_TC_CONFIG="cg,bn": forced tile configuration
  int plan_model = 1;         // PBX_PLAN_MODEL=0: round 1's threshold planner
  int tf32_presplit = -1;     // PBX_TF32_PRESPLIT=0/1: never / always a pre-pass for the fp32 lo halves (-1: by shape)
  int f32_split16 = 1;        // PBX_F32_SPLIT16=0: the 3xTF32 forms instead of tf32 + 2 x bf16
  int tf32_chunk_kb = 0;      // PBX_TF32_CHUNK_KB: K blocks per tensor-core accumulation chain (0: 16)
  int tf32_raw_hi = 1;        // PBX_TF32_RAW_HI=0: rounded tf32 hi operand instead of the raw fp32 tile
  int multicast_pace = 1;     // PBX_MULTICAST_PACE=0: the pusher sends a tile's peer copies at once
  int multicast_push = 1;     // PBX_MULTICAST_PUSH=0: peer copies stored by the epilogue warps (round 1)
  unsigned wait_hint_ns = 0;  // PBX_WAIT_HINT_NS: suspend-time hint of the long mbarrier waits (measured: no gain)
  int tma_store = 1;          // PBX_TMA_STORE=0: 16-bit C by direct stores
  int ilv_via_strided = -1;   // PBX_ILV_VIA_STRIDED=0/1: never / always re-lay interleaved batches out (-1: by shape)
  int group_m = 0;            // PBX_GROUP_M: M tiles per raster group (0: 8 for CTA pairs, 16 for single CTAs)
};

struct pbx_handle_s {
  int device = 0;
  cudaStream_t stream = nullptr;
  int sm_count = 0;
  int cc_major = 0, cc_minor = 0;
  int forced_kernel = PBX_KERNEL_AUTO;
  int forced_split_k = 0;
  int last_kernel = PBX_KERNEL_NONE;
  int last_split_k = 1;
  int last_repack = 0;
  int64_t launches = 0;
  std::string last_error;
  // split-K workspace pool (stream ordered; grows monotonically)
  void* ws = nullptr;
  int64_t ws_bytes = 0;
  // aligned copies of TMA-illegal operands (pbx_api.cu: repack_for_tma); grow monotonically
  void* pack[2] = {nullptr, nullptr};
  int64_t pack_bytes[2] = {0, 0};
  // fp32 pre-split: lo halves (tf32) of A and B computed once per call (gemm_tcgen05.cu); grow monotonically
  void* lo[2] = {nullptr, nullptr};
  int64_t lo_bytes[2] = {0, 0};
  int last_presplit = 0;
  // pooled temporaries of the GEMM-built routines (blas3_ext.cu: symmetrised A, planar complex operands,
  // inverted diagonal blocks / X of trsm); grow monotonically, stream ordered
  void* aux[4] = {nullptr, nullptr, nullptr, nullptr};
  int64_t aux_bytes[4] = {0, 0, 0, 0};
  int conj_transpose = 0;   // complex GEMM: 0 = 'c' behaves as 't' (the reference), 1 = BLAS conjugate-transpose
  // peer allocations opened through CUDA IPC (pbx_ipc_import): 64-byte handle -> mapped base pointer
  std::vector<std::pair<std::string, void*>> ipc_open;
  // dynamic tile schedule of the persistent tcgen05 kernel: {tiles handed out, groups done}, zero between launches
  // (the kernel re-arms them itself); PBX_DYNAMIC_SCHED=0 / PBX_PDL=0 switch the two launch features off (testing)
  unsigned int* tile_sched = nullptr;
  int dynamic_sched = 1;
  int pdl = 1;
  int pdl_reduce = 1;   // PBX_PDL_REDUCE=0: the split-K reduce kernel alone is launched without the attribute
  int last_grid_ctas = 1 << 30;   // CTAs of the last tcgen05 launch (the reduce kernel rides along only when SMs are free)
  PbxTmapCacheEntry tmap_cache[64];
  PbxKnobs knobs;
  // staging buffers for pbx_gemm_host
  void* stage[3] = {nullptr, nullptr, nullptr};
  int64_t stage_bytes[3] = {0, 0, 0};
  // copy-in / copy-out streams and events of the pipelined host path (pbx_host.cu)
  cudaStream_t s_in = nullptr, s_out = nullptr;
  std::vector<cudaEvent_t> events;
};

#define PBX_CUDA_CHECK(h, expr)                                                          \
  do {                                                                                   \
    cudaError_t _e = (expr);                                                             \
    if (_e != cudaSuccess) {                                                             \
      (h)->last_error = std::string(#expr) + ": " + cudaGetErrorString(_e);              \
      fprintf(stderr, "[pbx_gemm] CUDA error %s at %s:%d\n", (h)->last_error.c_str(),   \
              __FILE__, __LINE__);                                                       \
      return PBX_ERR_CUDA;                                                               \
    }                                                                                    \
  } while (0)

// Every entry point works on the handle's device and leaves the caller's current device as it found it: a
// single-process multi-GPU program (torch, or several SB_Handles in one thread) must not have its thread switched
// to another GPU by a library call.
struct PbxDeviceGuard {
  int prev = -1;
  bool changed = false, good = true;
  explicit PbxDeviceGuard(pbx_handle_t h) { enter(h->device); }
  explicit PbxDeviceGuard(int device) { enter(device); }
  ~PbxDeviceGuard() { if (changed) cudaSetDevice(prev); }
  PbxDeviceGuard(const PbxDeviceGuard&) = delete;
  PbxDeviceGuard& operator=(const PbxDeviceGuard&) = delete;
  bool ok() const { return good; }
 private:
  void enter(int device) {
    if (cudaGetDevice(&prev) != cudaSuccess) { good = false; return; }
    if (prev != device) { good = (cudaSetDevice(device) == cudaSuccess); changed = good; }
  }
};
#define PBX_DEVICE_GUARD(h)                                                              \
  PbxDeviceGuard _pbx_device_guard(h);                                                   \
  if (!_pbx_device_guard.ok()) { (h)->last_error = "cudaSetDevice failed"; return PBX_ERR_CUDA; }

static inline size_t pbx_in_size(int dtype) {
  switch (dtype) {
    case PBX_F32: return 4;
    case PBX_F64: return 8;
    default: return 2;
  }
}
static inline size_t pbx_out_size(int dtype) {
  switch (dtype) {
    case PBX_F32: case PBX_F16_F32: case PBX_BF16_F32: return 4;
    case PBX_F64: return 8;
    default: return 2;
  }
}

int pbx_ensure_workspace(pbx_handle_t h, int64_t bytes);
int pbx_ensure_aux(pbx_handle_t h, int i, int64_t bytes);
int pbx_ensure_lo(pbx_handle_t h, int i, int64_t bytes);
// lo = rn_tf32(x - trunc_tf32(x)) over a rows x cols column-major window (x batch), same ld / stride as the source
int pbx_launch_tf32_lo(pbx_handle_t h, const float* src, float* dst, int64_t rows, int64_t cols, int64_t ld,
                       int64_t stride, int64_t batch);
// EXPERIMENTAL (PBX_F32_SPLIT16=1): bf16(a) and bf16(a - trunc_tf32(a)) copies of an fp32 operand with their own ld / stride
int pbx_launch_split16(pbx_handle_t h, const float* src, void* hi, void* lo, int64_t rows, int64_t cols, int64_t ld,
                       int64_t stride, int64_t ld16, int64_t st16, int64_t batch);

// ---- kernel families (each returns a pbx_status_t) -----------------------
// slices > 1: the kernel writes raw fp32/fp64 partial sums to h->ws laid out
// [batch][slice][n][m] (compact, m contiguous) and pbx_launch_splitk_reduce
// applies alpha/beta.
int pbx_launch_simt(pbx_handle_t h, const PbxGemmCall& c, int slices);
int pbx_launch_interleaved(pbx_handle_t h, const PbxGemmCall& c);
int pbx_launch_scal(pbx_handle_t h, int dtype, int64_t m, int64_t n, double beta, void* C,
                    int64_t ldc, int64_t stridec, int64_t batch, int interleaved);
int pbx_launch_splitk_reduce(pbx_handle_t h, const PbxGemmCall& c, int slices);
bool pbx_tcgen05_eligible(pbx_handle_t h, const PbxGemmCall& c);
bool pbx_tcgen05_shape_ok(pbx_handle_t h, const PbxGemmCall& c);   // everything except operand alignment
bool pbx_tma_operand_ok(int dtype, const void* p, int64_t ld, int64_t stride);
// copy a rows x cols column-major window (x batch) to a new leading dimension / batch stride
int pbx_launch_repack(pbx_handle_t h, int elem_bytes, const void* src, void* dst, int64_t rows, int64_t cols,
                      int64_t ld_src, int64_t ld_dst, int64_t stride_src, int64_t stride_dst, int64_t batch);
// interleaved <-> strided re-layout of `batch` column-major rows x cols matrices (gemm_simt.cu)
int pbx_launch_ilv_relayout(pbx_handle_t h, int elem_bytes, const void* src, void* dst, int64_t rows, int64_t cols,
                            int64_t ld_i, int64_t ld_s, int64_t stride, int64_t batch, bool to_strided);
int pbx_tcgen05_slices(pbx_handle_t h, const PbxGemmCall& c);  // K slices the tcgen05 plan wants
int pbx_launch_tcgen05(pbx_handle_t h, const PbxGemmCall& c, int slices);
int pbx_launch_dmma(pbx_handle_t h, const PbxGemmCall& c, int slices);
