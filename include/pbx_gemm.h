/*
 * pbx_gemm.h -- C-ABI of the B200-native GEMM path (libpbx_gemm.so).
 *
 * This is the drop-in boundary.  It replaces the SYCL launcher seam of the
 * reference,
 *     Gemm_Launcher<...>::_select_gemm(sb_handle, M, N, K, alpha, a_, lda,
 *                                      stridea, b_, ldb, strideb, beta, C,
 *                                      ldc, stridec, batch_size, deps)
 *     (reference include/interface/gemm_launcher.h:37-52,
 *      src/interface/gemm_launcher.hpp:39-64)
 * together with everything below it (backend heuristics
 * src/interface/blas3/backend/nvidia_gpu.hpp:40-260, the Gemm<> kernels in
 * src/operations/blas3/, SB_Handle::execute(Gemm) in
 * src/sb_handle/portblas_handle.hpp:277-403 and the alpha==0 scal path
 * src/interface/blas1_interface.hpp:438-510).
 *
 * Conventions (identical to the reference's interface, README.md:282-289):
 *   - column-major storage only; dimensions are post-transpose (op(A) is MxK,
 *     op(B) is KxN, C is MxN); ld >= stored rows.
 *   - trans chars are case-insensitive 'n' / 't' / 'c' ('c' == 't' for the real
 *     types handled here, src/interface/gemm_interface.hpp:141-151).
 *   - alpha/beta live on the HOST.  They are passed by pointer to a host value
 *     of the *scalar type* listed per entry point.
 *   - A, B, C are DEVICE pointers owned by the caller (cudaMalloc / torch
 *     allocations; the reference accepts only device USM, README.md:186-187).
 *   - batch_type: 0 = strided, 1 = interleaved
 *     (blas::gemm_batch_type_t, include/operations/blas3_trees.h:59).
 *   - every call is asynchronous on the handle's CUDA stream.
 *
 * Plain pointers and sizes only: no torch / C++ types cross this boundary.
 * There is no CPU fallback behind these symbols.
 */
#ifndef PBX_GEMM_H
#define PBX_GEMM_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PBX_VERSION 100

/* ---- status codes ------------------------------------------------------ */
typedef enum pbx_status {
  PBX_OK = 0,
  PBX_ERR_INVALID_TRANSA = 1, /* reference throws invalid_argument("invalid _TransA") gemm_interface.hpp:144 */
  PBX_ERR_INVALID_TRANSB = 2, /* "invalid _TransB"  gemm_interface.hpp:146 */
  PBX_ERR_INVALID_STRIDEC = 3, /* "invalid _stridec" gemm_interface.hpp:159 */
  PBX_ERR_INVALID_STRIDEA = 4, /* "invalid _stridea" gemm_interface.hpp:161 */
  PBX_ERR_INVALID_STRIDEB = 5, /* "invalid _strideb" gemm_interface.hpp:163 */
  PBX_ERR_INVALID_ARG = 6,     /* null handle, unknown dtype/batch_type, negative dims */
  PBX_ERR_CUDA = 7,            /* a CUDA runtime / driver call failed; see pbx_last_error() */
  PBX_ERR_NO_DEVICE = 8,       /* no sm_100 device: this library has no fallback */
  PBX_ERR_WORKSPACE = 9,       /* a pooled temporary (split-K workspace, packed operand) could not be allocated */
  PBX_ERR_INVALID_UPLO = 10,   /* _symm: "invalid _uplo"  src/interface/symm_interface.hpp:51-53 */
  PBX_ERR_INVALID_SIDE = 11,   /* _symm: "invalid _side"  src/interface/symm_interface.hpp:70-72 */
  PBX_ERR_TRSM_SIZE = 12,      /* _trsm: "invalid matrix size argument"  src/interface/trsm_interface.hpp:112-114 */
  PBX_ERR_TRSM_SIDE = 13,      /* _trsm: "invalid Side argument"       trsm_interface.hpp:121-122 */
  PBX_ERR_TRSM_UPLO = 14,      /* _trsm: "invalid Triangle argument"   trsm_interface.hpp:123-124 */
  PBX_ERR_TRSM_TRANS = 15,     /* _trsm: "invalid Transpose argument"  trsm_interface.hpp:125-126 */
  PBX_ERR_TRSM_DIAG = 16       /* _trsm: "invalid Diagonal argument"   trsm_interface.hpp:127-128 */
} pbx_status_t;

/* ---- element types (in -> out) ----------------------------------------- *
 * Mirrors the (DATA_TYPE_IN, DATA_TYPE_OUT) matrix the reference instantiates
 * (src/interface/blas3/gemm.cpp.in:33-137) plus bf16 storage (BASELINE cfg 4). */
typedef enum pbx_dtype {
  PBX_F32 = 0,      /* float  -> float   scalar: float   (3xTF32 on tcgen05, fp32 accumulate) */
  PBX_F64 = 1,      /* double -> double  scalar: double  (DMMA mma.sync f64) */
  PBX_F16 = 2,      /* half   -> half    scalar: float   (tcgen05 kind::f16, fp32 accumulate) */
  PBX_F16_F32 = 3,  /* half   -> float   scalar: float   (mixed, gemm.cpp.in:36-39) */
  PBX_BF16 = 4,     /* bf16   -> bf16    scalar: float   (tcgen05 kind::f16) */
  PBX_BF16_F32 = 5  /* bf16   -> float   scalar: float */
} pbx_dtype_t;

/* ---- kernel families (for pbx_set_forced_kernel / pbx_last_kernel) ------ */
typedef enum pbx_kernel {
  PBX_KERNEL_AUTO = 0,
  PBX_KERNEL_SIMT = 1,        /* shared-memory tiled CUDA-core kernel: any alignment, any ld */
  PBX_KERNEL_TCGEN05 = 2,     /* TMA + tcgen05/TMEM (f16/bf16 direct, f32 as 3xTF32) */
  PBX_KERNEL_DMMA = 3,        /* fp64 tensor path (mma.sync m8n8k4 f64) */
  PBX_KERNEL_INTERLEAVED = 4, /* interleaved-batch kernel */
  PBX_KERNEL_SCAL = 5,        /* alpha == 0 : C = beta*C */
  PBX_KERNEL_NONE = 6         /* nothing launched (m==0 || n==0 || batch==0, or beta==1 scal) */
} pbx_kernel_t;

typedef struct pbx_handle_s* pbx_handle_t;

/* ---- handle ------------------------------------------------------------- *
 * Replaces blas::SB_Handle's device-facing half
 * (include/sb_handle/portblas_handle.h:46-200): one handle == one device +
 * one stream, caches the SM count (get_num_compute_units) and owns the
 * split-K workspace pool (Temp_Mem_Pool, include/sb_handle/temp_memory_pool.h:33-114).
 * Not thread-safe, like the reference (one handle per thread), and tied to ONE in-order stream at a time: the
 * persistent kernel's tile counters and the pooled temporaries live in the handle, so two calls of the same handle
 * must not run concurrently (change streams with pbx_set_stream only after pbx_synchronize).                    */
int pbx_create(pbx_handle_t* out, int device_ordinal, void* cuda_stream /* cudaStream_t or NULL */);
int pbx_destroy(pbx_handle_t h);
int pbx_set_stream(pbx_handle_t h, void* cuda_stream);
void* pbx_get_stream(pbx_handle_t h);
int pbx_get_num_compute_units(pbx_handle_t h);     /* SM count (148 on B200) */
int pbx_get_device(pbx_handle_t h);
int pbx_synchronize(pbx_handle_t h);               /* SB_Handle::wait() */
const char* pbx_last_error(pbx_handle_t h);        /* text of the last failure (may be "") */
const char* pbx_status_string(int status);         /* the reference's exception text for a status */

/* Testing / tuning hooks.  The PBX_* environment switches (DESIGN.md lists them) are read by pbx_create and again by
 * pbx_reload_env; a GEMM call itself reads only SB_ENABLE_JOINT_MATRIX, the variable the reference reads per call. */
int pbx_reload_env(pbx_handle_t h);
int pbx_set_forced_kernel(pbx_handle_t h, int kernel /* pbx_kernel_t */);
int pbx_set_split_k(pbx_handle_t h, int slices /* 0 = auto, 1 = never, >1 = force */);
int pbx_last_kernel(pbx_handle_t h);               /* pbx_kernel_t used by the last call */
int pbx_last_split_k(pbx_handle_t h);              /* K slices used by the last call */
int pbx_last_repack(pbx_handle_t h);               /* bit 0 / 1: A / B was re-laid out to a 16-byte-legal copy */
int pbx_last_presplit(pbx_handle_t h);             /* fp32 mode of the last call: 0 in-kernel 3xTF32 split, 1 pre-split tf32 lo halves, 2 single tf32 (SB_ENABLE_JOINT_MATRIX=1), 3 tf32 + 2 x bf16 (pre-split bf16 copies) */
int64_t pbx_launch_count(pbx_handle_t h);          /* kernels launched through this handle so far */
int64_t pbx_workspace_bytes(pbx_handle_t h);       /* current size of the pooled workspace */
/* The tensor-core tile plan for a shape as a pure function (no device, no handle): CTA group (1 or 2), tile width,
 * K slices, skinny-M operand swap.  What replaces the reference's shape heuristics (nvidia_gpu.hpp:116-171).   */
int pbx_plan_query(int sm_count, int dtype, int64_t m, int64_t n, int64_t k, int64_t batch, int* cta_group,
                   int* tile_n, int* k_slices, int* swapped);

/* ---- the GEMM entry point ------------------------------------------------
 *  C_b <- alpha * op(A_b) * op(B_b) + beta * C_b      b = 0 .. batch-1
 *
 *  strided     : X_b = X + b*strideX (elements); strideA/strideB may be 0
 *                (broadcast); for batch > 1 strideC >= ldc*N is required.
 *  interleaved : element (r, c, b) lives at X[(c*ldX + r)*batch + b]
 *                (reference src/operations/blas3/gemm_interleaved.hpp:265-271);
 *                strides are ignored.
 *
 *  Front-end rules restated from src/interface/gemm_interface.hpp:105-185:
 *   - alpha == 0 is tested FIRST (before trans validation): C <- beta*C over
 *     the MxN window(s); beta == 1 is a no-op, beta == 0 stores exact zeros.
 *   - beta == 0 never reads C (NaN-safe, gemm_interface.hpp:85-100).
 *   - m == 0 || n == 0 || batch == 0: nothing happens; k == 0: C <- beta*C.
 *  Returns a pbx_status_t.                                                    */
int pbx_gemm(pbx_handle_t h, int dtype /* pbx_dtype_t */, char transa, char transb,
             int64_t m, int64_t n, int64_t k,
             const void* alpha,
             const void* A, int64_t lda, int64_t stridea,
             const void* B, int64_t ldb, int64_t strideb,
             const void* beta,
             void* C, int64_t ldc, int64_t stridec,
             int64_t batch, int batch_type);

/* Per-type spellings (what a per-dtype explicit instantiation of
 * blas::internal::_gemm / _gemm_batched / _gemm_strided_batched binds to). */
int pbx_sgemm(pbx_handle_t h, char transa, char transb, int64_t m, int64_t n, int64_t k,
              const float* alpha, const float* A, int64_t lda, int64_t stridea,
              const float* B, int64_t ldb, int64_t strideb, const float* beta,
              float* C, int64_t ldc, int64_t stridec, int64_t batch, int batch_type);
int pbx_dgemm(pbx_handle_t h, char transa, char transb, int64_t m, int64_t n, int64_t k,
              const double* alpha, const double* A, int64_t lda, int64_t stridea,
              const double* B, int64_t ldb, int64_t strideb, const double* beta,
              double* C, int64_t ldc, int64_t stridec, int64_t batch, int batch_type);
int pbx_hgemm(pbx_handle_t h, char transa, char transb, int64_t m, int64_t n, int64_t k,
              const float* alpha, const void* A, int64_t lda, int64_t stridea,
              const void* B, int64_t ldb, int64_t strideb, const float* beta,
              void* C, int64_t ldc, int64_t stridec, int64_t batch, int batch_type);
int pbx_hsgemm(pbx_handle_t h, char transa, char transb, int64_t m, int64_t n, int64_t k,
               const float* alpha, const void* A, int64_t lda, int64_t stridea,
               const void* B, int64_t ldb, int64_t strideb, const float* beta,
               float* C, int64_t ldc, int64_t stridec, int64_t batch, int batch_type);
int pbx_bf16gemm(pbx_handle_t h, char transa, char transb, int64_t m, int64_t n, int64_t k,
                 const float* alpha, const void* A, int64_t lda, int64_t stridea,
                 const void* B, int64_t ldb, int64_t strideb, const float* beta,
                 void* C, int64_t ldc, int64_t stridec, int64_t batch, int batch_type);

/* C <- beta*C on an MxN column-major window (x batch, strided).  Replaces
 * blas::internal::_scal_matrix (src/interface/blas1_interface.hpp:468-510).
 * beta == 1: no launch.  beta == 0: stores zeros without reading C.          */
int pbx_scal_matrix(pbx_handle_t h, int dtype, int64_t m, int64_t n, const void* beta,
                    void* C, int64_t ldc, int64_t stridec, int64_t batch);

/* ---- routines built on the GEMM path -------------------------------
 *
 * pbx_symm:  C <- alpha*A*B + beta*C (side 'l', A is MxM) or alpha*B*A + beta*C (side 'r', A is NxN),
 *   A symmetric with only its `uplo` triangle referenced.  Replaces blas::internal::_symm
 *   (src/interface/symm_interface.hpp:35-75), which runs the GEMM kernels with a mirroring operand
 *   loader (gemm_local.hpp:813-873).  Here one HBM-bound pass mirrors the triangle into a pooled full
 *   matrix, then the ordinary tensor-core GEMM runs.  dtype: PBX_F32, PBX_F64, PBX_F16 or PBX_BF16.
 *   Checks, in the reference's order: uplo, side; then the GEMM front end (alpha == 0 shortcut ...).    */
int pbx_symm(pbx_handle_t h, int dtype, char side, char uplo, int64_t m, int64_t n, const void* alpha,
             const void* A, int64_t lda, const void* B, int64_t ldb, const void* beta, void* C, int64_t ldc);

/* pbx_trsm:  solves op(A)*X = alpha*B (side 'l') or X*op(A) = alpha*B (side 'r') in place of B (MxN);
 *   A triangular (`uplo`), `diag` 'u' = unit diagonal assumed, trans 'n' / 't' only (the reference rejects
 *   'c', trsm_interface.hpp:125).  Replaces blas::internal::_trsm (src/interface/trsm_interface.hpp:105-387):
 *   the same inverse-of-diagonal-blocks + GEMM scheme (DiagonalBlocksInverter, src/operations/blas3/trsm.hpp),
 *   but with 64/128-wide diagonal blocks and a recursive split so that the trailing updates are large
 *   tensor-core GEMMs instead of rank-16 ones.  The other triangle of A is never read (it may hold NaN).
 *   dtype: PBX_F32 or PBX_F64.                                                                           */
int pbx_trsm(pbx_handle_t h, int dtype, char side, char uplo, char trans, char diag, int64_t m, int64_t n,
             const void* alpha, const void* A, int64_t lda, void* B, int64_t ldb);

/* pbx_cgemm / pbx_zgemm:  complex GEMM (BLAS_ENABLE_COMPLEX builds of the reference: gemm.cpp.in,
 *   backend/default.hpp:202-246, nvidia_gpu.hpp:237-260), interleaved (re, im) storage, strided batches only
 *   (the reference hard-codes gemm_batch_type_t::strided for complex).  alpha / beta point to host
 *   {re, im} pairs.  Computed as ONE real GEMM of twice the size on the tensor-core path:
 *   [Cr; Ci] = [Ar -Ai; Ai Ar] * [Br; Bi], framed by an HBM-bound planar split and a combine pass that
 *   applies alpha and beta.  Same front-end rules as pbx_gemm.  'c' is treated like 't' (NO conjugation),
 *   exactly as the reference does (gemm_interface.hpp:150-151: _TrA = _TransA != 'n');
 *   pbx_set_conj_transpose(h, 1) switches to BLAS-standard conjugate transposes.                          */
int pbx_cgemm(pbx_handle_t h, char transa, char transb, int64_t m, int64_t n, int64_t k, const float* alpha,
              const void* A, int64_t lda, int64_t stridea, const void* B, int64_t ldb, int64_t strideb,
              const float* beta, void* C, int64_t ldc, int64_t stridec, int64_t batch);
int pbx_zgemm(pbx_handle_t h, char transa, char transb, int64_t m, int64_t n, int64_t k, const double* alpha,
              const void* A, int64_t lda, int64_t stridea, const void* B, int64_t ldb, int64_t strideb,
              const double* beta, void* C, int64_t ldc, int64_t stridec, int64_t batch);
int pbx_set_conj_transpose(pbx_handle_t h, int enable);

/* ---- multi-GPU: the C gather fused into the GEMM's stores ---------------------------------------------------
 * One process per GPU.  A rank exports its full C with pbx_ipc_export (64-byte CUDA IPC handle
 * of the underlying allocation + byte offset), the peers map it with pbx_ipc_import, and every rank then runs its
 * M-block with pbx_gemm_multicast: C_list[0] is the local copy (read when beta != 0), C_list[1..n_dst-1] the same
 * row block inside the peers' C.  The tensor-core epilogue stores each finished tile to all copies over NVLink
 * while the next tile is computed; no staging buffer, no collective call.  batch == 1, n_dst <= 8.               */
int pbx_gemm_multicast(pbx_handle_t h, int dtype, char transa, char transb, int64_t m, int64_t n, int64_t k,
                       const void* alpha, const void* A, int64_t lda, const void* B, int64_t ldb, const void* beta,
                       void* const* C_list, int n_dst, int64_t ldc);
int pbx_ipc_export(pbx_handle_t h, const void* dptr, void* handle_out /* 64 bytes */, int64_t* offset_out);
int pbx_ipc_import(pbx_handle_t h, const void* handle /* 64 bytes */, int64_t offset, void** dptr_out);

/* ---- multi-GPU from one host thread: a group of devices with peer access ----------------------------------------
 * The reference has no multi-device code (one SB_Handle == one sycl::queue, include/sb_handle/portblas_handle.h:51-60);
 * these entry points are the north star's "large square GEMMs are partitioned by M-blocks and strided-batched GEMMs by
 * batch across the 8 GPUs of one box" for a C / C++ caller (samples/gemm_multi_b200.cpp).  A group owns one handle and
 * one stream per device; every call is asynchronous on those streams, pbx_multi_synchronize waits for all of them.
 * pbx_shard_range is THE partition (pure function, also bound by portblas_b200/sharding.py): part `index` of `parts`
 * gets [start, start+count) of `total` units, shares are multiples of `align` (M-blocks: 256 rows; batches: 1).     */
typedef struct pbx_multi_s* pbx_multi_t;
int pbx_shard_range(int64_t total, int parts, int index, int64_t align, int64_t* start, int64_t* count);
int pbx_multi_create(pbx_multi_t* out, int n_dev, const int* device_ordinals /* NULL = 0..n_dev-1 */);
int pbx_multi_destroy(pbx_multi_t mh);
int pbx_multi_device_count(pbx_multi_t mh);
pbx_handle_t pbx_multi_handle(pbx_multi_t mh, int i);   /* device i's handle (its stream: pbx_get_stream) */
int pbx_multi_synchronize(pbx_multi_t mh);
const char* pbx_multi_last_error(pbx_multi_t mh);
/* C <- alpha*op(A)*op(B) + beta*C, M-block sharded, operands resident: A_blocks[g] = device g's rows
 * pbx_shard_range(m, G, g, 256) of op(A) (leading dimension lda), B_full[g] = all of B on device g, C_full[g] = device
 * g's m x n C.  gather == 0: device g writes its rows of its own C.  gather != 0: the GEMM epilogue of every device
 * stores its tiles into ALL devices' C over NVLink, so each device ends with the whole product (no collective).  */
int pbx_gemm_sharded(pbx_multi_t mh, int dtype, char transa, char transb, int64_t m, int64_t n, int64_t k,
                     const void* alpha, const void* const* A_blocks, int64_t lda, const void* const* B_full, int64_t ldb,
                     const void* beta, void* const* C_full, int64_t ldc, int gather);
/* strided batches cut into batch ranges pbx_shard_range(batch, G, g, 1); X_shards[g] = first entry device g owns */
int pbx_gemm_strided_batched_sharded(pbx_multi_t mh, int dtype, char transa, char transb, int64_t m, int64_t n, int64_t k,
                                     const void* alpha, const void* const* A_shards, int64_t lda, int64_t stridea,
                                     const void* const* B_shards, int64_t ldb, int64_t strideb, const void* beta,
                                     void* const* C_shards, int64_t ldc, int64_t stridec, int64_t batch);
/* blas::_gemm on HOST operands over the whole group (synchronous): every device uploads its M-block of A and 1/G of
 * B, the B panels are exchanged over NVLink, every device computes and downloads its block of C.                 */
int pbx_gemm_sharded_host(pbx_multi_t mh, int dtype, char transa, char transb, int64_t m, int64_t n, int64_t k,
                          const void* alpha, const void* A_host, int64_t lda, const void* B_host, int64_t ldb,
                          const void* beta, void* C_host, int64_t ldc);

/* ---- host-buffer convenience path (used for the end-to-end metric) -------
 * Same semantics as pbx_gemm, but A, B, C are HOST pointers (ideally pinned):
 * stages H2D on the handle's stream, runs the GEMM, copies C back and
 * synchronises.  Device staging buffers are cached in the handle.
 * This is what blas::helper::copy_to_device + _gemm + copy_to_host amount to
 * (reference include/portblas_helper.h:139-191, samples/gemm.cpp:50-63).      */
int pbx_gemm_host(pbx_handle_t h, int dtype, char transa, char transb,
                  int64_t m, int64_t n, int64_t k, const void* alpha,
                  const void* A_host, int64_t lda, int64_t stridea,
                  const void* B_host, int64_t ldb, int64_t strideb,
                  const void* beta, void* C_host, int64_t ldc, int64_t stridec,
                  int64_t batch, int batch_type);

/* ---- device memory helpers (blas::helper::allocate / copy_to_device / ...,
 * include/portblas_helper.h:54-81,139-219) -------------------------------- */
int pbx_malloc(pbx_handle_t h, void** dptr, int64_t bytes);
int pbx_free(pbx_handle_t h, void* dptr);
int pbx_copy_to_device(pbx_handle_t h, const void* host_src, void* dev_dst, int64_t bytes);
int pbx_copy_to_host(pbx_handle_t h, const void* dev_src, void* host_dst, int64_t bytes);
/* rows x cols window of a column-major matrix, each side with its own leading dimension (an M-block of A or C) */
int pbx_copy2d_to_device(pbx_handle_t h, const void* host_src, int64_t ld_src, void* dev_dst, int64_t ld_dst, int64_t rows,
                         int64_t cols, int elem_bytes);
int pbx_copy2d_to_host(pbx_handle_t h, const void* dev_src, int64_t ld_src, void* host_dst, int64_t ld_dst, int64_t rows,
                       int64_t cols, int elem_bytes);
int pbx_fill_bytes(pbx_handle_t h, void* dev_dst, int value, int64_t bytes);

int pbx_copy_device_to_device(pbx_handle_t h, const void* dev_src, void* dev_dst, int64_t bytes);
/* typed fill: `count` elements of `elem_bytes` (2, 4 or 8) bytes each set to *value */
int pbx_fill(pbx_handle_t h, void* dev_dst, const void* value, int elem_bytes, int64_t count);

/* ---- events (what the reference returns from every call: sb_handle_t::event_t =
 * std::vector<sycl::event>, include/sb_handle/portblas_handle.h:48; profiling as in
 * benchmark/portblas/utils.hpp:52-63) ---------------------------------------------------- */
int pbx_event_create(pbx_handle_t h, void** event_out);
int pbx_event_record(pbx_handle_t h, void* event);          /* on the handle's stream */
int pbx_event_synchronize(pbx_handle_t h, void* event);     /* sycl::event::wait() */
int pbx_event_elapsed_ms(pbx_handle_t h, void* start, void* end, float* ms_out);
int pbx_event_destroy(pbx_handle_t h, void* event);
int pbx_stream_wait_event(pbx_handle_t h, void* event);     /* handler::depends_on() */
int pbx_device_name(pbx_handle_t h, char* buf, int len);

#ifdef __cplusplus
}
#endif
#endif /* PBX_GEMM_H */
