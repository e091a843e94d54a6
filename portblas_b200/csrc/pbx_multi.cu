// pbx_multi.cu -- one GEMM over several B200s of one box, from ONE host thread (C-ABI: pbx_multi_* / pbx_gemm_sharded*).
//
// The reference is single-device: one SB_Handle wraps one sycl::queue (include/sb_handle/portblas_handle.h:51-60) and
// blas::_gemm (include/interface/blas3_interface.h:88-95) runs on that queue's device.  This file is the multi-GPU
// counterpart the north star asks for, behind the same argument conventions:
//
//   * large GEMMs are cut into M-blocks: device g owns rows [row0_g, row0_g + rows_g) of op(A) and of C, B is replicated
//     (SURVEY.md section 8e).  In column-major storage a row block of C is a pointer offset with the ORIGINAL ldc, so every
//     device runs an ordinary GEMM on its own stream; with `gather` the tensor-core epilogue of each device stores its
//     finished tiles into the same rows of EVERY device's C over NVLink (pbx_gemm_multicast: peer pointers are plain
//     device pointers here, peer access is enabled between all devices of the group), so that all devices end with
//     the whole product and no collective runs;
//   * strided batches are cut into batch ranges, one pbx_gemm per device;
//   * pbx_gemm_sharded_host takes HOST operands: every device uploads its M-block of A and 1/G of B over its own PCIe
//     link, the B panels are exchanged device to device over NVLink (so B crosses PCIe once, not G times), each
//     device computes its block and downloads it.
//
// The partition itself (pbx_shard_range) is a pure function and is what portblas_b200/sharding.py binds, so the
// one-process-per-GPU path (torch.distributed, bench.py) and this one cut problems identically.
#include <ctype.h>
#include <stdio.h>

#include <vector>

#include "pbx_internal.cuh"

struct pbx_multi_s {
  std::vector<pbx_handle_t> h;
  std::vector<cudaStream_t> stream;
  std::vector<cudaEvent_t> ev;        // one per device: "my uploads / my GEMM are done"
  // pooled device operands of pbx_gemm_sharded_host: A block, full B, C block per device
  std::vector<void*> dA, dB, dC;
  std::vector<int64_t> dA_bytes, dB_bytes, dC_bytes;
  std::string last_error;
};

namespace {

int ensure_buf(pbx_multi_t mh, int g, std::vector<void*>& buf, std::vector<int64_t>& cap, int64_t bytes) {
  if (bytes <= cap[g]) return PBX_OK;
  PbxDeviceGuard guard(mh->h[g]);
  if (!guard.ok()) return PBX_ERR_CUDA;
  if (buf[g]) {
    if (cudaStreamSynchronize(mh->stream[g]) != cudaSuccess || cudaFree(buf[g]) != cudaSuccess) return PBX_ERR_WORKSPACE;
    buf[g] = nullptr; cap[g] = 0;
  }
  if (cudaMalloc(&buf[g], (size_t)(bytes > 0 ? bytes : 1)) != cudaSuccess) { cudaGetLastError(); return PBX_ERR_WORKSPACE; }
  cap[g] = bytes;
  return PBX_OK;
}

cudaError_t copy2d(void* dst, int64_t ld_dst, const void* src, int64_t ld_src, int64_t rows, int64_t cols, int64_t es,
                   cudaMemcpyKind kind, cudaStream_t s) {
  if (rows <= 0 || cols <= 0) return cudaSuccess;
  if (rows == ld_dst && rows == ld_src)
    return cudaMemcpyAsync(dst, src, (size_t)(rows * cols * es), kind, s);
  return cudaMemcpy2DAsync(dst, (size_t)(ld_dst * es), src, (size_t)(ld_src * es), (size_t)(rows * es), (size_t)cols, kind, s);
}

}  // namespace

extern "C" {

// [start, start + count) of part `index` when `total` units are cut into `parts` shares that are multiples of `align`
// (except possibly the last non-empty one); earlier parts take the remainder.
int pbx_shard_range(int64_t total, int parts, int index, int64_t align, int64_t* start, int64_t* count) {
  if (total < 0 || parts <= 0 || index < 0 || index >= parts || align <= 0 || !start || !count) return PBX_ERR_INVALID_ARG;
  const int64_t units = (total + align - 1) / align;
  const int64_t base = units / parts, rem = units % parts;
  const int64_t u0 = index * base + (index < rem ? index : rem);
  const int64_t cnt = base + (index < rem ? 1 : 0);
  int64_t s = u0 * align, e = (u0 + cnt) * align;
  if (s > total) s = total;
  if (e > total) e = total;
  *start = s;
  *count = e - s;
  return PBX_OK;
}

int pbx_multi_create(pbx_multi_t* out, int n_dev, const int* device_ordinals) {
  if (!out || n_dev < 1 || n_dev > 8) return PBX_ERR_INVALID_ARG;
  *out = nullptr;
  int have = 0;
  if (cudaGetDeviceCount(&have) != cudaSuccess || have < 1) return PBX_ERR_NO_DEVICE;
  pbx_multi_s* mh = new pbx_multi_s();
  int prev = 0;
  cudaGetDevice(&prev);
  int st = PBX_OK;
  for (int g = 0; g < n_dev && st == PBX_OK; ++g) {
    const int dev = device_ordinals ? device_ordinals[g] : g;
    if (dev < 0 || dev >= have) { st = PBX_ERR_NO_DEVICE; break; }
    cudaStream_t s = nullptr;
    cudaEvent_t e = nullptr;
    if (cudaSetDevice(dev) != cudaSuccess || cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) { st = PBX_ERR_CUDA; break; }
    pbx_handle_t h = nullptr;
    st = pbx_create(&h, dev, (void*)s);
    if (st != PBX_OK) { cudaStreamDestroy(s); cudaEventDestroy(e); break; }
    mh->h.push_back(h); mh->stream.push_back(s); mh->ev.push_back(e);
  }
  // every device may address every other device's memory (NVLink / NVSwitch on an 8 x B200 box)
  for (size_t a = 0; a < mh->h.size() && st == PBX_OK; ++a) {
    cudaSetDevice(mh->h[a]->device);
    for (size_t b = 0; b < mh->h.size(); ++b) {
      if (a == b || mh->h[a]->device == mh->h[b]->device) continue;
      int can = 0;
      cudaDeviceCanAccessPeer(&can, mh->h[a]->device, mh->h[b]->device);
      if (!can) { mh->last_error = "devices of the group cannot address each other's memory"; st = PBX_ERR_CUDA; break; }
      const cudaError_t e = cudaDeviceEnablePeerAccess(mh->h[b]->device, 0);
      if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { st = PBX_ERR_CUDA; break; }
      cudaGetLastError();
    }
  }
  cudaSetDevice(prev);
  const size_t n = mh->h.size();
  mh->dA.assign(n, nullptr); mh->dB.assign(n, nullptr); mh->dC.assign(n, nullptr);
  mh->dA_bytes.assign(n, 0); mh->dB_bytes.assign(n, 0); mh->dC_bytes.assign(n, 0);
  if (st != PBX_OK) { pbx_multi_destroy(mh); return st; }
  *out = mh;
  return PBX_OK;
}

int pbx_multi_destroy(pbx_multi_t mh) {
  if (!mh) return PBX_ERR_INVALID_ARG;
  for (size_t g = 0; g < mh->h.size(); ++g) {
    PbxDeviceGuard guard(mh->h[g]);
    cudaStreamSynchronize(mh->stream[g]);
    if (g < mh->dA.size()) {
      if (mh->dA[g]) cudaFree(mh->dA[g]);
      if (mh->dB[g]) cudaFree(mh->dB[g]);
      if (mh->dC[g]) cudaFree(mh->dC[g]);
    }
    pbx_destroy(mh->h[g]);
    cudaEventDestroy(mh->ev[g]);
    cudaStreamDestroy(mh->stream[g]);
  }
  delete mh;
  return PBX_OK;
}

int pbx_multi_device_count(pbx_multi_t mh) { return mh ? (int)mh->h.size() : 0; }
pbx_handle_t pbx_multi_handle(pbx_multi_t mh, int i) { return (mh && i >= 0 && i < (int)mh->h.size()) ? mh->h[i] : nullptr; }
const char* pbx_multi_last_error(pbx_multi_t mh) { return mh ? mh->last_error.c_str() : "null group"; }

int pbx_multi_synchronize(pbx_multi_t mh) {
  if (!mh) return PBX_ERR_INVALID_ARG;
  for (size_t g = 0; g < mh->h.size(); ++g) {
    const int st = pbx_synchronize(mh->h[g]);
    if (st != PBX_OK) { mh->last_error = pbx_last_error(mh->h[g]); return st; }
  }
  return PBX_OK;
}

// M-block sharded GEMM on device-resident shards.  A_blocks[g]: device g's rows of op(A) (its M-block; for transa == 'n'
// a (rows_g x K) matrix with leading dimension lda, for 't' a (K x rows_g) one); B_full[g]: the whole op(B) operand on
// device g; C_full[g]: device g's m x n matrix C (leading dimension ldc).  Device g computes rows
// pbx_shard_range(m, G, g, 256).  gather == 0: it writes them into its own C only.  gather != 0: into every device's
// C_full (and reads beta*C from its own), so that after pbx_multi_synchronize all devices hold the full result.
int pbx_gemm_sharded(pbx_multi_t mh, int dtype, char transa, char transb, int64_t m, int64_t n, int64_t k, const void* alpha,
                     const void* const* A_blocks, int64_t lda, const void* const* B_full, int64_t ldb, const void* beta,
                     void* const* C_full, int64_t ldc, int gather) {
  if (!mh || !A_blocks || !B_full || !C_full || m < 0 || n < 0 || k < 0) return PBX_ERR_INVALID_ARG;
  const int G = (int)mh->h.size();
  const size_t eo = pbx_out_size(dtype);
  for (int g = 0; g < G; ++g) {
    int64_t row0 = 0, rows = 0;
    pbx_shard_range(m, G, g, 256, &row0, &rows);
    if (rows == 0) continue;
    void* c_list[8];
    int n_dst = 0;
    c_list[n_dst++] = (char*)C_full[g] + (size_t)row0 * eo;
    if (gather)
      for (int x = 0; x < G; ++x)
        if (x != g) c_list[n_dst++] = (char*)C_full[x] + (size_t)row0 * eo;
    const int st = pbx_gemm_multicast(mh->h[g], dtype, transa, transb, rows, n, k, alpha, A_blocks[g], lda, B_full[g], ldb,
                                      beta, c_list, n_dst, ldc);
    if (st != PBX_OK) { mh->last_error = pbx_last_error(mh->h[g]); return st; }
  }
  return PBX_OK;
}

// Batch-sharded strided GEMM on device-resident shards: device g holds batch entries pbx_shard_range(batch, G, g, 1)
// of A, B and C (A_shards[g] etc. point at the FIRST entry the device owns; strides as in pbx_gemm).
int pbx_gemm_strided_batched_sharded(pbx_multi_t mh, int dtype, char transa, char transb, int64_t m, int64_t n, int64_t k,
                                     const void* alpha, const void* const* A_shards, int64_t lda, int64_t stridea,
                                     const void* const* B_shards, int64_t ldb, int64_t strideb, const void* beta,
                                     void* const* C_shards, int64_t ldc, int64_t stridec, int64_t batch) {
  if (!mh || !A_shards || !B_shards || !C_shards || batch < 0) return PBX_ERR_INVALID_ARG;
  const int G = (int)mh->h.size();
  for (int g = 0; g < G; ++g) {
    int64_t b0 = 0, cnt = 0;
    pbx_shard_range(batch, G, g, 1, &b0, &cnt);
    if (cnt == 0) continue;
    const int st = pbx_gemm(mh->h[g], dtype, transa, transb, m, n, k, alpha, A_shards[g], lda, stridea, B_shards[g], ldb,
                            strideb, beta, C_shards[g], ldc, stridec, cnt, 0);
    if (st != PBX_OK) { mh->last_error = pbx_last_error(mh->h[g]); return st; }
  }
  return PBX_OK;
}

// blas::_gemm with HOST operands over all devices of the group (same arguments as pbx_gemm_host, batch == 1).
// Synchronous: returns when C_host holds the result.
int pbx_gemm_sharded_host(pbx_multi_t mh, int dtype, char transa, char transb, int64_t m, int64_t n, int64_t k,
                          const void* alpha, const void* A_host, int64_t lda, const void* B_host, int64_t ldb,
                          const void* beta, void* C_host, int64_t ldc) {
  if (!mh || dtype < PBX_F32 || dtype > PBX_BF16_F32 || !alpha || !beta || m < 0 || n < 0 || k < 0) return PBX_ERR_INVALID_ARG;
  const int G = (int)mh->h.size();
  const int ta_c = tolower((unsigned char)transa), tb_c = tolower((unsigned char)transb);
  const double al = dtype == PBX_F64 ? *(const double*)alpha : (double)*(const float*)alpha;
  const double be = dtype == PBX_F64 ? *(const double*)beta : (double)*(const float*)beta;
  const bool valid = (ta_c == 'n' || ta_c == 't' || ta_c == 'c') && (tb_c == 'n' || tb_c == 't' || tb_c == 'c');
  if (G == 1 || al == 0.0 || !valid || m == 0 || n == 0 || k == 0 || !A_host || !B_host || !C_host)
    return pbx_gemm_host(mh->h[0], dtype, transa, transb, m, n, k, alpha, A_host, lda, 0, B_host, ldb, 0, beta, C_host, ldc, 0,
                         1, 0);   // front-end shortcuts and errors: the single-device rules
  const bool ta = ta_c != 'n', tb = tb_c != 'n';
  const int64_t es = (int64_t)pbx_in_size(dtype), eo = (int64_t)pbx_out_size(dtype);
  // stored shape of B is (b_rows x b_cols); device g uploads the slice of B that holds columns [n0_g, n0_g + nn_g) of op(B)
  const int64_t b_rows = tb ? n : k, b_cols = tb ? k : n;
  const int64_t ldb_d = b_rows;   // compact device copies
  struct Part { int64_t row0, rows, n0, nn; };
  std::vector<Part> part(G);
  for (int g = 0; g < G; ++g) {
    pbx_shard_range(m, G, g, 256, &part[g].row0, &part[g].rows);
    pbx_shard_range(n, G, g, 256, &part[g].n0, &part[g].nn);
  }
  int st = PBX_OK;
  // ---- 1. uploads: A block, own B slice (and own C block when beta != 0), all devices concurrently ----
  for (int g = 0; g < G; ++g) {
    const Part& p = part[g];
    const int64_t a_rows = ta ? k : p.rows, a_cols = ta ? p.rows : k;
    if ((st = ensure_buf(mh, g, mh->dA, mh->dA_bytes, a_rows * a_cols * es)) ||
        (st = ensure_buf(mh, g, mh->dB, mh->dB_bytes, b_rows * b_cols * es)) ||
        (st = ensure_buf(mh, g, mh->dC, mh->dC_bytes, p.rows * n * eo)))
      return st;
    PbxDeviceGuard guard(mh->h[g]);
    cudaStream_t s = mh->stream[g];
    const char* a_src = (const char*)A_host + (ta ? p.row0 * lda : p.row0) * es;
    PBX_CUDA_CHECK(mh->h[g], copy2d(mh->dA[g], a_rows, a_src, lda, a_rows, a_cols, es, cudaMemcpyHostToDevice, s));
    if (p.nn > 0) {
      if (tb) {   // B stored n x k: rows [n0, n0 + nn) of every column
        PBX_CUDA_CHECK(mh->h[g], copy2d((char*)mh->dB[g] + p.n0 * es, ldb_d, (const char*)B_host + p.n0 * es, ldb, p.nn, k, es,
                                        cudaMemcpyHostToDevice, s));
      } else {    // B stored k x n: columns [n0, n0 + nn)
        PBX_CUDA_CHECK(mh->h[g], copy2d((char*)mh->dB[g] + p.n0 * ldb_d * es, ldb_d, (const char*)B_host + p.n0 * ldb * es, ldb, k,
                                        p.nn, es, cudaMemcpyHostToDevice, s));
      }
    }
    if (be != 0.0 && p.rows > 0)
      PBX_CUDA_CHECK(mh->h[g], copy2d(mh->dC[g], p.rows, (const char*)C_host + p.row0 * eo, ldc, p.rows, n, eo,
                                      cudaMemcpyHostToDevice, s));
    PBX_CUDA_CHECK(mh->h[g], cudaEventRecord(mh->ev[g], s));
  }
  // ---- 2. exchange the B slices device to device (each device pulls the others' slices over NVLink) ----
  for (int g = 0; g < G; ++g) {
    PbxDeviceGuard guard(mh->h[g]);
    cudaStream_t s = mh->stream[g];
    for (int o = 1; o < G; ++o) {
      const int x = (g + o) % G;   // staggered so that no device is everybody's first source
      const Part& q = part[x];
      if (q.nn == 0) continue;
      PBX_CUDA_CHECK(mh->h[g], cudaStreamWaitEvent(s, mh->ev[x], 0));
      if (tb) {
        PBX_CUDA_CHECK(mh->h[g], copy2d((char*)mh->dB[g] + q.n0 * es, ldb_d, (const char*)mh->dB[x] + q.n0 * es, ldb_d, q.nn, k,
                                        es, cudaMemcpyDeviceToDevice, s));
      } else {
        PBX_CUDA_CHECK(mh->h[g], cudaMemcpyAsync((char*)mh->dB[g] + q.n0 * ldb_d * es, (const char*)mh->dB[x] + q.n0 * ldb_d * es,
                                                 (size_t)(k * q.nn * es), cudaMemcpyDeviceToDevice, s));
      }
    }
  }
  // ---- 3. compute and download ----
  for (int g = 0; g < G; ++g) {
    const Part& p = part[g];
    if (p.rows == 0) continue;
    const int64_t lda_d = ta ? k : p.rows;
    st = pbx_gemm(mh->h[g], dtype, transa, transb, p.rows, n, k, alpha, mh->dA[g], lda_d, 0, mh->dB[g], ldb_d, 0, beta, mh->dC[g],
                  p.rows, 0, 1, 0);
    if (st != PBX_OK) { mh->last_error = pbx_last_error(mh->h[g]); return st; }
    PbxDeviceGuard guard(mh->h[g]);
    PBX_CUDA_CHECK(mh->h[g], copy2d((char*)C_host + p.row0 * eo, ldc, mh->dC[g], p.rows, p.rows, n, eo, cudaMemcpyDeviceToHost,
                                    mh->stream[g]));
  }
  // a device's B must stay intact until every peer has pulled its slice: the final synchronize covers that
  return pbx_multi_synchronize(mh);
}

}  // extern "C"
