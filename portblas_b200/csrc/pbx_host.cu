// pbx_host.cu -- pbx_gemm_host: the GEMM path with HOST operands (end-to-end metric).
//
// What a portBLAS caller does around one GEMM (reference samples/gemm.cpp:50-63,
// include/portblas_helper.h:139-191): copy_to_device(A), copy_to_device(B), _gemm, copy_to_host(C),
// wait.  Done naively that serialises PCIe and the tensor pipe.  Here the call is cut into
// independent column panels of op(B)/C (batch == 1) or batch ranges (strided batches):
//
//     copy-in stream :  A | B_0 | B_1 | B_2 | ...
//     handle stream  :      |gemm_0|gemm_1|gemm_2| ...          (each panel is an ordinary pbx_gemm)
//     copy-out stream:             | C_0 | C_1 | C_2 | ...
//
// so the upload of panel j+1 and the download of panel j-1 overlap the compute of panel j
// (H2D and D2H use separate copy engines).  Only the M x N window of C is ever written on the host
// (2-D copies), so ld padding in the caller's buffer is preserved; C is uploaded only when beta != 0.
//
// Deep single GEMMs whose output is kept in the accumulate precision (fp32 / fp64 C) go one step further:
// the scheme above cannot start computing before ALL of A has crossed PCIe, and cannot finish before the
// last C panel has gone back.  So the column blocks shrink geometrically (n/2, n/4, ..., the last one is the
// download tail) and the FIRST block is additionally streamed along K,
//
//     copy-in :  A_k0 B0_k0 | A_k1 B0_k1 | ... | A_k7 B0_k7 | B1 | B2 | B3
//     compute :             | C0 = a A_k0 B0_k0 + b C0 | C0 += a A_k1 B0_k1 | ... | C1 | C2 | C3
//     copy-out:                                                              | C0 | C1 | C2 | C3
//
// so the first MMA starts after 1/8 of A and 1/16 of B have arrived and A is resident once block 0 is done.
#include <ctype.h>
#include <stdio.h>
#include <stdlib.h>

#include "pbx_internal.cuh"

namespace {

struct Copy2D {
  // rows x cols window of a column-major matrix with leading dimension ld (elements of es bytes)
  static cudaError_t run(void* dst, const void* src, int64_t rows, int64_t cols, int64_t ld, int64_t es,
                         cudaMemcpyKind kind, cudaStream_t s) {
    if (rows <= 0 || cols <= 0) return cudaSuccess;
    if (rows == ld || cols == 1)
      return cudaMemcpyAsync(dst, src, (size_t)(((cols - 1) * ld + rows) * es), kind, s);
    return cudaMemcpy2DAsync(dst, (size_t)(ld * es), src, (size_t)(ld * es), (size_t)(rows * es), (size_t)cols,
                             kind, s);
  }
};

int ensure_stage(pbx_handle_t h, int i, int64_t bytes) {
  if (bytes <= h->stage_bytes[i]) return PBX_OK;
  if (h->stage[i]) {
    PBX_CUDA_CHECK(h, cudaStreamSynchronize(h->stream));
    PBX_CUDA_CHECK(h, cudaFree(h->stage[i]));
    h->stage[i] = nullptr; h->stage_bytes[i] = 0;
  }
  PBX_CUDA_CHECK(h, cudaMalloc(&h->stage[i], (size_t)(bytes > 0 ? bytes : 1)));
  h->stage_bytes[i] = bytes;
  return PBX_OK;
}

int ensure_pipe(pbx_handle_t h, int n_events) {
  if (!h->s_in) PBX_CUDA_CHECK(h, cudaStreamCreateWithFlags(&h->s_in, cudaStreamNonBlocking));
  if (!h->s_out) PBX_CUDA_CHECK(h, cudaStreamCreateWithFlags(&h->s_out, cudaStreamNonBlocking));
  while ((int)h->events.size() < n_events) {
    cudaEvent_t e;
    PBX_CUDA_CHECK(h, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    h->events.push_back(e);
  }
  return PBX_OK;
}

}  // namespace

extern "C" int pbx_gemm_host(pbx_handle_t h, int dtype, char transa, char transb, int64_t m, int64_t n,
                             int64_t k, const void* alpha, const void* A_host, int64_t lda, int64_t stridea,
                             const void* B_host, int64_t ldb, int64_t strideb, const void* beta, void* C_host,
                             int64_t ldc, int64_t stridec, int64_t batch, int batch_type) {
  if (!h || dtype < PBX_F32 || dtype > PBX_BF16_F32 || !alpha || !beta || m < 0 || n < 0 || k < 0 || batch < 0 ||
      (batch_type != 0 && batch_type != 1))
    return PBX_ERR_INVALID_ARG;
  PBX_DEVICE_GUARD(h);
  const int ta_c = tolower((unsigned char)transa), tb_c = tolower((unsigned char)transb);
  {
    // pbx_gemm's front-end rules, in its order and with its codes, BEFORE any buffer size is derived from the strides
    // (gemm_interface.hpp:112-166): alpha == 0 skips the validation there; this entry point additionally refuses
    // negative strides in that case, because it sizes host <-> device copies from them.
    const double al0 = dtype == PBX_F64 ? *(const double*)alpha : (double)*(const float*)alpha;
    const bool strided = batch > 1 && batch_type == 0;
    if (al0 != 0.0) {
      if (ta_c != 'n' && ta_c != 't' && ta_c != 'c') return PBX_ERR_INVALID_TRANSA;
      if (tb_c != 'n' && tb_c != 't' && tb_c != 'c') return PBX_ERR_INVALID_TRANSB;
      if (strided) {
        if (stridec < ldc * n || stridec < 0) return PBX_ERR_INVALID_STRIDEC;
        if (stridea < 0) return PBX_ERR_INVALID_STRIDEA;
        if (strideb < 0) return PBX_ERR_INVALID_STRIDEB;
      }
    } else if (strided && (stridea < 0 || strideb < 0 || stridec < 0)) {
      return PBX_ERR_INVALID_ARG;
    }
    if (batch == 0 || m == 0 || n == 0) return PBX_OK;   // nothing to compute, nothing to copy (pbx_gemm: no-op)
  }
  const bool ta = ta_c != 'n', tb = tb_c != 'n';
  const int64_t es = (int64_t)pbx_in_size(dtype), eo = (int64_t)pbx_out_size(dtype);
  const int64_t a_rows = ta ? k : m, a_cols = ta ? m : k;   // stored shapes
  const int64_t b_rows = tb ? n : k, b_cols = tb ? k : n;
  const bool ilv = (batch_type == 1 && batch > 1);
  if (batch == 1) stridea = strideb = stridec = 0;
  int64_t a_el, b_el, c_el;
  if (ilv) {
    a_el = lda * a_cols * batch; b_el = ldb * b_cols * batch; c_el = ldc * n * batch;
  } else {
    a_el = (batch - 1) * stridea + lda * a_cols;
    b_el = (batch - 1) * strideb + ldb * b_cols;
    c_el = (batch - 1) * stridec + ldc * n;
  }
  int st;
  if ((st = ensure_stage(h, 0, a_el * es)) || (st = ensure_stage(h, 1, b_el * es)) ||
      (st = ensure_stage(h, 2, c_el * eo)))
    return st;
  char* dA = (char*)h->stage[0];
  char* dB = (char*)h->stage[1];
  char* dC = (char*)h->stage[2];
  const char* hA = (const char*)A_host;
  const char* hB = (const char*)B_host;
  char* hC = (char*)C_host;
  const double al = dtype == PBX_F64 ? *(const double*)alpha : (double)*(const float*)alpha;
  const double be = dtype == PBX_F64 ? *(const double*)beta : (double)*(const float*)beta;
  const bool need_c_in = (be != 0.0);
  const bool valid_trans = (ta_c == 'n' || ta_c == 't' || ta_c == 'c') && (tb_c == 'n' || tb_c == 't' || tb_c == 'c');
  const double work = 2.0 * (double)m * (double)n * (double)k * (double)batch;
  const double bytes = (double)(a_el + b_el) * es + (double)c_el * eo;

  // ---- simple path: interleaved batches, alpha == 0 / invalid arguments (let pbx_gemm decide), small calls ----
  if (ilv || al == 0.0 || !valid_trans || k == 0 || m == 0 || n == 0 || (bytes < 16e6 && work < 4e10)) {
    if (ilv || al == 0.0 || !valid_trans || k == 0) {
      // whole buffers (interleaved layout has no simple 2-D window)
      if (a_el > 0 && hA) PBX_CUDA_CHECK(h, cudaMemcpyAsync(dA, hA, (size_t)(a_el * es), cudaMemcpyHostToDevice, h->stream));
      if (b_el > 0 && hB) PBX_CUDA_CHECK(h, cudaMemcpyAsync(dB, hB, (size_t)(b_el * es), cudaMemcpyHostToDevice, h->stream));
      if (c_el > 0) PBX_CUDA_CHECK(h, cudaMemcpyAsync(dC, hC, (size_t)(c_el * eo), cudaMemcpyHostToDevice, h->stream));
      st = pbx_gemm(h, dtype, transa, transb, m, n, k, alpha, dA, lda, stridea, dB, ldb, strideb, beta, dC, ldc,
                    stridec, batch, batch_type);
      if (st != PBX_OK) return st;
      if (c_el > 0) PBX_CUDA_CHECK(h, cudaMemcpyAsync(hC, dC, (size_t)(c_el * eo), cudaMemcpyDeviceToHost, h->stream));
      PBX_CUDA_CHECK(h, cudaStreamSynchronize(h->stream));
      return PBX_OK;
    }
    for (int64_t b = 0; b < batch; ++b) {
      if (b == 0 || stridea > 0)
        PBX_CUDA_CHECK(h, Copy2D::run(dA + b * stridea * es, hA + b * stridea * es, a_rows, a_cols, lda, es,
                                      cudaMemcpyHostToDevice, h->stream));
      if (b == 0 || strideb > 0)
        PBX_CUDA_CHECK(h, Copy2D::run(dB + b * strideb * es, hB + b * strideb * es, b_rows, b_cols, ldb, es,
                                      cudaMemcpyHostToDevice, h->stream));
      if (need_c_in)
        PBX_CUDA_CHECK(h, Copy2D::run(dC + b * stridec * eo, hC + b * stridec * eo, m, n, ldc, eo,
                                      cudaMemcpyHostToDevice, h->stream));
    }
    st = pbx_gemm(h, dtype, transa, transb, m, n, k, alpha, dA, lda, stridea, dB, ldb, strideb, beta, dC, ldc,
                  stridec, batch, batch_type);
    if (st != PBX_OK) return st;
    for (int64_t b = 0; b < batch; ++b)
      PBX_CUDA_CHECK(h, Copy2D::run(hC + b * stridec * eo, dC + b * stridec * eo, m, n, ldc, eo,
                                    cudaMemcpyDeviceToHost, h->stream));
    PBX_CUDA_CHECK(h, cudaStreamSynchronize(h->stream));
    return PBX_OK;
  }

  // ---- pipelined path ----
  // stride validation happens inside pbx_gemm per chunk as well, but the chunked copies below already
  // dereference with the strides: validate first (same order and codes as pbx_gemm).
  if (batch > 1) {
    if (stridec < ldc * n) return PBX_ERR_INVALID_STRIDEC;
    if (stridea < 0) return PBX_ERR_INVALID_STRIDEA;
    if (strideb < 0) return PBX_ERR_INVALID_STRIDEB;
  }
  const int MAXP = 32;
  int panels;
  if (batch > 1) {
    panels = (int)(batch < 8 ? batch : 8);
  } else {
    // column panels of >= 1024 columns, at most 8; keep panels multiples of 256 columns (tile width)
    panels = (int)(n / 1024);
    if (panels > 8) panels = 8;
    if (panels < 1) panels = 1;
  }
  if (panels > MAXP) panels = MAXP;
  if ((st = ensure_pipe(h, 2 + 2 * MAXP)) != PBX_OK) return st;
  cudaEvent_t ev_start = h->events[0], ev_a = h->events[1];
  cudaEvent_t* ev_in = &h->events[2];
  cudaEvent_t* ev_done = &h->events[2 + MAXP];

  PBX_CUDA_CHECK(h, cudaEventRecord(ev_start, h->stream));   // order after earlier work on the handle's stream
  PBX_CUDA_CHECK(h, cudaStreamWaitEvent(h->s_in, ev_start, 0));
  PBX_CUDA_CHECK(h, cudaStreamWaitEvent(h->s_out, ev_start, 0));

  static const int kstream_env = getenv("PBX_HOST_KSTREAM") ? atoi(getenv("PBX_HOST_KSTREAM")) : 1;
  if (batch == 1 && eo >= 4 && k >= 2048 && n >= 2048 && kstream_env) {
    // ---- K-streamed first block + geometrically shrinking column blocks ----
    int64_t blk_n0[8], blk_nb[8];
    int nblk = 0;
    {
      int64_t n0 = 0, rest = n;
      while (rest > 0 && nblk < 7) {
        int64_t nb = ((rest / 2) + 255) / 256 * 256;
        if (nb < 1024 || rest - nb < 512) nb = rest;
        blk_n0[nblk] = n0; blk_nb[nblk] = nb; ++nblk;
        n0 += nb; rest -= nb;
      }
      if (rest > 0) { blk_n0[nblk] = n0; blk_nb[nblk] = rest; ++nblk; }
    }
    int kp_count = (int)(k / 1024);
    if (kp_count > 8) kp_count = 8;
    const int64_t kper = (((k + kp_count - 1) / kp_count) + 255) / 256 * 256;
    const double one_d = 1.0;
    const float one_f = 1.0f;
    const void* one = (dtype == PBX_F64) ? (const void*)&one_d : (const void*)&one_f;
    int e = 0;
    const int64_t nb0 = blk_nb[0];
    for (int64_t k0 = 0; k0 < k; k0 += kper) {
      const int64_t kb = (k - k0 < kper) ? (k - k0) : kper;
      const int64_t a_off = (ta ? k0 : k0 * lda) * es;            // op(A)[:, k0:] inside the stored matrix
      const int64_t b_off = (tb ? k0 * ldb : k0) * es;            // op(B)[k0:, 0:nb0]
      PBX_CUDA_CHECK(h, Copy2D::run(dA + a_off, hA + a_off, ta ? kb : m, ta ? m : kb, lda, es, cudaMemcpyHostToDevice,
                                    h->s_in));
      PBX_CUDA_CHECK(h, Copy2D::run(dB + b_off, hB + b_off, tb ? nb0 : kb, tb ? kb : nb0, ldb, es,
                                    cudaMemcpyHostToDevice, h->s_in));
      if (k0 == 0 && need_c_in)
        PBX_CUDA_CHECK(h, Copy2D::run(dC, hC, m, nb0, ldc, eo, cudaMemcpyHostToDevice, h->s_in));
      PBX_CUDA_CHECK(h, cudaEventRecord(ev_in[e], h->s_in));
      PBX_CUDA_CHECK(h, cudaStreamWaitEvent(h->stream, ev_in[e], 0));
      ++e;
      st = pbx_gemm(h, dtype, transa, transb, m, nb0, kb, alpha, dA + a_off, lda, 0, dB + b_off, ldb, 0,
                    k0 == 0 ? beta : one, dC, ldc, 0, 1, 0);
      if (st != PBX_OK) { cudaStreamSynchronize(h->s_in); return st; }
    }
    for (int j = 0; j < nblk; ++j) {
      const int64_t n0 = blk_n0[j], nb = blk_nb[j];
      if (j > 0) {   // A is resident: one upload of op(B)[:, n0:n0+nb] and one full-K GEMM per block
        const int64_t b_off = (tb ? n0 : n0 * ldb) * es;
        PBX_CUDA_CHECK(h, Copy2D::run(dB + b_off, hB + b_off, tb ? nb : k, tb ? k : nb, ldb, es, cudaMemcpyHostToDevice,
                                      h->s_in));
        if (need_c_in)
          PBX_CUDA_CHECK(h, Copy2D::run(dC + n0 * ldc * eo, hC + n0 * ldc * eo, m, nb, ldc, eo, cudaMemcpyHostToDevice,
                                        h->s_in));
        PBX_CUDA_CHECK(h, cudaEventRecord(ev_in[e], h->s_in));
        PBX_CUDA_CHECK(h, cudaStreamWaitEvent(h->stream, ev_in[e], 0));
        ++e;
        st = pbx_gemm(h, dtype, transa, transb, m, nb, k, alpha, dA, lda, 0, dB + b_off, ldb, 0, beta,
                      dC + n0 * ldc * eo, ldc, 0, 1, 0);
        if (st != PBX_OK) { cudaStreamSynchronize(h->s_in); return st; }
      }
      PBX_CUDA_CHECK(h, cudaEventRecord(ev_done[j], h->stream));
      PBX_CUDA_CHECK(h, cudaStreamWaitEvent(h->s_out, ev_done[j], 0));
      PBX_CUDA_CHECK(h, Copy2D::run(hC + n0 * ldc * eo, dC + n0 * ldc * eo, m, nb, ldc, eo, cudaMemcpyDeviceToHost,
                                    h->s_out));
    }
  } else if (batch == 1) {
    PBX_CUDA_CHECK(h, Copy2D::run(dA, hA, a_rows, a_cols, lda, es, cudaMemcpyHostToDevice, h->s_in));
    PBX_CUDA_CHECK(h, cudaEventRecord(ev_a, h->s_in));
    PBX_CUDA_CHECK(h, cudaStreamWaitEvent(h->stream, ev_a, 0));
    const int64_t per = (((n + panels - 1) / panels) + 255) / 256 * 256;
    int j = 0;
    for (int64_t n0 = 0; n0 < n; n0 += per, ++j) {
      const int64_t nb = (n - n0 < per) ? (n - n0) : per;
      // op(B) columns [n0, n0+nb): stored columns (tb == false) or stored rows (tb == true)
      if (!tb) {
        PBX_CUDA_CHECK(h, Copy2D::run(dB + n0 * ldb * es, hB + n0 * ldb * es, k, nb, ldb, es,
                                      cudaMemcpyHostToDevice, h->s_in));
      } else {
        PBX_CUDA_CHECK(h, Copy2D::run(dB + n0 * es, hB + n0 * es, nb, k, ldb, es, cudaMemcpyHostToDevice, h->s_in));
      }
      if (need_c_in)
        PBX_CUDA_CHECK(h, Copy2D::run(dC + n0 * ldc * eo, hC + n0 * ldc * eo, m, nb, ldc, eo,
                                      cudaMemcpyHostToDevice, h->s_in));
      PBX_CUDA_CHECK(h, cudaEventRecord(ev_in[j], h->s_in));
      PBX_CUDA_CHECK(h, cudaStreamWaitEvent(h->stream, ev_in[j], 0));
      const char* dBj = tb ? dB + n0 * es : dB + n0 * ldb * es;
      st = pbx_gemm(h, dtype, transa, transb, m, nb, k, alpha, dA, lda, 0, dBj, ldb, 0, beta, dC + n0 * ldc * eo,
                    ldc, 0, 1, 0);
      if (st != PBX_OK) { cudaStreamSynchronize(h->s_in); return st; }
      PBX_CUDA_CHECK(h, cudaEventRecord(ev_done[j], h->stream));
      PBX_CUDA_CHECK(h, cudaStreamWaitEvent(h->s_out, ev_done[j], 0));
      PBX_CUDA_CHECK(h, Copy2D::run(hC + n0 * ldc * eo, dC + n0 * ldc * eo, m, nb, ldc, eo,
                                    cudaMemcpyDeviceToHost, h->s_out));
    }
  } else {
    const int64_t per = (batch + panels - 1) / panels;
    // footprints: contiguous ranges when the stride equals the matrix footprint (one copy per chunk)
    auto copy_range = [&](char* d, const char* s, int64_t b0, int64_t cnt, int64_t stride, int64_t rows,
                          int64_t cols, int64_t ld, int64_t e, cudaMemcpyKind kind, cudaStream_t strm) -> cudaError_t {
      if (rows == ld && stride == ld * cols)
        return cudaMemcpyAsync(d + b0 * stride * e, s + b0 * stride * e, (size_t)(cnt * stride * e), kind, strm);
      for (int64_t b = b0; b < b0 + cnt; ++b) {
        cudaError_t r = Copy2D::run(d + b * stride * e, s + b * stride * e, rows, cols, ld, e, kind, strm);
        if (r != cudaSuccess) return r;
      }
      return cudaSuccess;
    };
    if (stridea == 0) PBX_CUDA_CHECK(h, Copy2D::run(dA, hA, a_rows, a_cols, lda, es, cudaMemcpyHostToDevice, h->s_in));
    if (strideb == 0) PBX_CUDA_CHECK(h, Copy2D::run(dB, hB, b_rows, b_cols, ldb, es, cudaMemcpyHostToDevice, h->s_in));
    int j = 0;
    for (int64_t b0 = 0; b0 < batch; b0 += per, ++j) {
      const int64_t cnt = (batch - b0 < per) ? (batch - b0) : per;
      if (stridea > 0)
        PBX_CUDA_CHECK(h, copy_range(dA, hA, b0, cnt, stridea, a_rows, a_cols, lda, es, cudaMemcpyHostToDevice, h->s_in));
      if (strideb > 0)
        PBX_CUDA_CHECK(h, copy_range(dB, hB, b0, cnt, strideb, b_rows, b_cols, ldb, es, cudaMemcpyHostToDevice, h->s_in));
      if (need_c_in)
        PBX_CUDA_CHECK(h, copy_range(dC, hC, b0, cnt, stridec, m, n, ldc, eo, cudaMemcpyHostToDevice, h->s_in));
      PBX_CUDA_CHECK(h, cudaEventRecord(ev_in[j], h->s_in));
      PBX_CUDA_CHECK(h, cudaStreamWaitEvent(h->stream, ev_in[j], 0));
      st = pbx_gemm(h, dtype, transa, transb, m, n, k, alpha, dA + b0 * stridea * es, lda, stridea,
                    dB + b0 * strideb * es, ldb, strideb, beta, dC + b0 * stridec * eo, ldc, stridec, cnt, 0);
      if (st != PBX_OK) { cudaStreamSynchronize(h->s_in); return st; }
      PBX_CUDA_CHECK(h, cudaEventRecord(ev_done[j], h->stream));
      PBX_CUDA_CHECK(h, cudaStreamWaitEvent(h->s_out, ev_done[j], 0));
      PBX_CUDA_CHECK(h, copy_range(hC, dC, b0, cnt, stridec, m, n, ldc, eo, cudaMemcpyDeviceToHost, h->s_out));
    }
  }
  PBX_CUDA_CHECK(h, cudaStreamSynchronize(h->s_out));
  PBX_CUDA_CHECK(h, cudaStreamSynchronize(h->stream));
  return PBX_OK;
}
