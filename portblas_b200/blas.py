"""Host-side mirror of the reference's GEMM interface, in Python, over the C-ABI.

Names, argument order, argument meaning and error behaviour follow the reference:

  blas::SB_Handle                    include/sb_handle/portblas_handle.h:46-200
  blas::gemm_batch_type_t            include/operations/blas3_trees.h:59
  blas::_gemm                        include/interface/blas3_interface.h:88-95
  blas::_gemm_batched                include/interface/blas3_interface.h:99-109
  blas::_gemm_strided_batched        include/interface/blas3_interface.h:113-123
  blas::_trsm                        include/interface/blas3_interface.h:125-135
  blas::_symm                        include/interface/blas3_interface.h:137-147
  complex _gemm* (BLAS_ENABLE_COMPLEX): torch.complex64 / complex128 containers, complex alpha / beta

so that tests/ read like test/unittest/blas3/*gemm*.  (The C++ mirror of the same interface lives
in include/portblas.hpp; both funnel into the same extern "C" entry points.)

Containers are 1-D torch CUDA tensors -- the analogue of the reference's USM pointers /
BufferIterator (``buf[offset:]`` plays the role of ``buffer + offset``).  std::invalid_argument
maps to ValueError with the reference's exact message.  Calls are asynchronous on the handle's
stream, as in the reference; ``SB_Handle.wait()`` synchronises.
"""
from __future__ import annotations

import ctypes
import enum
from typing import Optional

import torch

from . import _lib


class gemm_batch_type_t(enum.IntEnum):
    strided = 0
    interleaved = 1


class PbxError(RuntimeError):
    pass


def _check(h: "SB_Handle", status: int) -> None:
    if status == _lib.OK:
        return
    lib = _lib.load()
    text = lib.pbx_status_string(status).decode()
    if 1 <= status <= 5 or 10 <= status <= 16:
        # reference: std::invalid_argument(text), gemm_interface.hpp:144-165, symm_interface.hpp:51-72,
        # trsm_interface.hpp:112-128
        raise ValueError(text)
    detail = lib.pbx_last_error(h._h).decode() if h is not None and h._h else ""
    raise PbxError(f"{text}: {detail}" if detail else text)


class SB_Handle:
    """One device + one stream (reference: one sycl::queue)."""

    def __init__(self, device: int = 0, stream: Optional[object] = None):
        self._lib = _lib.load()
        self._h = ctypes.c_void_p()
        if stream is None:
            sptr = torch.cuda.current_stream(device).cuda_stream if torch.cuda.is_available() else 0
        elif isinstance(stream, int):
            sptr = stream
        else:
            sptr = stream.cuda_stream
        st = self._lib.pbx_create(ctypes.byref(self._h), int(device), ctypes.c_void_p(sptr))
        if st != _lib.OK:
            self._h = None
            raise PbxError(self._lib.pbx_status_string(st).decode())
        self.device = int(device)

    def close(self) -> None:
        if getattr(self, "_h", None):
            self._lib.pbx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- reference surface -------------------------------------------------------------------
    def wait(self) -> None:
        _check(self, self._lib.pbx_synchronize(self._h))

    @property
    def stream_ptr(self) -> int:
        """The cudaStream_t every call of this handle is ordered on.  It is fixed at construction (torch's current stream
        of the device at that moment) and changed only by ``set_stream`` -- a later ``torch.cuda.stream(...)`` context
        does not move it (the reference's SB_Handle likewise keeps the one queue it was built with)."""
        return int(self._lib.pbx_get_stream(self._h) or 0)

    def get_num_compute_units(self) -> int:
        return self._lib.pbx_get_num_compute_units(self._h)

    def set_stream(self, stream) -> None:
        sptr = stream if isinstance(stream, int) else stream.cuda_stream
        _check(self, self._lib.pbx_set_stream(self._h, ctypes.c_void_p(sptr)))

    # -- testing / tuning hooks -----------------------------------------------------------------
    def reload_env(self) -> None:
        """Re-read the PBX_* switches from the environment (they are read once when the handle is created)."""
        _check(self, self._lib.pbx_reload_env(self._h))

    def set_forced_kernel(self, kernel: int) -> None:
        _check(self, self._lib.pbx_set_forced_kernel(self._h, int(kernel)))

    def set_split_k(self, slices: int) -> None:
        _check(self, self._lib.pbx_set_split_k(self._h, int(slices)))

    def set_conj_transpose(self, enable: bool) -> None:
        """complex GEMM: False (default) = 'c' behaves like 't', as in the reference; True = BLAS conjugate."""
        _check(self, self._lib.pbx_set_conj_transpose(self._h, int(bool(enable))))

    @property
    def last_kernel(self) -> str:
        return _lib.KERNEL_NAMES[self._lib.pbx_last_kernel(self._h)]

    @property
    def last_split_k(self) -> int:
        return self._lib.pbx_last_split_k(self._h)

    @property
    def last_repack(self) -> int:
        return self._lib.pbx_last_repack(self._h)

    @property
    def last_presplit(self) -> int:
        return self._lib.pbx_last_presplit(self._h)

    @property
    def launch_count(self) -> int:
        return int(self._lib.pbx_launch_count(self._h))


_DTYPES = {
    (torch.float32, torch.float32): _lib.F32,
    (torch.float64, torch.float64): _lib.F64,
    (torch.float16, torch.float16): _lib.F16,
    (torch.float16, torch.float32): _lib.F16_F32,
    (torch.bfloat16, torch.bfloat16): _lib.BF16,
    (torch.bfloat16, torch.float32): _lib.BF16_F32,
}


def _dtype_of(a: torch.Tensor, b: torch.Tensor, c: torch.Tensor) -> int:
    if a.dtype != b.dtype:
        raise TypeError("A and B must have the same element type")
    try:
        return _DTYPES[(a.dtype, c.dtype)]
    except KeyError:
        raise TypeError(f"unsupported (in, out) element types {a.dtype}, {c.dtype}") from None


def _scalar(dtype: int, v: float):
    return ctypes.c_double(v) if dtype == _lib.F64 else ctypes.c_float(v)


def _gemm_backend(sb_handle: SB_Handle, transa: str, transb: str, m: int, n: int, k: int, alpha, a, lda,
                  stridea, b, ldb, strideb, beta, c, ldc, stridec, batch_size, batch_type) -> None:
    for t in (a, b, c):
        if not (isinstance(t, torch.Tensor) and t.is_cuda and t.is_contiguous()):
            raise TypeError("containers must be contiguous CUDA tensors (device USM analogue)")
    if c.dtype in (torch.complex64, torch.complex128):
        if batch_type != gemm_batch_type_t.strided and batch_size > 1:
            raise TypeError("complex GEMM supports strided batches only (as the reference)")
        return _gemm_complex(sb_handle, transa, transb, m, n, k, alpha, a, lda, stridea, b, ldb, strideb, beta, c,
                             ldc, stridec, batch_size)
    dt = _dtype_of(a, b, c)
    al, be = _scalar(dt, float(alpha)), _scalar(dt, float(beta))
    st = sb_handle._lib.pbx_gemm(
        sb_handle._h, dt, str(transa).encode()[:1], str(transb).encode()[:1], int(m), int(n), int(k),
        ctypes.cast(ctypes.pointer(al), ctypes.c_void_p), ctypes.c_void_p(a.data_ptr()), int(lda), int(stridea),
        ctypes.c_void_p(b.data_ptr()), int(ldb), int(strideb),
        ctypes.cast(ctypes.pointer(be), ctypes.c_void_p), ctypes.c_void_p(c.data_ptr()), int(ldc), int(stridec),
        int(batch_size), int(batch_type))
    _check(sb_handle, st)


def _gemm_complex(sb_handle, transa, transb, m, n, k, alpha, a, lda, stridea, b, ldb, strideb, beta, c, ldc, stridec,
                  batch_size) -> None:
    if not (a.dtype == b.dtype == c.dtype):
        raise TypeError("complex GEMM: A, B and C must share one element type")
    z = c.dtype == torch.complex128
    ct = ctypes.c_double if z else ctypes.c_float
    al, be = complex(alpha), complex(beta)
    al_a, be_a = (ct * 2)(al.real, al.imag), (ct * 2)(be.real, be.imag)
    fn = sb_handle._lib.pbx_zgemm if z else sb_handle._lib.pbx_cgemm
    st = fn(sb_handle._h, str(transa).encode()[:1], str(transb).encode()[:1], int(m), int(n), int(k),
            ctypes.cast(al_a, ctypes.c_void_p), ctypes.c_void_p(a.data_ptr()), int(lda), int(stridea),
            ctypes.c_void_p(b.data_ptr()), int(ldb), int(strideb), ctypes.cast(be_a, ctypes.c_void_p),
            ctypes.c_void_p(c.data_ptr()), int(ldc), int(stridec), int(batch_size))
    _check(sb_handle, st)


def _symm(sb_handle, _side, _uplo, _M, _N, _alpha, a_, _lda, b_, _ldb, _beta, _C, _ldc) -> None:
    """C <- alpha*A*B + beta*C (side 'l') or alpha*B*A + beta*C (side 'r'), A symmetric
    (blas3_interface.h:137-147 -> symm_interface.hpp:35-75)."""
    for t in (a_, b_, _C):
        if not (isinstance(t, torch.Tensor) and t.is_cuda and t.is_contiguous()):
            raise TypeError("containers must be contiguous CUDA tensors (device USM analogue)")
    dt = _dtype_of(a_, b_, _C)
    al, be = _scalar(dt, float(_alpha)), _scalar(dt, float(_beta))
    st = sb_handle._lib.pbx_symm(
        sb_handle._h, dt, str(_side).encode()[:1], str(_uplo).encode()[:1], int(_M), int(_N),
        ctypes.cast(ctypes.pointer(al), ctypes.c_void_p), ctypes.c_void_p(a_.data_ptr()), int(_lda),
        ctypes.c_void_p(b_.data_ptr()), int(_ldb), ctypes.cast(ctypes.pointer(be), ctypes.c_void_p),
        ctypes.c_void_p(_C.data_ptr()), int(_ldc))
    _check(sb_handle, st)


def _trsm(sb_handle, side, uplo, trans, diag, M, N, alpha, A, lda, B, ldb) -> None:
    """op(A)*X = alpha*B or X*op(A) = alpha*B, X overwrites B (blas3_interface.h:125-135 ->
    trsm_interface.hpp:105-387)."""
    for t in (A, B):
        if not (isinstance(t, torch.Tensor) and t.is_cuda and t.is_contiguous()):
            raise TypeError("containers must be contiguous CUDA tensors (device USM analogue)")
    dt = _dtype_of(A, B, B)
    al = _scalar(dt, float(alpha))
    st = sb_handle._lib.pbx_trsm(
        sb_handle._h, dt, str(side).encode()[:1], str(uplo).encode()[:1], str(trans).encode()[:1],
        str(diag).encode()[:1], int(M), int(N), ctypes.cast(ctypes.pointer(al), ctypes.c_void_p),
        ctypes.c_void_p(A.data_ptr()), int(lda), ctypes.c_void_p(B.data_ptr()), int(ldb))
    _check(sb_handle, st)


def _gemm(sb_handle, _TransA, _TransB, _M, _N, _K, _alpha, a_, _lda, b_, _ldb, _beta, _C, _ldc) -> None:
    """C <- alpha*op(A)*op(B) + beta*C   (gemm_interface.hpp:189-198: strides 0, batch 1)."""
    _gemm_backend(sb_handle, _TransA, _TransB, _M, _N, _K, _alpha, a_, _lda, 0, b_, _ldb, 0, _beta, _C, _ldc, 0,
                  1, gemm_batch_type_t.strided)


def _gemm_batched(sb_handle, _TransA, _TransB, _M, _N, _K, _alpha, a_, _lda, b_, _ldb, _beta, _C, _ldc,
                  batch_size, batch_type: gemm_batch_type_t = gemm_batch_type_t.strided) -> None:
    """Default strides = matrix footprints for strided; 0 for interleaved (gemm_interface.hpp:202-226)."""
    sa = sb = sc = 0
    if batch_type == gemm_batch_type_t.strided:
        sa = (_M if str(_TransA).lower() != "n" else _K) * _lda
        sb = (_K if str(_TransB).lower() != "n" else _N) * _ldb
        sc = _ldc * _N
    _gemm_backend(sb_handle, _TransA, _TransB, _M, _N, _K, _alpha, a_, _lda, sa, b_, _ldb, sb, _beta, _C, _ldc, sc,
                  batch_size, batch_type)


def _gemm_strided_batched(sb_handle, _TransA, _TransB, _M, _N, _K, _alpha, a_, _lda, _stridea, b_, _ldb,
                          _strideb, _beta, _C, _ldc, _stridec, batch_size) -> None:
    """User strides straight through (gemm_interface.hpp:230-240)."""
    _gemm_backend(sb_handle, _TransA, _TransB, _M, _N, _K, _alpha, a_, _lda, _stridea, b_, _ldb, _strideb, _beta,
                  _C, _ldc, _stridec, batch_size, gemm_batch_type_t.strided)


def gemm_host(sb_handle: SB_Handle, transa, transb, m, n, k, alpha, a_host: torch.Tensor, lda, b_host, ldb, beta,
              c_host, ldc, *, stridea=0, strideb=0, stridec=0, batch_size=1,
              batch_type=gemm_batch_type_t.strided) -> None:
    """HOST buffers in, HOST result out (copy_to_device + _gemm + copy_to_host + wait,
    reference samples/gemm.cpp:50-66).  Used for the end-to-end metric."""
    dt = _dtype_of(a_host, b_host, c_host)
    al, be = _scalar(dt, float(alpha)), _scalar(dt, float(beta))
    st = sb_handle._lib.pbx_gemm_host(
        sb_handle._h, dt, str(transa).encode()[:1], str(transb).encode()[:1], int(m), int(n), int(k),
        ctypes.cast(ctypes.pointer(al), ctypes.c_void_p), ctypes.c_void_p(a_host.data_ptr()), int(lda), int(stridea),
        ctypes.c_void_p(b_host.data_ptr()), int(ldb), int(strideb),
        ctypes.cast(ctypes.pointer(be), ctypes.c_void_p), ctypes.c_void_p(c_host.data_ptr()), int(ldc), int(stridec),
        int(batch_size), int(batch_type))
    _check(sb_handle, st)


def copy2d_to_device(sb_handle: SB_Handle, host_src: torch.Tensor, ld_src: int, dev_dst: torch.Tensor, ld_dst: int, rows: int,
                     cols: int) -> None:
    """rows x cols window of a column-major host matrix -> device (asynchronous on the handle's stream; pinned memory
    for a truly asynchronous copy)."""
    _check(sb_handle, sb_handle._lib.pbx_copy2d_to_device(
        sb_handle._h, ctypes.c_void_p(host_src.data_ptr()), int(ld_src), ctypes.c_void_p(dev_dst.data_ptr()), int(ld_dst),
        int(rows), int(cols), int(host_src.element_size())))


def copy2d_to_host(sb_handle: SB_Handle, dev_src: torch.Tensor, ld_src: int, host_dst: torch.Tensor, ld_dst: int, rows: int,
                   cols: int) -> None:
    _check(sb_handle, sb_handle._lib.pbx_copy2d_to_host(
        sb_handle._h, ctypes.c_void_p(dev_src.data_ptr()), int(ld_src), ctypes.c_void_p(host_dst.data_ptr()), int(ld_dst),
        int(rows), int(cols), int(dev_src.element_size())))


# ---- multi-GPU: the gather of C fused into the GEMM's stores (include/pbx_gemm.h: pbx_gemm_multicast) --------------
def ipc_export(sb_handle: SB_Handle, t: torch.Tensor):
    """(64-byte CUDA IPC handle, byte offset) of the allocation behind a CUDA tensor -- picklable, for the peers."""
    hbuf = ctypes.create_string_buffer(64)
    off = ctypes.c_int64(0)
    _check(sb_handle, sb_handle._lib.pbx_ipc_export(sb_handle._h, ctypes.c_void_p(t.data_ptr()), hbuf, ctypes.byref(off)))
    return bytes(hbuf.raw), int(off.value)


def ipc_import(sb_handle: SB_Handle, exported) -> int:
    """Device pointer (int) in THIS process for a peer's exported tensor."""
    hbytes, off = exported
    hbuf = ctypes.create_string_buffer(hbytes, 64)
    out = ctypes.c_void_p()
    _check(sb_handle, sb_handle._lib.pbx_ipc_import(sb_handle._h, hbuf, ctypes.c_int64(off), ctypes.byref(out)))
    return int(out.value)


def _gemm_multicast(sb_handle: SB_Handle, _TransA, _TransB, _M, _N, _K, _alpha, a_, _lda, b_, _ldb, _beta, c_ptrs,
                    _ldc, c_dtype: torch.dtype) -> None:
    """``_gemm`` whose result is written to every pointer of ``c_ptrs`` (ints; [0] = this GPU's C, the rest are peer
    copies from ``ipc_import``), all with leading dimension ``_ldc``."""
    dt = _DTYPES[(a_.dtype, c_dtype)]
    al, be = _scalar(dt, float(_alpha)), _scalar(dt, float(_beta))
    arr = (ctypes.c_void_p * len(c_ptrs))(*[ctypes.c_void_p(int(p)) for p in c_ptrs])
    st = sb_handle._lib.pbx_gemm_multicast(
        sb_handle._h, dt, str(_TransA).encode()[:1], str(_TransB).encode()[:1], int(_M), int(_N), int(_K),
        ctypes.cast(ctypes.pointer(al), ctypes.c_void_p), ctypes.c_void_p(a_.data_ptr()), int(_lda),
        ctypes.c_void_p(b_.data_ptr()), int(_ldb), ctypes.cast(ctypes.pointer(be), ctypes.c_void_p), arr, len(c_ptrs),
        int(_ldc))
    _check(sb_handle, st)


# ---- multi-GPU from one process: include/pbx_gemm.h pbx_multi_* / pbx_gemm_sharded* (C++ mirror: blas::multi) -------
class SB_Handle_Group:
    """A group of devices with peer access, one handle and one stream each (no reference counterpart: a
    blas::SB_Handle is one sycl::queue, include/sb_handle/portblas_handle.h:51-60).  ``devices`` may repeat an ordinal
    (several shards on one GPU)."""

    def __init__(self, devices):
        self._lib = _lib.load()
        self._mh = ctypes.c_void_p()
        devs = [int(d) for d in devices]
        arr = (ctypes.c_int * len(devs))(*devs)
        st = self._lib.pbx_multi_create(ctypes.byref(self._mh), len(devs), arr)
        if st != _lib.OK:
            self._mh = None
            raise PbxError(self._lib.pbx_status_string(st).decode())
        self.devices = devs

    def close(self) -> None:
        if getattr(self, "_mh", None):
            self._lib.pbx_multi_destroy(self._mh)
            self._mh = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __len__(self) -> int:
        return self._lib.pbx_multi_device_count(self._mh)

    def wait(self) -> None:
        self._check(self._lib.pbx_multi_synchronize(self._mh))

    def _check(self, status: int) -> None:
        if status == _lib.OK:
            return
        text = self._lib.pbx_status_string(status).decode()
        if 1 <= status <= 5:
            raise ValueError(text)
        raise PbxError(f"{text}: {self._lib.pbx_multi_last_error(self._mh).decode()}")

    def _ptrs(self, tensors):
        return (ctypes.c_void_p * len(tensors))(*[ctypes.c_void_p(t.data_ptr()) for t in tensors])

    def gemm_sharded(self, transa, transb, m, n, k, alpha, a_blocks, lda, b_full, ldb, beta, c_full, ldc, gather=True) -> None:
        """M-block sharded ``_gemm``: ``a_blocks[g]`` = device g's rows ``sharding.split_range(m, G, g, 256)`` of op(A),
        ``b_full[g]`` = all of B on device g, ``c_full[g]`` = device g's m x n C.  Asynchronous (``wait()``)."""
        dt = _dtype_of(a_blocks[0], b_full[0], c_full[0])
        al, be = _scalar(dt, float(alpha)), _scalar(dt, float(beta))
        self._check(self._lib.pbx_gemm_sharded(
            self._mh, dt, str(transa).encode()[:1], str(transb).encode()[:1], int(m), int(n), int(k),
            ctypes.cast(ctypes.pointer(al), ctypes.c_void_p), self._ptrs(a_blocks), int(lda), self._ptrs(b_full), int(ldb),
            ctypes.cast(ctypes.pointer(be), ctypes.c_void_p), self._ptrs(c_full), int(ldc), int(bool(gather))))

    def gemm_strided_batched_sharded(self, transa, transb, m, n, k, alpha, a_shards, lda, stridea, b_shards, ldb, strideb, beta,
                                     c_shards, ldc, stridec, batch_size) -> None:
        """Batch-range sharded ``_gemm_strided_batched``: ``x_shards[g]`` starts at the first entry device g owns."""
        dt = _dtype_of(a_shards[0], b_shards[0], c_shards[0])
        al, be = _scalar(dt, float(alpha)), _scalar(dt, float(beta))
        self._check(self._lib.pbx_gemm_strided_batched_sharded(
            self._mh, dt, str(transa).encode()[:1], str(transb).encode()[:1], int(m), int(n), int(k),
            ctypes.cast(ctypes.pointer(al), ctypes.c_void_p), self._ptrs(a_shards), int(lda), int(stridea), self._ptrs(b_shards),
            int(ldb), int(strideb), ctypes.cast(ctypes.pointer(be), ctypes.c_void_p), self._ptrs(c_shards), int(ldc),
            int(stridec), int(batch_size)))

    def gemm_sharded_host(self, transa, transb, m, n, k, alpha, a_host, lda, b_host, ldb, beta, c_host, ldc) -> None:
        """``_gemm`` on HOST tensors over the whole group (synchronous)."""
        dt = _dtype_of(a_host, b_host, c_host)
        al, be = _scalar(dt, float(alpha)), _scalar(dt, float(beta))
        self._check(self._lib.pbx_gemm_sharded_host(
            self._mh, dt, str(transa).encode()[:1], str(transb).encode()[:1], int(m), int(n), int(k),
            ctypes.cast(ctypes.pointer(al), ctypes.c_void_p), ctypes.c_void_p(a_host.data_ptr()), int(lda),
            ctypes.c_void_p(b_host.data_ptr()), int(ldb), ctypes.cast(ctypes.pointer(be), ctypes.c_void_p),
            ctypes.c_void_p(c_host.data_ptr()), int(ldc)))
