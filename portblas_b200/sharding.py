"""Multi-GPU partitioning of the GEMM path: one process per GPU, no data-path collective.

The reference is single-device (one SB_Handle == one sycl::queue,
include/sb_handle/portblas_handle.h:51-60); sharding is new in this build:

  * large GEMMs are split into M-blocks: rank g owns rows [r0, r0+rows) of op(A) and of C, B is
    replicated.  In column-major storage a row block is just a pointer offset with the ORIGINAL
    leading dimensions, so each rank runs an ordinary (rows x N x K) ``_gemm``;
  * strided-batched GEMMs are split by batch range: pointer offsets b0*stride, local batch count.

Both partitions are embarrassingly parallel.  The only collective is the optional gather of C
into one buffer (``gather_c_mblocks`` / ``gather_c_batches``), NCCL all_gather over NVLink on
GPUs (gloo on CPU in the tests).  ``gemm_mblock_gather_overlapped`` hides that gather behind the
compute: the local row block is produced in column panels and panel j travels over NVLink (on a
side stream) while panel j+1 is still on the tensor cores.
"""
from __future__ import annotations

import dataclasses
from typing import List, Tuple

import torch
import torch.distributed as dist


def split_range(total: int, world: int, rank: int, align: int = 1) -> Tuple[int, int]:
    """[start, count) of ``rank``'s share of ``total`` units; shares are multiples of ``align``
    except possibly the last non-empty one; earlier ranks take the remainder.  A binding of the C-ABI's
    ``pbx_shard_range`` (portblas_b200/csrc/pbx_multi.cu), so that the one-process-per-GPU path here and the
    single-process ``pbx_gemm_sharded`` cut every problem identically."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad world/rank")
    import ctypes
    from . import _lib
    start, count = ctypes.c_int64(0), ctypes.c_int64(0)
    st = _lib.load().pbx_shard_range(int(total), int(world), int(rank), int(align), ctypes.byref(start), ctypes.byref(count))
    if st != _lib.OK:
        raise ValueError("pbx_shard_range: invalid argument")
    return int(start.value), int(count.value)


@dataclasses.dataclass(frozen=True)
class MBlockShard:
    row0: int        # first row of C / op(A) owned by this rank
    rows: int        # local M
    a_offset: int    # element offset into A
    c_offset: int    # element offset into C


def shard_mblock(transa: str, m: int, lda: int, world: int, rank: int, align: int = 128) -> MBlockShard:
    """Rows of op(A): A is stored M x K (rows contiguous) for 'n' -> offset row0;
    stored K x M for 't' -> offset row0*lda."""
    row0, rows = split_range(m, world, rank, align)
    a_off = row0 if transa.lower() == "n" else row0 * lda
    return MBlockShard(row0, rows, a_off, row0)


@dataclasses.dataclass(frozen=True)
class BatchShard:
    batch0: int
    batches: int
    a_offset: int
    b_offset: int
    c_offset: int


def shard_batch(batch: int, stridea: int, strideb: int, stridec: int, world: int, rank: int) -> BatchShard:
    b0, cnt = split_range(batch, world, rank)
    return BatchShard(b0, cnt, b0 * stridea, b0 * strideb, b0 * stridec)


def gather_c_mblocks(c_local: torch.Tensor, m: int, n: int, world: int, align: int = 128, group=None) -> torch.Tensor:
    """All-gather compact (rows x n, column-major, ld == rows) C row-blocks into a full compact
    m x n column-major matrix on every rank."""
    shapes = [split_range(m, world, r, align) for r in range(world)]
    max_rows = max(s[1] for s in shapes)
    if all(s[1] == max_rows for s in shapes) and hasattr(dist, "all_gather_into_tensor"):
        # equal blocks: one all-gather straight into a [world][n][rows] staging tensor, then one strided copy
        stage = torch.empty(world * n * max_rows, dtype=c_local.dtype, device=c_local.device)
        dist.all_gather_into_tensor(stage, c_local[:n * max_rows].contiguous(), group=group)
        full = torch.empty(m * n, dtype=c_local.dtype, device=c_local.device)
        full.view(n, world, max_rows).copy_(stage.view(world, n, max_rows).permute(1, 0, 2))
        return full
    pad = torch.zeros(max_rows * n, dtype=c_local.dtype, device=c_local.device)
    rows = shapes[dist.get_rank(group)][1]
    pad.view(n, max_rows)[:, :rows] = c_local.view(n, rows)
    out: List[torch.Tensor] = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad, group=group)
    full = torch.empty(m * n, dtype=c_local.dtype, device=c_local.device)
    fv = full.view(n, m)
    for r, (r0, cnt) in enumerate(shapes):
        if cnt:
            fv[:, r0:r0 + cnt] = out[r].view(n, max_rows)[:, :cnt]
    return full


def gather_c_batches(c_local: torch.Tensor, per_matrix: int, batch: int, world: int, group=None) -> torch.Tensor:
    """All-gather contiguous batch ranges of C (stride_c == per_matrix)."""
    shapes = [split_range(batch, world, r) for r in range(world)]
    max_b = max(s[1] for s in shapes)
    if all(s[1] == max_b for s in shapes) and hasattr(dist, "all_gather_into_tensor"):
        # equal batch ranges: the concatenation of the shards IS the full buffer -- one all-gather, no copies
        full = torch.empty(batch * per_matrix, dtype=c_local.dtype, device=c_local.device)
        dist.all_gather_into_tensor(full, c_local[:max_b * per_matrix].contiguous(), group=group)
        return full
    pad = torch.zeros(max_b * per_matrix, dtype=c_local.dtype, device=c_local.device)
    cnt = shapes[dist.get_rank(group)][1]
    pad[:cnt * per_matrix] = c_local[:cnt * per_matrix]
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad, group=group)
    return torch.cat([out[r][:shapes[r][1] * per_matrix] for r in range(world)])


def panel_ranges(n: int, panels: int, align: int = 256) -> List[Tuple[int, int]]:
    """[(n0, nb)] column panels of about n/panels columns, multiples of ``align`` except the last."""
    per = max(align, ((n + panels - 1) // panels + align - 1) // align * align)
    return [(n0, min(per, n - n0)) for n0 in range(0, n, per)]


def gemm_mblock_gather_overlapped(gemm_panel, m: int, n: int, rows: int, world: int, dtype, device, panels: int = 8,
                                  group=None, side_stream=None) -> torch.Tensor:
    """M-block sharded GEMM whose C gather overlaps the compute.

    ``gemm_panel(n0, nb, c_panel)`` must compute columns [n0, n0+nb) of this rank's ``rows`` x n row block into the
    compact column-major tensor ``c_panel`` (rows*nb elements, ld == rows) -- on the GPU an ordinary ``blas._gemm`` on
    the handle's stream.  Every rank owns the same number of rows (m == rows*world).  After panel j is enqueued its
    all-gather is issued on ``side_stream`` (ordered after the panel's GEMM by an event), so the NVLink transfer of
    panel j runs under the GEMM of panel j+1; the gathered panels land directly in their final place of the full
    compact m x n matrix, which is returned on every rank."""
    if rows * world != m:
        raise ValueError("gemm_mblock_gather_overlapped needs equal row blocks (m == rows * world)")
    cuda = torch.device(device).type == "cuda"
    c_loc = torch.empty(rows * n, dtype=dtype, device=device)
    full = torch.empty(m * n, dtype=dtype, device=device)
    stages = []
    main = torch.cuda.current_stream(device) if cuda else None
    if cuda and side_stream is None:
        side_stream = torch.cuda.Stream(device=device)
    for n0, nb in panel_ranges(n, panels):
        c_panel = c_loc[n0 * rows:(n0 + nb) * rows]
        gemm_panel(n0, nb, c_panel)
        stage = torch.empty(world * nb * rows, dtype=dtype, device=device)   # [world][nb][rows]
        stages.append(stage)
        if cuda:
            ev = torch.cuda.Event()
            ev.record(main)
            side_stream.wait_event(ev)
            with torch.cuda.stream(side_stream):
                dist.all_gather_into_tensor(stage, c_panel, group=group)
                full.view(n, world, rows)[n0:n0 + nb].copy_(stage.view(world, nb, rows).permute(1, 0, 2))
        else:
            dist.all_gather_into_tensor(stage, c_panel, group=group)
            full.view(n, world, rows)[n0:n0 + nb].copy_(stage.view(world, nb, rows).permute(1, 0, 2))
    if cuda:
        main.wait_stream(side_stream)
    return full


def share_full_c(sb_handle, c_full: torch.Tensor, group=None) -> List[int]:
    """Exchange CUDA IPC handles of every rank's full C and map the peers' buffers: returns device pointers
    (ints) indexed by rank, valid in this process ([rank] is the local tensor itself)."""
    from . import blas
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    exported = [None] * world
    dist.all_gather_object(exported, blas.ipc_export(sb_handle, c_full), group=group)
    return [c_full.data_ptr() if r == rank else blas.ipc_import(sb_handle, exported[r]) for r in range(world)]


def gemm_mblock_fused_gather(sb_handle, transa: str, transb: str, m: int, n: int, k: int, alpha, a_local, lda: int, b,
                             ldb: int, beta, c_ptrs: List[int], ldc: int, c_dtype, world: int, rank: int,
                             align: int = 128) -> MBlockShard:
    """This rank's M-block of C <- alpha*op(A)*op(B) + beta*C, written by the GEMM epilogue into the same rows of EVERY
    rank's full C (``c_ptrs`` from ``share_full_c``): when all ranks have run it and synchronised, each holds the whole
    result -- the gather costs no pass over C and no collective.  ``a_local`` already points at the rank's rows of A
    (``shard_mblock(...).a_offset``)."""
    from . import blas
    sh = shard_mblock(transa, m, lda, world, rank, align)
    es = torch.empty(0, dtype=c_dtype).element_size()
    ptrs = [c_ptrs[rank] + sh.c_offset * es] + [c_ptrs[r] + sh.c_offset * es for r in range(world) if r != rank]
    blas._gemm_multicast(sb_handle, transa, transb, sh.rows, n, k, alpha, a_local, lda, b, ldb, beta, ptrs, ldc, c_dtype)
    return sh


def gemm_mblock_host(sb_handle, transa: str, transb: str, m: int, n: int, k: int, alpha, a_host: torch.Tensor, lda: int,
                     b_host: torch.Tensor, ldb: int, beta, c_host: torch.Tensor, ldc: int, world: int, rank: int,
                     group=None, align: int = 256) -> MBlockShard:
    """``_gemm`` on HOST operands, M-block sharded, one process per GPU -- the torch.distributed twin of the C-ABI's
    ``pbx_gemm_sharded_host``.  Every rank uploads its rows of op(A) and ONE column panel of op(B) over its own PCIe
    link; the panels are exchanged over NVLink (the path's one real exchange step: an NCCL all-gather of B), so B
    crosses PCIe once instead of ``world`` times; then the rank computes its row block and downloads it into its rows
    of ``c_host``.  Synchronous.  Needs equal panels (n a multiple of align * world) and 'n' for transb; other calls fall
    back to uploading all of B per rank."""
    from . import blas
    sh = shard_mblock(transa, m, lda, world, rank, align)
    dev = torch.device("cuda", sb_handle.device)
    ta, tb = transa.lower() != "n", transb.lower() != "n"
    dt_in, dt_out = a_host.dtype, c_host.dtype
    rows = sh.rows
    # The handle's stream carries the copies and the GEMM; torch's current stream carries the NCCL exchange.
    hs = torch.cuda.ExternalStream(sb_handle.stream_ptr, device=dev)
    cur = torch.cuda.current_stream(dev)
    # ---- this rank's rows of op(A): a window of the host matrix -> compact device copy (2-D DMA, no host-side repack) ----
    a_rows, a_cols = (k, rows) if ta else (rows, k)
    a_dev = torch.empty(max(a_rows * a_cols, 1), dtype=dt_in, device=dev)
    lda_d = max(a_rows, 1)
    blas.copy2d_to_device(sb_handle, a_host[sh.a_offset:], lda, a_dev, lda_d, a_rows, a_cols)
    # ---- B: one panel per rank over PCIe, the rest over NVLink ----
    n0, nn = split_range(n, world, rank, align)
    equal = (not tb) and nn * world == n and world > 1
    if equal:
        b_dev = torch.empty(n * k, dtype=dt_in, device=dev)
        panel = b_dev[n0 * k:(n0 + nn) * k]              # the panel lands in its final place
        blas.copy2d_to_device(sb_handle, b_host[n0 * ldb:], ldb, panel, k, k, nn)
        cur.wait_stream(hs)
        dist.all_gather_into_tensor(b_dev, panel, group=group)
        hs.wait_stream(cur)
        ldb_d = k
    else:
        b_rows, b_cols = (n, k) if tb else (k, n)
        b_dev = torch.empty(b_rows * b_cols, dtype=dt_in, device=dev)
        blas.copy2d_to_device(sb_handle, b_host, ldb, b_dev, b_rows, b_rows, b_cols)
        ldb_d = b_rows
    c_dev = torch.empty(n * max(rows, 1), dtype=dt_out, device=dev)
    if rows > 0:
        if float(beta) != 0.0:
            blas.copy2d_to_device(sb_handle, c_host[sh.c_offset:], ldc, c_dev, rows, rows, n)
        blas._gemm(sb_handle, transa, transb, rows, n, k, alpha, a_dev, lda_d, b_dev, ldb_d, beta, c_dev, rows)
        blas.copy2d_to_host(sb_handle, c_dev, rows, c_host[sh.c_offset:], ldc, rows, n)
    sb_handle.wait()
    return sh
