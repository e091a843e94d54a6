"""GPU tests of pbx_gemm_multicast: the GEMM whose epilogue stores every tile into several copies of C -- the fused
form of "M-block sharded GEMM + gather of C" (SURVEY.md section 8e).

* one GPU: two local copies; all transposes, ragged shapes, beta != 0, every tcgen05 dtype, and the kernels that
  fall back to compute-then-copy (fp64 DMMA, tiny SIMT shapes);
* two GPUs (skipped on a one-GPU box): one process per GPU, full C exported over CUDA IPC, each rank multicasts its
  row block into both C's; afterwards every rank must hold the whole product.
"""
from __future__ import annotations

import itertools
import os
import socket

import pytest
import torch

from portblas_b200 import blas

pytestmark = pytest.mark.gpu

TOL = {torch.float64: 1e-12, torch.float32: 1e-5, torch.float16: 2e-3, torch.bfloat16: 1.6e-2}


def _ref(a, b, c0, ta, tb, m, n, k, lda, ldb, ldc, alpha, beta):
    f64 = torch.float64
    A = (a.view(-1, lda).T[:m] if ta == "n" else a.view(-1, lda)[:m, :k]).to(f64)[:, :k]
    B = (b.view(-1, ldb).T[:k] if tb == "n" else b.view(-1, ldb)[:k, :n]).to(f64)[:, :n]
    C0 = c0.view(n, ldc).T[:m].to(f64)
    return alpha * (A @ B) + beta * C0, abs(alpha) * (A.abs() @ B.abs()) + abs(beta) * C0.abs()


@pytest.mark.parametrize("tin,tout", [(torch.float32, torch.float32), (torch.bfloat16, torch.float32),
                                      (torch.float16, torch.float16), (torch.bfloat16, torch.bfloat16),
                                      (torch.float64, torch.float64)])
def test_multicast_two_local_copies(handle, tin, tout):
    gen = torch.Generator(device="cuda").manual_seed(21)
    shapes = [(392, 520, 264), (1000, 136, 1096), (128, 128, 64), (7, 5, 9)]
    for (m, n, k), (ta, tb), beta in itertools.product(shapes, [("n", "n"), ("t", "n"), ("n", "t"), ("t", "t")], [0.0, 0.5]):
        lda, ldb, ldc = (m if ta == "n" else k) + 8, (k if tb == "n" else n) + 8, m + 8
        a = (torch.rand(lda * (k if ta == "n" else m), device="cuda", generator=gen) * 7 - 2).to(tin)
        b = (torch.rand(ldb * (n if tb == "n" else k), device="cuda", generator=gen) * 7 - 2).to(tin)
        c0 = (torch.rand(ldc * n, device="cuda", generator=gen) * 7 - 2).to(tout)
        c1, c2 = c0.clone(), torch.full_like(c0, 77.0)
        blas._gemm_multicast(handle, ta, tb, m, n, k, 1.5, a, lda, b, ldb, beta, [c1.data_ptr(), c2.data_ptr()], ldc, tout)
        handle.wait()
        want, bound = _ref(a, b, c0, ta, tb, m, n, k, lda, ldb, ldc, 1.5, beta)
        g1, g2 = c1.view(n, ldc).T, c2.view(n, ldc).T
        rel = float(((g1[:m].double() - want).abs() / bound).max())
        assert rel <= TOL[tout], f"{tin} {ta}{tb} {m}x{n}x{k} beta={beta}: {rel:.2e} kernel={handle.last_kernel}"
        assert torch.equal(g1[:m], g2[:m]), "the second copy differs from the first"
        assert torch.equal(g1[m:], c0.view(n, ldc).T[m:]) and bool((g2[m:] == 77.0).all()), "ld padding was written"
    if tin != torch.float64:
        assert handle.last_kernel in ("tcgen05", "simt")


@pytest.mark.parametrize("tin,tout", [(torch.float32, torch.float32), (torch.bfloat16, torch.bfloat16)])
def test_multicast_copies_that_tma_cannot_address(handle, tin, tout):
    """A leading dimension of C that is not a multiple of 16 bytes rules out the pusher's (and the 16-bit epilogue's)
    tensor maps: the epilogue warps then store every tile to all copies themselves (round 1's form), three copies here."""
    gen = torch.Generator(device="cuda").manual_seed(22)
    for (m, n, k), (ta, tb), beta in itertools.product([(392, 264, 520), (520, 136, 264)], [("n", "n"), ("t", "t")], [0.0, 0.5]):
        lda, ldb, ldc = (m if ta == "n" else k), (k if tb == "n" else n), m + 9
        a = (torch.rand(lda * (k if ta == "n" else m), device="cuda", generator=gen) * 7 - 2).to(tin)
        b = (torch.rand(ldb * (n if tb == "n" else k), device="cuda", generator=gen) * 7 - 2).to(tin)
        c0 = (torch.rand(ldc * n, device="cuda", generator=gen) * 7 - 2).to(tout)
        cs = [c0.clone(), torch.full_like(c0, 77.0), torch.full_like(c0, -5.0)]
        blas._gemm_multicast(handle, ta, tb, m, n, k, 1.5, a, lda, b, ldb, beta, [c.data_ptr() for c in cs], ldc, tout)
        handle.wait()
        assert handle.last_kernel == "tcgen05"
        want, bound = _ref(a, b, c0, ta, tb, m, n, k, lda, ldb, ldc, 1.5, beta)
        g = [c.view(n, ldc).T for c in cs]
        rel = float(((g[0][:m].double() - want).abs() / bound).max())
        assert rel <= TOL[tout], f"{tin} {ta}{tb} {m}x{n}x{k} beta={beta}: {rel:.2e}"
        assert torch.equal(g[0][:m], g[1][:m]) and torch.equal(g[0][:m], g[2][:m])
        assert bool((g[1][m:] == 77.0).all()) and bool((g[2][m:] == -5.0).all()), "ld padding was written"


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch.distributed as dist
    from portblas_b200 import SB_Handle, sharding
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.cuda.set_device(rank)
        dev = torch.device("cuda", rank)
        h = SB_Handle(rank)
        for tdt, (ta, tb), beta in itertools.product([torch.float32, torch.float64, torch.bfloat16], [("n", "n"), ("t", "t")],
                                                     [0.0, 0.5]):
            m, n, k = 1024 * world, 1536, 2048
            gen = torch.Generator(device=dev).manual_seed(5)       # same seed: every rank generates the same operands
            a = (torch.rand(m * k, device=dev, generator=gen) * 7 - 2).to(tdt)
            b = (torch.rand(k * n, device=dev, generator=gen) * 7 - 2).to(tdt)
            c0 = (torch.rand(m * n, device=dev, generator=gen) * 7 - 2).to(tdt)
            lda, ldb = (m if ta == "n" else k), (k if tb == "n" else n)
            c_full = c0.clone()
            ptrs = sharding.share_full_c(h, c_full)
            torch.cuda.synchronize()
            dist.barrier()
            sh = sharding.shard_mblock(ta, m, lda, world, rank, align=128)
            sharding.gemm_mblock_fused_gather(h, ta, tb, m, n, k, 1.5, a[sh.a_offset:], lda, b, ldb, beta, ptrs, m, tdt,
                                              world, rank)
            h.wait()
            dist.barrier()          # every rank's stores (local and remote) are complete
            torch.cuda.synchronize()
            want, bound = _ref(a, b, c0, ta, tb, m, n, k, lda, ldb, m, 1.5, beta)
            rel = float(((c_full.view(n, m).T.double() - want).abs() / bound).max())
            assert rel <= TOL[tdt], f"rank {rank} {tdt} {ta}{tb} beta={beta}: {rel:.2e} kernel={h.last_kernel}"
            dist.barrier()
        # HOST operands: every rank uploads its rows of A and one panel of B, the panels are exchanged over NVLink
        # (an NCCL group beside the test's gloo one), each rank fills its rows of the host C
        pg = dist.new_group(backend="nccl")   # the B panels travel GPU to GPU
        for tdt, ta, beta in itertools.product([torch.float32, torch.float64], ["n", "t"], [0.0, 0.5]):
            m, n, k = 512 * world, 1024, 520
            gen = torch.Generator().manual_seed(9)
            a_h = (torch.rand(m * k, generator=gen) * 7 - 2).to(tdt).pin_memory()
            b_h = (torch.rand(k * n, generator=gen) * 7 - 2).to(tdt).pin_memory()
            c0_h = (torch.rand(m * n, generator=gen) * 7 - 2).to(tdt)
            c_h = c0_h.clone().pin_memory()
            lda = m if ta == "n" else k
            sh = sharding.gemm_mblock_host(h, ta, "n", m, n, k, 1.5, a_h, lda, b_h, k, beta, c_h, m, world, rank, group=pg)
            want, bound = _ref(a_h, b_h, c0_h, ta, "n", m, n, k, lda, k, m, 1.5, beta)
            got = c_h.view(n, m).T.double()
            rows = slice(sh.row0, sh.row0 + sh.rows)
            rel = float(((got[rows] - want[rows]).abs() / bound[rows]).max())
            assert rel <= TOL[tdt], f"rank {rank} host path {tdt} {ta} beta={beta}: {rel:.2e}"
            other = torch.ones(m, dtype=torch.bool)
            other[rows] = False
            assert torch.equal(got[other], c0_h.view(n, m).T.double()[other]), "rows of other ranks were written"
            dist.barrier()
        h.close()
        q.put((rank, "ok"))
    except Exception as e:  # noqa: BLE001
        import traceback
        q.put((rank, f"{type(e).__name__}: {e}\n{traceback.format_exc()}"))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_multicast_two_gpus_every_rank_holds_full_c():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs on one box (run with gpurun --gpus 2)")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=280) for _ in procs]
    for p in procs:
        p.join(timeout=30)
    assert all(r[1] == "ok" for r in res), res
