"""Host-side multi-GPU logic on CPU: partition arithmetic, and a world_size-2 gloo run in which
each rank computes its M-block / batch range (numpy stands in for the device GEMM) and the C
gather reproduces the unsharded result."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from portblas_b200 import sharding  # noqa: E402


def test_split_range_covers_everything():
    for total, world, align in [(16384, 8, 128), (8192, 3, 128), (1000, 4, 128), (5, 8, 1), (4096, 8, 1), (0, 2, 1)]:
        pos = 0
        for r in range(world):
            s, c = sharding.split_range(total, world, r, align)
            assert s == pos and c >= 0
            if c and s + c < total:
                assert c % align == 0
            pos += c
        assert pos == total


def test_split_range_properties():
    """pbx_shard_range over random problems: the parts tile [0, total) in rank order, every boundary but the end is a
    multiple of `align`, and no two parts differ by more than one aligned unit (the balance bench.py's strong scaling
    relies on); invalid arguments are rejected."""
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=300, deadline=None)
    @given(st.integers(0, 1 << 40), st.integers(1, 16), st.sampled_from([1, 2, 7, 128, 256, 1000]))
    def check(total, world, align):
        parts = [sharding.split_range(total, world, r, align) for r in range(world)]
        pos = 0
        for s, c in parts:
            assert s == pos and c >= 0
            pos += c
            assert pos == total or pos % align == 0
        assert pos == total
        units = [-(-c // align) for _, c in parts]
        assert max(units) - min(units) <= 1
        assert sorted(units, reverse=True) == units          # earlier ranks take the remainder
    check()
    for bad in [(-1, 2, 0, 1), (10, 0, 0, 1), (10, 2, 2, 1), (10, 2, -1, 1), (10, 2, 0, 0)]:
        with pytest.raises(ValueError):
            sharding.split_range(*bad)


def test_shard_offsets():
    sh = sharding.shard_mblock("n", 16384, 16384, 8, 3)
    assert (sh.row0, sh.rows, sh.a_offset, sh.c_offset) == (6144, 2048, 6144, 6144)
    sh = sharding.shard_mblock("t", 16384, 20000, 8, 3)
    assert sh.a_offset == 6144 * 20000
    bs = sharding.shard_batch(4096, 65536, 0, 65536, 8, 7)
    assert (bs.batch0, bs.batches, bs.a_offset, bs.b_offset, bs.c_offset) == (3584, 512, 3584 * 65536, 0, 3584 * 65536)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(7)
        m, n, k = 300, 40, 50
        for transa in ("n", "t"):
            lda = (k if transa == "t" else m) + 3
            A = rng.uniform(-2, 5, lda * (m if transa == "t" else k))
            B = rng.uniform(-2, 5, k * n)
            sh = sharding.shard_mblock(transa, m, lda, world, rank, align=128)
            a_loc = A[sh.a_offset:]
            if transa == "n":
                opa = np.lib.stride_tricks.as_strided(a_loc, (sh.rows, k), (8, lda * 8))
            else:
                opa = np.lib.stride_tricks.as_strided(a_loc, (k, sh.rows), (8, lda * 8)).T
            c_loc = (opa @ B.reshape(n, k).T).T.copy().reshape(-1)  # compact rows x n column-major
            full = sharding.gather_c_mblocks(torch.from_numpy(c_loc), m, n, world, align=128).numpy()
            if transa == "n":
                opa_full = np.lib.stride_tricks.as_strided(A, (m, k), (8, lda * 8))
            else:
                opa_full = np.lib.stride_tricks.as_strided(A, (k, m), (8, lda * 8)).T
            want = (opa_full @ B.reshape(n, k).T).T.reshape(-1)
            assert np.allclose(full, want, rtol=1e-12, atol=1e-12)
        # panel-overlapped gather: numpy stands in for the device GEMM of each column panel
        m2, n2, k2 = 64 * world, 1000, 37
        A2, B2 = rng.uniform(-2, 5, m2 * k2), rng.uniform(-2, 5, k2 * n2)
        rows2 = m2 // world
        opa2 = A2.reshape(k2, m2).T[rank * rows2:(rank + 1) * rows2]
        opb2 = B2.reshape(n2, k2).T

        def gemm_panel(n0, nb, c_panel):
            c_panel.copy_(torch.from_numpy((opa2 @ opb2[:, n0:n0 + nb]).T.copy().reshape(-1)))
        full2 = sharding.gemm_mblock_gather_overlapped(gemm_panel, m2, n2, rows2, world, torch.float64, "cpu",
                                                       panels=3).numpy()
        want2 = (A2.reshape(k2, m2).T @ opb2).T.reshape(-1)
        assert np.allclose(full2, want2, rtol=1e-12, atol=1e-12)
        assert sharding.panel_ranges(1000, 3) == [(0, 512), (512, 488)]
        # batch sharding
        batch, per = 7, 12
        Cb = rng.uniform(-2, 5, batch * per)
        bs = sharding.shard_batch(batch, per, per, per, world, rank)
        loc = torch.from_numpy(Cb[bs.c_offset: bs.c_offset + bs.batches * per].copy())
        got = sharding.gather_c_batches(loc, per, batch, world).numpy()
        assert np.array_equal(got, Cb)
        q.put((rank, "ok"))
    except Exception as e:  # noqa: BLE001
        q.put((rank, f"{type(e).__name__}: {e}"))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_gloo_world2_gather_reproduces_unsharded():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=100) for _ in procs]
    for p in procs:
        p.join(timeout=30)
    assert all(r[1] == "ok" for r in res), res
