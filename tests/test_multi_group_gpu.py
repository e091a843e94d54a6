"""The C-ABI's multi-GPU entry points (include/pbx_gemm.h: pbx_multi_*, pbx_gemm_sharded, pbx_gemm_strided_batched_sharded,
pbx_gemm_sharded_host) and their C++ mirror blas::multi (include/interface/blas3_interface_multi.h).

The reference is single-device (include/sb_handle/portblas_handle.h:51-60), so the oracle here is the single-GPU
product itself and an fp64 torch product: a sharded call must give what one `_gemm` gives.  A group may name the same
device several times, so the partition, the peer stores of the fused gather and the B exchange of the host path all run
on a one-GPU box; with two or more GPUs the same tests also run across real peers.
"""
from __future__ import annotations

import subprocess
from pathlib import Path

import pytest
import torch

from portblas_b200 import SB_Handle_Group, sharding

ROOT = Path(__file__).resolve().parents[1]
pytestmark = pytest.mark.gpu


def _groups():
    n = torch.cuda.device_count()
    gs = [[0, 0], [0, 0, 0]]
    if n >= 2:
        gs.append(list(range(min(n, 8))))
    return gs


def _rand(shape, dt, dev, seed):
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    return (torch.rand(*shape, device=dev, generator=g) * 7 - 2).to(dt)


@pytest.mark.parametrize("dt,tol", [(torch.float32, 1e-5), (torch.bfloat16, 1.6e-2), (torch.float64, 1e-12)])
@pytest.mark.parametrize("ta,tb", [("n", "n"), ("t", "t")])
@pytest.mark.parametrize("gather", [True, False])
def test_gemm_sharded_matches_fp64_product(handle, dt, tol, ta, tb, gather):
    m, n, k = 1000, 392, 520      # 1000 rows: M-blocks 512 + 488 (two shards) / 512 + 256 + 232 (three): ragged last block
    for devs in _groups():
        grp = SB_Handle_Group(devs)
        G = len(devs)
        a = _rand((k, m) if ta == "n" else (m, k), dt, "cuda:0", 1)        # [col][row]: column-major with ld = rows
        b = _rand((n, k) if tb == "n" else (k, n), dt, "cuda:0", 2)
        opa = a.double().T if ta == "n" else a.double()
        opb = b.double().T if tb == "n" else b.double()
        want = 1.5 * (opa @ opb)
        bound = 1.5 * (opa.abs() @ opb.abs())
        lda = a.shape[1]
        a_blocks, b_full, c_full, ranges = [], [], [], []
        for g, d in enumerate(devs):
            r0, rows = sharding.split_range(m, G, g, 256)
            ranges.append((r0, rows))
            # device g's rows of op(A), as a pointer offset into a full copy of A with the ORIGINAL lda
            a_g = a.to(f"cuda:{d}").contiguous().view(-1)
            a_blocks.append(a_g[(r0 if ta == "n" else r0 * lda):] if rows else a_g)
            b_full.append(b.to(f"cuda:{d}").contiguous().view(-1))
            c_full.append(torch.full((n * m,), float("nan"), device=f"cuda:{d}", dtype=dt))
        torch.cuda.synchronize()
        grp.gemm_sharded(ta, tb, m, n, k, 1.5, a_blocks, lda, b_full, b.shape[1], 0.0, c_full, m, gather=gather)
        grp.wait()
        for g, d in enumerate(devs):
            got = c_full[g].view(n, m).T.double().to("cuda:0")
            r0, rows = ranges[g]
            rows_held = slice(0, m) if gather else slice(r0, r0 + rows)
            err = (got[rows_held] - want[rows_held]).abs()
            assert torch.isfinite(got[rows_held]).all(), (devs, g)
            assert (err <= tol * bound[rows_held] + 1e-300).all(), (devs, g, float((err / bound[rows_held]).max()))
            if not gather and rows < m:   # the other devices' rows were not touched
                other = torch.ones(m, dtype=torch.bool); other[r0:r0 + rows] = False
                assert torch.isnan(c_full[g].view(n, m).T[other.to(c_full[g].device)]).all()
        if gather:   # every copy is bitwise the same matrix
            for g in range(1, G):
                assert torch.equal(c_full[0].to("cuda:0"), c_full[g].to("cuda:0"))
        grp.close()


@pytest.mark.parametrize("dt,tol", [(torch.bfloat16, 1.6e-2), (torch.float16, 2e-3), (torch.float32, 1e-5)])
def test_strided_batched_sharded(handle, dt, tol):
    m = n = k = 136
    batch = 11
    for devs in _groups():
        grp = SB_Handle_Group(devs)
        G = len(devs)
        a, b = _rand((batch, k, m), dt, "cuda:0", 3), _rand((batch, n, k), dt, "cuda:0", 4)
        want = a.double().transpose(1, 2) @ b.double().transpose(1, 2)
        bound = a.double().abs().transpose(1, 2) @ b.double().abs().transpose(1, 2)
        a_s, b_s, c_s, rng = [], [], [], []
        for g, d in enumerate(devs):
            b0, cnt = sharding.split_range(batch, G, g, 1)
            rng.append((b0, cnt))
            a_s.append(a[b0:b0 + max(cnt, 1)].to(f"cuda:{d}").contiguous().view(-1))
            b_s.append(b[b0:b0 + max(cnt, 1)].to(f"cuda:{d}").contiguous().view(-1))
            c_s.append(torch.zeros(max(cnt, 1) * n * m, device=f"cuda:{d}", dtype=dt))
        torch.cuda.synchronize()
        grp.gemm_strided_batched_sharded("n", "n", m, n, k, 1.0, a_s, m, m * k, b_s, k, k * n, 0.0, c_s, m, m * n, batch)
        grp.wait()
        for g, (b0, cnt) in enumerate(rng):
            if cnt == 0:
                continue
            got = c_s[g].view(-1, n, m).transpose(1, 2).double().to("cuda:0")[:cnt]
            err = (got - want[b0:b0 + cnt]).abs()
            assert (err <= tol * bound[b0:b0 + cnt]).all(), (devs, g, float((err / bound[b0:b0 + cnt]).max()))
        grp.close()


@pytest.mark.parametrize("dt,tol", [(torch.float32, 1e-5), (torch.float64, 1e-12), (torch.bfloat16, 1.6e-2)])
@pytest.mark.parametrize("ta,tb,beta", [("n", "n", 0.0), ("t", "n", 0.5), ("n", "t", 0.0)])
def test_gemm_sharded_host(handle, dt, tol, ta, tb, beta):
    m, n, k = 1100, 700, 520
    for devs in _groups():
        grp = SB_Handle_Group(devs)
        a = _rand((k, m) if ta == "n" else (m, k), dt, "cuda:0", 5).cpu()
        b = _rand((n, k) if tb == "n" else (k, n), dt, "cuda:0", 6).cpu()
        ldc = m + 8
        c0 = _rand((n, ldc), dt, "cuda:0", 7).cpu()
        c = c0.clone()
        opa = a.double().T if ta == "n" else a.double()
        opb = b.double().T if tb == "n" else b.double()
        want = 1.5 * (opa @ opb) + beta * c0.double().T[:m]
        bound = 1.5 * (opa.abs() @ opb.abs()) + abs(beta) * c0.double().T[:m].abs()
        grp.gemm_sharded_host(ta, tb, m, n, k, 1.5, a.view(-1), a.shape[1], b.view(-1), b.shape[1], beta, c.view(-1), ldc)
        got = c.double().T
        err = (got[:m] - want).abs()
        assert (err <= tol * bound + 1e-300).all(), (devs, float((err / bound).max()))
        assert torch.equal(c.T[m:], c0.T[m:]), "ld padding of the host C was written"
        grp.close()


def test_cpp_sample_runs_over_the_group(handle):
    """samples/gemm_multi_b200.cpp through blas::multi: two shards on device 0, and all devices when there are several."""
    exe = ROOT / "build" / "gemm_multi_b200"
    if not exe.exists():
        pytest.skip("build/gemm_multi_b200 was not prebuilt")
    runs = [["2048", "2", "1"]]
    n_dev = torch.cuda.device_count()
    if n_dev >= 2:
        runs.append([str(2048 * (n_dev // 2) if n_dev > 2 else 2048), str(n_dev), "0"])
    for argv in runs:
        r = subprocess.run([str(exe), *argv], capture_output=True, text=True, timeout=300)
        assert r.returncode == 0 and "ALL PASS" in r.stdout, (argv, r.stdout[-2000:], r.stderr[-2000:])
