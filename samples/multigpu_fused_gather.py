"""M-block sharded GEMM on N GPUs of one box whose C gather is fused into the GEMM's stores.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 samples/multigpu_fused_gather.py [m n k]

One process per GPU.  Every rank holds A (or just its row block of it), B, and a full-size C; the ranks exchange CUDA
IPC handles of their C once, then each runs its M-block with pbx_gemm_multicast: the tensor-core epilogue stores every
finished tile into the same rows of every rank's C over NVLink while the next tile computes.  After a barrier each rank
holds the whole product -- no all-gather, no staging buffer.  (The same steps through the C-ABI: pbx_ipc_export,
pbx_ipc_import, pbx_gemm_multicast -- include/pbx_gemm.h.)
"""
import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from portblas_b200 import SB_Handle, sharding  # noqa: E402


def main():
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    m, n, k = (int(x) for x in sys.argv[1:4]) if len(sys.argv) >= 4 else (8192, 8192, 8192)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29533")
    dist.init_process_group("nccl" if world > 1 else "gloo", rank=rank, world_size=world, device_id=dev if world > 1 else None)
    h = SB_Handle(local)
    gen = torch.Generator(device=dev).manual_seed(1)          # same operands on every rank
    a = torch.rand(m * k, device=dev, generator=gen) * 7 - 2  # column-major m x k, lda = m
    b = torch.rand(k * n, device=dev, generator=gen) * 7 - 2  # column-major k x n, ldb = k
    c = torch.zeros(m * n, device=dev)                        # the FULL C, on every rank
    ptrs = sharding.share_full_c(h, c)                        # CUDA IPC: peers' C mapped into this process
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    sh = sharding.shard_mblock("n", m, m, world, rank, align=256)
    sharding.gemm_mblock_fused_gather(h, "n", "n", m, n, k, 1.0, a[sh.a_offset:], m, b, k, 0.0, ptrs, m, torch.float32,
                                      world, rank, align=256)
    e1.record()
    h.wait()
    dist.barrier()                                            # every rank's local and remote stores are complete
    torch.cuda.synchronize()
    # spot check against an fp64 product of 64 x 64 sampled entries -- rows owned by OTHER ranks included
    ri = torch.randint(0, m, (64,), device=dev)
    ci = torch.randint(0, n, (64,), device=dev)
    want = a.view(k, m).T[ri].double() @ b.view(n, k).T[:, ci].double()
    bound = a.view(k, m).T[ri].double().abs() @ b.view(n, k).T[:, ci].double().abs()
    rel = float(((c.view(n, m).T[ri][:, ci].double() - want).abs() / bound).max())
    print(f"rank {rank}/{world}: rows [{sh.row0}, {sh.row0 + sh.rows}) in {e0.elapsed_time(e1):.2f} ms, "
          f"{2.0 * sh.rows * n * k / e0.elapsed_time(e1) / 1e9:.1f} TFLOP/s, full-C max rel err {rel:.2e}", flush=True)
    assert rel <= 1e-5
    h.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
