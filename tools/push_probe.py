"""Cost of the multicast GEMM's peer copies on ONE GPU: plain _gemm vs pbx_gemm_multicast with n local copies of C (the
pusher warp reads every tile back and stores it n-1 more times, all within the device: no NVLink, no receiver), with the
pusher (default) and with the round-1 epilogue stores (PBX_MULTICAST_PUSH=0).  Measurement aid.

    python tools/push_probe.py [--m 2048] [--n 16384] [--k 16384] [--copies 2]
"""
import argparse
import os
import statistics
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import torch  # noqa: E402

from portblas_b200 import SB_Handle, blas  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--m", type=int, default=2048)
ap.add_argument("--n", type=int, default=16384)
ap.add_argument("--k", type=int, default=16384)
ap.add_argument("--copies", type=int, default=2)
ap.add_argument("--dtype", default="f32")
args = ap.parse_args()
dt = {"f32": torch.float32, "bf16": torch.bfloat16}[args.dtype]
dev = torch.device("cuda", 0)
h = SB_Handle(0)
m, n, k = args.m, args.n, args.k
a = (torch.rand(m * k, device=dev) * 7 - 2).to(dt)
b = (torch.rand(k * n, device=dev) * 7 - 2).to(dt)
cs = [torch.zeros(m * n, device=dev, dtype=dt) for _ in range(args.copies)]
ptrs = [c.data_ptr() for c in cs]


def plain():
    blas._gemm(h, "n", "n", m, n, k, 1.0, a, m, b, k, 0.0, cs[0], m)


def multi():
    blas._gemm_multicast(h, "n", "n", m, n, k, 1.0, a, m, b, k, 0.0, ptrs, m, dt)


os.environ["PBX_MULTICAST_PUSH"] = "0"
h_nopush = SB_Handle(0)          # the switch is read into the handle when it is created
os.environ.pop("PBX_MULTICAST_PUSH", None)


def multi_nopush():
    blas._gemm_multicast(h_nopush, "n", "n", m, n, k, 1.0, a, m, b, k, 0.0, ptrs, m, dt)


runs = {"plain": plain, "multicast_push": multi, "multicast_epilogue_stores": multi_nopush}
res = {kname: [] for kname in runs}
for fn in runs.values():
    for _ in range(3):
        fn()
torch.cuda.synchronize()
for _ in range(5):
    for name, fn in runs.items():
        time.sleep(0.3)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            fn()
        e1.record()
        torch.cuda.synchronize()
        res[name].append(e0.elapsed_time(e1) / 5)
for name, v in res.items():
    print(f"{name:28s} median {statistics.median(v):8.4f} ms  min {min(v):8.4f} ms   ({args.dtype} {m}x{n}x{k}, {args.copies} copies)", flush=True)
assert torch.equal(cs[0], cs[-1])
