"""Fixed per-launch overhead of the batched tcgen05 path: device time of bf16 strided-batched 256^3 GEMMs as a
function of the batch count (the intercept of the line is what strong scaling over N GPUs pays N times)."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import torch
from portblas_b200 import SB_Handle, blas

h = SB_Handle(0)
dev = torch.device("cuda", 0)
m = n = k = 256
bmax = 4096
a = (torch.rand(m * k * bmax, device=dev) * 7 - 2).to(torch.bfloat16)
b = (torch.rand(m * k * bmax, device=dev) * 7 - 2).to(torch.bfloat16)
c = torch.zeros(m * n * bmax, device=dev, dtype=torch.bfloat16)
for batch in (1, 37, 74, 148, 296, 512, 592, 1024, 1184, 2048, 2368, 4096):
    def run():
        blas._gemm_strided_batched(h, "n", "n", m, n, k, 1.0, a, m, m * k, b, k, k * n, 0.0, c, m, m * n, batch)
    for _ in range(5):
        run()
    torch.cuda.synchronize()
    iters = 30
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(iters + 1)]
    evs[0].record()
    for i in range(iters):
        run()
        evs[i + 1].record()
    torch.cuda.synchronize()
    ts = sorted(evs[i].elapsed_time(evs[i + 1]) * 1e3 for i in range(iters))
    # single isolated launch (GPU idle before): event pair around one call
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); s.record(); run(); e.record(); torch.cuda.synchronize()
    print(f"batch {batch:5d}: back-to-back median {ts[len(ts)//2]:8.1f} us  min {ts[0]:8.1f} us   isolated {s.elapsed_time(e)*1e3:8.1f} us   kernel={h.last_kernel}", flush=True)
