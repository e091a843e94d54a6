"""One interleaved-batch shape, a few launches (for ncu captures and quick timing).
usage: python tools/ilv_probe.py m n k batch [iters]"""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import torch
from portblas_b200 import SB_Handle, blas, gemm_batch_type_t

m, n, k, batch = (int(x) for x in sys.argv[1:5])
iters = int(sys.argv[5]) if len(sys.argv) > 5 else 20
h = SB_Handle(0)
dev = torch.device("cuda", 0)
a = torch.rand(m * k * batch, device=dev) * 7 - 2
b = torch.rand(k * n * batch, device=dev) * 7 - 2
c = torch.zeros(m * n * batch, device=dev)
run = lambda: blas._gemm_batched(h, "n", "n", m, n, k, 1.0, a, m, b, k, 0.0, c, m, batch, gemm_batch_type_t.interleaved)  # noqa: E731
for _ in range(3):
    run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(iters):
    run()
e1.record()
torch.cuda.synchronize()
us = e0.elapsed_time(e1) / iters * 1e3
byts = (m * k + k * n + m * n) * 4 * batch
print(f"{m}x{n}x{k} x{batch} interleaved: {us:.1f} us  {2.0 * m * n * k * batch / us / 1e6:.2f} TFLOP/s  {byts / us / 1e3:.0f} GB/s  kernel={h.last_kernel}")
