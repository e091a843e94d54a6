"""Device time of one GEMM shape (CUDA-graph replays).  usage: python tools/shape_probe.py dtype ta tb m n k [batch]"""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import torch
from portblas_b200 import SB_Handle, blas

dtn, ta, tb = sys.argv[1:4]
m, n, k = (int(x) for x in sys.argv[4:7])
batch = int(sys.argv[7]) if len(sys.argv) > 7 else 1
dt = {"f32": torch.float32, "f64": torch.float64, "f16": torch.float16, "bf16": torch.bfloat16}[dtn]
dev = torch.device("cuda", 0)
h = SB_Handle(0)
a = (torch.rand(m * k * batch, device=dev) * 7 - 2).to(dt)
b = (torch.rand(k * n * batch, device=dev) * 7 - 2).to(dt)
c = torch.zeros(m * n * batch, device=dev, dtype=dt)
lda, ldb = (m if ta == "n" else k), (k if tb == "n" else n)


def run():
    if batch == 1:
        blas._gemm(h, ta, tb, m, n, k, 1.0, a, lda, b, ldb, 0.0, c, m)
    else:
        blas._gemm_strided_batched(h, ta, tb, m, n, k, 1.0, a, lda, m * k, b, ldb, k * n, 0.0, c, m, m * n, batch)


for _ in range(3):
    run()
torch.cuda.synchronize()
iters = 20
evs = [torch.cuda.Event(enable_timing=True) for _ in range(iters + 1)]
evs[0].record()
for i in range(iters):
    run()
    evs[i + 1].record()
torch.cuda.synchronize()
ts = sorted(evs[i].elapsed_time(evs[i + 1]) for i in range(iters))
ms = ts[len(ts) // 2]
fl = 2.0 * m * n * k * batch
es = a.element_size()
print(f"{dtn} {ta}{tb} {m}x{n}x{k} x{batch}: {ms * 1e3:.1f} us  {fl / ms / 1e9:.1f} TFLOP/s  "
      f"{(m * k + k * n + m * n) * es * batch / ms / 1e6:.0f} GB/s  kernel={h.last_kernel} split_k={h.last_split_k}")
