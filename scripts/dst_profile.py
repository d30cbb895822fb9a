"""Times the three passes of the FFT DST solve (run under `ncu --metrics gpu__time_duration.sum` for per-kernel times).
usage: python scripts/dst_profile.py [n] [B] [f32|f64]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from poisson_cnn_b200 import ops
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
B = int(sys.argv[2]) if len(sys.argv) > 2 else 16
dt = torch.float32 if (len(sys.argv) > 3 and sys.argv[3] == "f32") else torch.float64
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(1)
f = torch.randn((B, 1, n, n), device=dev, generator=g)
bc = [torch.randn((B, 1, n), device=dev, generator=g) for _ in range(4)]
dx = 5e-3 + 4.5e-2 * torch.rand((B, 1), device=dev, generator=g)
for _ in range(2):
    u = ops.dst_solve(f, bc[0], bc[1], bc[2], bc[3], dx, dtype=dt)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    u = ops.dst_solve(f, bc[0], bc[1], bc[2], bc[3], dx, dtype=dt)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print("DST fft %s %dx%dx%d: %.3f ms, %.0f solutions/s, %.1f GB/s at 8 B/pt" % (dt, B, n, n, ms, B / ms * 1e3, B * n * n * 8 / ms / 1e6))
