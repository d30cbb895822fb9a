"""The memory-bound check kernels at the config-4 shape, for `ncu --set full`: Laplacian residual (stencil 3 and 5) and the
FFT DST solve (float64).  usage: profile_checks.py [n] [B]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from poisson_cnn_b200 import ops
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
B = int(sys.argv[2]) if len(sys.argv) > 2 else 16
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(1)
u = torch.randn((B, 1, n, n), device=dev, generator=g)
f = torch.randn((B, 1, n, n), device=dev, generator=g)
bc = [torch.randn((B, 1, n), device=dev, generator=g) for _ in range(4)]
dx = 5e-3 + 4.5e-2 * torch.rand((B, 1), device=dev, generator=g)
gs = torch.cat([dx, dx], 1)
r3 = ops.laplacian_residual(f, u, gs, 3)
r5 = ops.laplacian_residual(f, u, gs, 5)
sol = ops.dst_solve(f, bc[0], bc[1], bc[2], bc[3], dx)
torch.cuda.synchronize()
print("residual sums", float(r3.sum()), float(r5.sum()), "dst finite", bool(torch.isfinite(sol).all()))
