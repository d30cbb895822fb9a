"""Throughput of the pressure-projection solver (csrc/krylov.cu): conjugate-gradient iterations per second and the
HBM bandwidth they amount to.  Per iteration and grid point the two kernels move 36 B of fp32 vectors (p = r + b p, x += a p_old,
q = A p: read r p x, write p x q; r -= a q: read r q, write r): HBM-bound.  (The first version of the solver ran three kernels
and 44 B per point: profiles/r02_pressure_cg_scalar.json / r02_pressure_cg_3kernel.json.)
  python scripts/pressure_bench.py [iterations]
Prints one JSON line: per shape ms per iteration, GB/s at 36 B/pt, fraction of the measured HBM peak; and, for a smooth problem,
the iterations to a 1e-6 relative residual from a zero guess vs from the Neumann HPNN's prediction (seeded weights: plumbing only)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from poisson_cnn_b200.solvers import pressure_poisson_solve

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 200
dev = torch.device("cuda", 0)
peak = 6536.4
try:
    peak = float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
out = {"bytes_per_pt_per_iteration": 36, "hbm_peak_gbs": peak, "iterations_timed": iters, "shapes": {}}
g = torch.Generator().manual_seed(3)
for B, n in ((64, 256), (16, 512), (4, 1024), (1, 2048), (16, 2048)):
    rhs = torch.randn(B, 1, n, n, generator=g).to(dev)
    rhs -= rhs.mean(dim=(2, 3), keepdim=True)
    dx = torch.full((B, 1), 1.0 / (n - 1), device=dev)
    for _ in range(2):
        pressure_poisson_solve(rhs, dx, max_iter=8, rel_tol=0.0)
    ts = []
    for k in (8, 8 + iters):                                   # the difference removes the setup / centring launches
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            pressure_poisson_solve(rhs, dx, max_iter=k, rel_tol=0.0)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) / 3)
    ms = (ts[1] - ts[0]) / iters
    gbs = 36.0 * B * n * n / (ms * 1e-3) / 1e9
    out["shapes"]["%dx%dx%d" % (B, n, n)] = {"ms_per_iteration": ms, "gbs": gbs, "frac_of_hbm_peak": gbs / peak, "solve_overhead_ms": ts[0] - 8 * ms}
print(json.dumps(out))
