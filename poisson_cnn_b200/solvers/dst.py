"""GPU DST-I direct Poisson solve: ground truth for accuracy checks.

Solves the same discrete system as the reference's multigrid_poisson_solve
(poisson_CNN/dataset/solvers/multigrid.py:98-150; right-hand side assembled like poisson_RHS,
poisson_CNN/dataset/solvers/cholesky.py:45-119) directly, in double precision on the device.
"""
from .. import ops


def dst_poisson_solve(rhses, boundaries, dx, method="fft", dtype=None):
    """rhses [B,1,nx,ny]; boundaries dict with 'left','right' [B,ny] (or [B,1,ny]) and 'top','bottom'
    [B,nx]; dx [B] or [B,1].  Returns [B,1,nx,ny] float32 on the same device.
    method 'fft' (Bluestein/FFT transforms, HBM-bound; float64 arithmetic unless dtype=torch.float32) or 'gemm' (dense
    sine-matrix products in float64, the independent cross-check)."""
    import torch
    B = rhses.shape[0]
    b = {k: v.reshape(B, -1) for k, v in boundaries.items()}
    return ops.dst_solve(rhses, b["left"], b["top"], b["right"], b["bottom"], dx.reshape(B), method=method,
                         dtype=torch.float64 if dtype is None else dtype)
