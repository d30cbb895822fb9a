"""Pressure-Poisson solve of a projection-method Navier-Stokes step, with the Neumann HPNN as the initial guess.

The caller AFTER the hot path.  Reference: Navier_Stokes_2D/solvers.py
  :29-33    the model: Homogeneous_Poisson_NN_Legacy with the hpnn_neumann_piloss_smalldomain config (bc_type 'neumann'),
  :153-186  Poisson_pressure_matrix: minus the cell-centred 5-point Laplacian with homogeneous Neumann boundaries plus a
            zero-integral Lagrange row,
  :204-334  Poisson_pressure_solver: the (commented-out) network prediction, rescaled by (dx (n-1))^2 / sf, is the x0 of a
            Krylov iteration on that system.
The reference iterates with scipy's BiCGStab + ILU on the CPU and writes plots; here the whole solve is batched conjugate
gradients on the GPU (csrc/krylov.cu), the network prediction comes from the hot path, nothing synchronises with the host.
"""
import torch

from .. import ops
from .._lib import lib, check


def neumann_laplacian_apply(v, dx):
    """(A v) with A = minus the cell-centred Neumann 5-point Laplacian (solvers.py:164-173).  v [B,1,m,n], dx [B,1] or [B]."""
    ops._chk(v, "v")
    ops._chk(dx, "dx")
    v = v.contiguous()
    B, _, H, W = v.shape
    out = torch.empty_like(v)
    check(lib.pcnn_neumann_laplacian_apply_f32(v.data_ptr(), dx.contiguous().data_ptr(), out.data_ptr(), B, H, W, ops._stream()), "neumann_laplacian_apply")
    return out


def hpnn_initial_guess(model, rhs, dx):
    """The reference's use of the network (solvers.py:246-250): max-normalise the right-hand side, predict, undo the
    normalisations: pred * (dx (n-1))^2 / sf with sf = 1 / max|rhs|.  Returns (raw prediction, per-sample scale)."""
    m = ops.maxabs(rhs)
    pred = model([ops.scale_inv(rhs, m), dx])
    n = rhs.shape[3]
    scale = (dx.reshape(-1) * float(n - 1)) ** 2 * m            # [B] plumbing on a B-element vector
    return pred, scale.contiguous()


def pressure_poisson_solve(rhs, dx, model=None, x0=None, max_iter=200, rel_tol=1e-6, return_history=False):
    """Solve  laplace(p) = rhs  with homogeneous Neumann boundaries and zero mean (Poisson_pressure_solver, solvers.py:204-334).
    rhs [B,1,m,n] (CUDA), dx [B,1]; model: a Neumann Homogeneous_Poisson_NN_Legacy whose prediction seeds the iteration, or
    x0: an explicit initial guess [B,1,m,n] (neither: zero initial guess, the reference's current default).  Returns p
    [B,1,m,n] (and the [max_iter,B] history of relative residuals |r_k|/|b| when return_history).  No host synchronisation.
    Vectors are fp32: the iteration reaches ~eps * cond * max(|x0|, |p|), so a guess far larger than the solution costs digits."""
    ops._chk(rhs, "rhs")
    ops._chk(dx, "dx")
    if rhs.dim() != 4 or rhs.shape[1] != 1:
        raise ValueError("rhs must be [batch, 1, m, n] (channels_first)")
    B, _, H, W = rhs.shape
    rhs = rhs.contiguous()
    dxv = dx.reshape(B).contiguous()
    if model is not None and x0 is not None:
        raise ValueError("pass either model= or x0=, not both")
    if model is not None:
        x, scale = hpnn_initial_guess(model, rhs, dx.reshape(B, 1).contiguous())
        x = x.contiguous()
    elif x0 is not None:
        ops._chk(x0, "x0")
        if tuple(x0.shape) != tuple(rhs.shape):
            raise ValueError("x0 must have the shape of rhs")
        x, scale = x0.clone().contiguous(), None
    else:
        x, scale = torch.zeros_like(rhs), None
    work = torch.empty((lib.pcnn_neumann_cg_workspace_bytes(B, H, W) + 7) // 8, device=rhs.device, dtype=torch.float64)
    hist = torch.zeros((max(max_iter, 1), B), device=rhs.device, dtype=torch.float64) if return_history else None
    check(lib.pcnn_neumann_cg_solve(rhs.data_ptr(), dxv.data_ptr(), None if scale is None else scale.data_ptr(), x.data_ptr(), B, H, W,
                                    int(max_iter), float(rel_tol), None if hist is None else hist.data_ptr(), work.data_ptr(),
                                    ops._stream()), "neumann_cg_solve")
    return (x, hist) if return_history else x
