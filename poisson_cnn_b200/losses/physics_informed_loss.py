"""linear_operator_loss: the finite-difference Laplacian residual check on the GPU.

Mirrors poisson_CNN/losses/physics_informed_loss.py:6-50 (same constructor arguments, same
__call__(rhs, solution, grid_spacings) -> scalar mean squared residual) and additionally exposes
per-sample squared norms, which the accuracy sweeps need.
"""
import torch

from .. import ops


class linear_operator_loss:
    def __init__(self, stencil_sizes, orders, ndims=None, data_format="channels_first", normalize=False,
                 inputs_have_max_domain_size_squared_normalization=False):
        if ndims is None:
            try:
                ndims = len(stencil_sizes)
            except TypeError:
                try:
                    ndims = len(orders)
                except TypeError:
                    raise ValueError("If ndims is not supplied, one of stencil_sizes or orders must be a list containing as many elements as there are dimensions")
        self.ndims = ndims
        sizes = [stencil_sizes] * ndims if isinstance(stencil_sizes, int) else list(stencil_sizes)
        ords = [orders] * ndims if isinstance(orders, int) else list(orders)
        if ndims != 2 or len(set(sizes)) != 1 or sizes[0] not in (3, 5) or any(o != 2 for o in ords):
            raise NotImplementedError("the CUDA residual covers the 2-D Laplacian with stencil size 3 or 5 (orders 2)")
        if data_format != "channels_first":
            raise NotImplementedError("channels_first only")
        self.stencil_size = sizes[0]
        self.data_format = data_format
        self.normalize = normalize
        self.inputs_have_max_domain_size_squared_normalization = inputs_have_max_domain_size_squared_normalization

    def per_sample_squared_sums(self, rhs, solution, grid_spacings):
        """[B] float64: sum over the interior of (rhs - Lap(solution))^2 (divided by max|rhs|^2 if normalize)."""
        gs = grid_spacings
        if gs.shape[1] == 1:
            gs = gs.repeat(1, 2)
        if self.inputs_have_max_domain_size_squared_normalization:
            # q = (Lmax/dx)^2  <=>  effective spacing dx/Lmax  (physics_informed_loss.py:36-37)
            H, W = solution.shape[2], solution.shape[3]
            sizes = gs * torch.tensor([H - 1, W - 1], device=gs.device, dtype=gs.dtype)
            gs = gs / sizes.max(1, keepdim=True).values
        m = ops.maxabs(rhs) if self.normalize else None
        return ops.laplacian_residual(rhs, solution, gs, self.stencil_size, m)

    def __call__(self, rhs, solution, grid_spacings):
        sq = self.per_sample_squared_sums(rhs, solution, grid_spacings)
        h = self.stencil_size // 2
        n = (rhs.shape[2] - 2 * h) * (rhs.shape[3] - 2 * h)
        return (sq.sum() / (rhs.shape[0] * n)).to(torch.float32)
