"""Random smooth Poisson problems generated on the GPU, with DST ground truth.

Reference: poisson_CNN/dataset/generators/numerical.py:10-150 (generate_random_RHS,
generate_random_boundaries, numerical_dataset), dataset/utils/image_resize.py:5-30 (legacy bicubic
resize with align_corners=True) and dataset/utils/set_max_magnitude.py:4-24.  The random control points
come from torch's device RNG (plumbing); the up-sampling, the max-normalisation and the ground-truth
solve are libpcnn kernels (pcnn_resize_f32 with the legacy-bicubic tables, pcnn_maxabs_f32 /
pcnn_scale_inv_f32, pcnn_dst_solve).  The reference's multigrid / Cholesky solvers are replaced by the
DST-I direct solve of the same discrete system (solver_method='dst')."""
import math

import numpy as np
import torch

from .. import ops
from ..config import RESIZE_BICUBIC_LEGACY_AC

_SIDES = ("left", "top", "right", "bottom")


def _device(device):
    dev = torch.device("cuda" if device is None else device)
    if dev.type != "cuda":
        raise ValueError("the dataset generators run on a CUDA device (there is no CPU path)")
    return dev


def image_resize(image, newshape):
    """image [B,C,h,w] (or [B,n]: one axis, like the reference's 2-D call) -> newshape, legacy bicubic with
    align_corners=True (dataset/utils/image_resize.py:20)."""
    if image.dim() == 2:
        return ops.resize(image[:, None, None, :].contiguous(), (1, int(newshape[-1])), RESIZE_BICUBIC_LEGACY_AC)[:, 0, 0, :]
    return ops.resize(image, (int(newshape[0]), int(newshape[1])), RESIZE_BICUBIC_LEGACY_AC)


def set_max_magnitude_in_batch(arr, max_magnitude):
    """arr * (max_magnitude / max|arr|) per sample (dataset/utils/set_max_magnitude.py:4-24)."""
    shape = arr.shape
    x = arr.reshape(shape[0], 1, 1, -1).contiguous()
    m = ops.maxabs(x)
    mm = torch.as_tensor(max_magnitude, dtype=torch.float32, device=arr.device).expand(shape[0]).contiguous()
    return ops.scale_inv(x, m / mm).reshape(shape)


def _smoothness(smoothness, n_outputpts, rng):
    if smoothness is None:       # numerical.py:22-23: randint(5, n//1.5) per dimension
        return [int(rng.integers(5, max(6, int(n // 1.5)))) for n in n_outputpts]
    if isinstance(smoothness, (int, np.integer)):
        return [int(smoothness)] * len(n_outputpts)
    return [int(s) for s in smoothness]


def generate_random_RHS(batch_size, n_outputpts, smoothness=None, max_magnitude=np.inf, device=None, generator=None):
    """[batch_size, 1, nx, ny]: a 2*U(0,1)-1 field on `smoothness` control points per side, bicubically
    super-sampled to n_outputpts (numerical.py:10-35)."""
    dev = _device(device)
    rng = np.random.default_rng(None if generator is None else int(generator.initial_seed()))
    ctrl = _smoothness(smoothness, n_outputpts, rng)
    rhs = 2 * torch.rand([batch_size, 1] + ctrl, device=dev, dtype=torch.float32, generator=generator) - 1
    rhs = image_resize(rhs, n_outputpts)
    if max_magnitude != np.inf:
        rhs = set_max_magnitude_in_batch(rhs, max_magnitude)
    return rhs


def generate_random_boundaries(n_outputpts, batch_size=1, max_magnitude=None, smoothness=None,
                               nonzero_boundaries=("left", "right", "bottom", "top"), return_with_expanded_dims=False,
                               data_format="channels_first", device=None, generator=None):
    """dict left/right [B,ny], top/bottom [B,nx] ([B,1,n] with return_with_expanded_dims), numerical.py:37-78."""
    dev = _device(device)
    rng = np.random.default_rng(None if generator is None else int(generator.initial_seed()) + 1)
    if max_magnitude is None:
        max_magnitude = {k: np.inf for k in _SIDES}
    lengths = {"left": n_outputpts[1], "right": n_outputpts[1], "top": n_outputpts[0], "bottom": n_outputpts[0]}
    if isinstance(smoothness, (int, np.integer)):
        smoothness = {k: int(smoothness) for k in _SIDES}
    elif smoothness is None:
        smoothness = {k: int(rng.integers(5, max(6, int(lengths[k] // 1.5)))) for k in _SIDES}
    out = {}
    for side in ("left", "right", "top", "bottom"):
        n = int(lengths[side])
        if side in nonzero_boundaries:
            b = 2 * torch.rand((batch_size, int(smoothness[side])), device=dev, dtype=torch.float32, generator=generator) - 1
            b = image_resize(b, [batch_size, n])
            if max_magnitude[side] != np.inf:
                b = set_max_magnitude_in_batch(b, max_magnitude[side])
        else:
            b = torch.zeros((batch_size, n), device=dev, dtype=torch.float32)
        if return_with_expanded_dims:
            b = b.unsqueeze(1 if data_format == "channels_first" else 2)
        out[side] = b
    return out


def numerical_dataset(batch_size=1, output_shape=(64, 64), dx="random", boundaries="random", rhses="random",
                      rhs_smoothness=None, boundary_smoothness=None, rhs_max_magnitude=1.0, boundary_max_magnitude=None,
                      nonzero_boundaries=("left", "right", "bottom", "top"), solver_method="dst", return_rhs=True,
                      return_boundaries=False, return_dx=False, random_dx_range=(0.005, 0.05),
                      normalize_by_domain_size=False, device=None, generator=None):
    """RHS / BC / dx triples with their ground-truth solution (numerical.py:80-150).  One grid shape per batch
    (`output_shape`); solver_method 'dst' (the GPU DST-I solve) or a callable (rhses, boundaries, dx) -> solution.
    Returns ([rhs?, boundaries?, dx?], solution) or just the solution when nothing else is requested."""
    dev = _device(device)
    output_shape = [int(n) for n in output_shape]
    if boundary_max_magnitude is None:
        boundary_max_magnitude = {k: 1.0 for k in _SIDES}
    if isinstance(dx, str) and dx == "random":
        dx = torch.rand((batch_size, 1), device=dev, dtype=torch.float32, generator=generator) * (random_dx_range[1] - random_dx_range[0]) + random_dx_range[0]
    elif isinstance(dx, float):
        dx = torch.full((batch_size, 1), dx, device=dev, dtype=torch.float32)
    if isinstance(rhses, str) and rhses == "random":
        rhses = generate_random_RHS(batch_size, output_shape, smoothness=rhs_smoothness, max_magnitude=rhs_max_magnitude, device=dev, generator=generator)
    elif isinstance(rhses, str) and rhses == "zero":
        rhses = torch.zeros([batch_size, 1] + output_shape, device=dev, dtype=torch.float32)
    if isinstance(boundaries, str):
        nz = nonzero_boundaries if boundaries == "random" else ()
        boundaries = generate_random_boundaries(output_shape, batch_size=batch_size, max_magnitude=boundary_max_magnitude,
                                                return_with_expanded_dims=True, nonzero_boundaries=nz,
                                                smoothness=boundary_smoothness, device=dev, generator=generator)
    if callable(solver_method):
        out = solver_method(rhses, boundaries, dx)
    elif solver_method == "dst":
        from ..solvers.dst import dst_poisson_solve
        out = dst_poisson_solve(rhses, boundaries, dx)
    else:
        raise ValueError("solver_method must be a callable or 'dst' (the reference's multigrid/cholesky solvers are replaced by the GPU DST solve)")
    if normalize_by_domain_size:
        domainsize = dx.reshape(-1) ** len(output_shape) * float(math.prod(n - 1 for n in output_shape))
        out = 10 * out / domainsize.view(-1, 1, 1, 1)
    inp = []
    if return_rhs:
        inp.append(rhses)
    if return_boundaries:
        inp.append(boundaries)
    if return_dx:
        inp.append(dx)
    return (inp, out) if inp else out
