"""Seeded synthetic Poisson problems in the style of the reference's numerical dataset
(poisson_CNN/dataset/generators/numerical.py:10-72): a coarse 2*U(0,1)-1 field with a few control
points per side, bicubically up-sampled with align_corners=True (dataset/utils/image_resize.py:20)
and max-normalised to 1; boundaries the same in 1-D; dx ~ U(5e-3, 5e-2) per sample
(experiments/pcnn_end_to_end.json dataset section).  Host-side (CPU torch); used by tests and bench.
"""
import torch
import torch.nn.functional as F


def smooth_field_2d(gen, B, H, W, ctrl=(3, 8)):
    n = int(torch.randint(ctrl[0], ctrl[1] + 1, (1,), generator=gen))
    coarse = 2 * torch.rand(B, 1, n, n, generator=gen) - 1
    f = F.interpolate(coarse, size=(H, W), mode="bicubic", align_corners=True)
    return f / f.abs().amax(dim=(1, 2, 3), keepdim=True)


def smooth_field_1d(gen, B, n_out, ctrl=(3, 8)):
    n = int(torch.randint(ctrl[0], ctrl[1] + 1, (1,), generator=gen))
    coarse = 2 * torch.rand(B, 1, 1, n, generator=gen) - 1
    f = F.interpolate(coarse, size=(1, n_out), mode="bicubic", align_corners=True)[:, :, 0, :]
    return f / f.abs().amax(dim=(1, 2), keepdim=True)


def make_problem(B, nx, ny, seed, magnitudes=True):
    """Returns dict of CPU float32 tensors: rhs [B,1,nx,ny], left/right [B,1,ny], top/bottom [B,1,nx], dx [B,1].
    With magnitudes=True each field gets a random per-sample amplitude so the max-normalisation and the
    1/scaling-factor rescale of Poisson_CNN_Legacy are exercised."""
    gen = torch.Generator().manual_seed(int(seed))
    p = {
        "rhs": smooth_field_2d(gen, B, nx, ny),
        "left": smooth_field_1d(gen, B, ny),
        "top": smooth_field_1d(gen, B, nx),
        "right": smooth_field_1d(gen, B, ny),
        "bottom": smooth_field_1d(gen, B, nx),
        "dx": 5e-3 + (5e-2 - 5e-3) * torch.rand(B, 1, generator=gen),
    }
    if magnitudes:
        for k in ("rhs", "left", "top", "right", "bottom"):
            amp = 0.25 + 1.75 * torch.rand(B, generator=gen)
            p[k] = p[k] * amp.view(-1, *([1] * (p[k].dim() - 1)))
    return {k: v.float().contiguous() for k, v in p.items()}
