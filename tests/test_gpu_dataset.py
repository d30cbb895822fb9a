"""GPU problem generator (dataset/generators/numerical.py + dataset/utils/image_resize.py)
against the oracle restatement, and the generated problems against the discrete system they claim to solve."""
import numpy as np
import pytest
import torch

from oracle import poisson_oracle as O
from tests.helpers import rel_l2

pytestmark = pytest.mark.gpu


def test_image_resize_matches_oracle():
    from poisson_cnn_b200 import dataset
    g = torch.Generator().manual_seed(11)
    for shape, out in [((4, 1, 5, 7), (64, 96)), ((2, 1, 8, 3), (200, 300)), ((3, 1, 20, 20), (256, 256))]:
        x = 2 * torch.rand(*shape, generator=g) - 1
        got = dataset.image_resize(x.cuda(), out)
        ref = O.image_resize(x.double(), out)
        assert got.shape == ref.shape
        assert float((got.double().cpu() - ref).abs().max()) < 2e-6          # fp32 evaluation of the same tables
    b = 2 * torch.rand(6, 9, generator=g) - 1
    got = dataset.image_resize(b.cuda(), [6, 130])
    ref = O.image_resize(b.double()[:, None, None, :], (1, 130))[:, 0, 0, :]
    assert got.shape == (6, 130) and float((got.double().cpu() - ref).abs().max()) < 2e-6


def test_set_max_magnitude_in_batch():
    from poisson_cnn_b200 import dataset
    g = torch.Generator().manual_seed(12)
    x = torch.randn(5, 1, 33, 47, generator=g) * torch.tensor([0.1, 1.0, 7.0, 30.0, 1e-3]).view(5, 1, 1, 1)
    got = dataset.set_max_magnitude_in_batch(x.cuda(), 2.5)
    ref, _ = O.set_max_magnitude(x.double(), 2.5)
    assert rel_l2(got, ref) < 3e-7
    np.testing.assert_allclose(got.abs().amax(dim=(1, 2, 3)).cpu().numpy(), 2.5, rtol=3e-7)


def test_generators_shapes_ranges_and_reproducibility():
    from poisson_cnn_b200 import dataset
    gen = torch.Generator(device="cuda").manual_seed(3)
    rhs = dataset.generate_random_RHS(4, [96, 80], smoothness=6, max_magnitude=1.0, generator=gen)
    assert rhs.shape == (4, 1, 96, 80) and rhs.dtype == torch.float32 and rhs.is_cuda
    np.testing.assert_allclose(rhs.abs().amax(dim=(1, 2, 3)).cpu().numpy(), 1.0, rtol=3e-7)
    bcs = dataset.generate_random_boundaries([96, 80], batch_size=4, smoothness=5, nonzero_boundaries=["left", "top"],
                                             return_with_expanded_dims=True, generator=gen)
    assert bcs["left"].shape == (4, 1, 80) and bcs["top"].shape == (4, 1, 96)
    assert float(bcs["right"].abs().max()) == 0 and float(bcs["bottom"].abs().max()) == 0 and float(bcs["left"].abs().max()) > 0
    again = dataset.generate_random_RHS(4, [96, 80], smoothness=6, max_magnitude=1.0, generator=torch.Generator(device="cuda").manual_seed(3))
    assert torch.equal(rhs, again)
    with pytest.raises(ValueError):
        dataset.generate_random_RHS(1, [8, 8], device="cpu")


@pytest.mark.parametrize("shape", [(64, 64), (120, 88)])
def test_numerical_dataset_solution_solves_the_reference_system(shape):
    """numerical.py:80-150 with the DST solve in place of pyamg: the returned solution reproduces its boundary data
    and satisfies the reference's 5-point system (the oracle's float64 DST solve of the same inputs agrees)."""
    from poisson_cnn_b200 import dataset
    gen = torch.Generator(device="cuda").manual_seed(21)
    (rhs, bcs, dx), sol = dataset.numerical_dataset(batch_size=3, output_shape=shape, return_boundaries=True, return_dx=True,
                                                    rhs_smoothness=7, boundary_smoothness=5, generator=gen)
    assert sol.shape == (3, 1) + shape and dx.shape == (3, 1)
    assert float(dx.min()) >= 0.005 and float(dx.max()) <= 0.05
    ref = O.dst_poisson_solve(rhs.double().cpu(), bcs["left"][:, 0].double().cpu(), bcs["top"][:, 0].double().cpu(),
                              bcs["right"][:, 0].double().cpu(), bcs["bottom"][:, 0].double().cpu(), dx[:, 0].double().cpu())
    assert rel_l2(sol, ref) < 1e-5
    # ring := BCs (multigrid.py:145-148): left = u[0, :], right = u[-1, :], bottom = u[:, 0], top = u[:, -1]
    np.testing.assert_allclose(sol[:, 0, 0, :].cpu().numpy(), bcs["left"][:, 0].cpu().numpy(), atol=1e-6)
    np.testing.assert_allclose(sol[:, 0, 1:-1, -1].cpu().numpy(), bcs["top"][:, 0, 1:-1].cpu().numpy(), atol=1e-6)
    # the interior satisfies the 3-point Laplacian system
    gs = torch.cat([dx, dx], 1)
    res = O.laplacian_residual(rhs.double().cpu(), sol.double().cpu(), gs.double().cpu(), 3)
    assert float(res) < 1e-4 * float((rhs.double() ** 2).mean())
