"""GPU tests of the caller after the hot path: the pressure-Poisson solve of the reference's
Navier-Stokes projection solver (Navier_Stokes_2D/solvers.py:153-334) as batched conjugate gradients seeded by the Neumann
HPNN, against the oracle's direct solve of the reference's augmented system and its float64 CG restatement."""
import pytest
import torch

from tests.helpers import rel_l2
from oracle import poisson_oracle as O

pytestmark = pytest.mark.gpu


def _rhs(B, m, n, seed):
    from poisson_cnn_b200.synthetic import make_problem
    p = make_problem(B, m, n, seed=seed)
    return p["rhs"], p["dx"]


def test_neumann_operator_matches_reference_matrix():
    from poisson_cnn_b200.solvers import neumann_laplacian_apply
    g = torch.Generator().manual_seed(3)
    v = torch.randn(2, 1, 9, 12, generator=g)
    dx = torch.tensor([[0.03], [0.011]])
    got = neumann_laplacian_apply(v.cuda(), dx.cuda()).cpu().double()
    for b in range(2):
        A = O.pressure_poisson_matrix(9, 12, float(dx[b]))
        ref = torch.from_numpy(A[:-1, :-1] @ v[b, 0].double().reshape(-1).numpy()).reshape(9, 12)
        assert rel_l2(got[b, 0], ref) < 1e-5


@pytest.mark.parametrize("B,m,n", [(2, 48, 40), (3, 64, 64), (1, 33, 130)])
def test_pressure_solve_converges_to_the_reference_system(B, m, n):
    from poisson_cnn_b200.solvers import pressure_poisson_solve
    rhs, dx = _rhs(B, m, n, seed=60 + m)
    ref = O.pressure_poisson_reference(rhs, dx)
    p, hist = pressure_poisson_solve(rhs.cuda(), dx.cuda(), max_iter=600, rel_tol=1e-6, return_history=True)
    assert p.shape == (B, 1, m, n)
    assert rel_l2(p, ref) < 2e-4                       # fp32 vectors, condition number ~ (n/pi)^2
    assert float(p.mean(dim=(1, 2, 3)).abs().max()) < 1e-6 * float(p.abs().max()) + 1e-9
    assert float(hist[-1].max()) <= 1.01e-6            # every sample reached the tolerance and froze there
    # the first iterations follow the float64 CG restatement
    _, h64 = O.neumann_cg(rhs, dx, torch.zeros_like(rhs), 10)
    assert torch.allclose(hist[:10].cpu(), h64, rtol=1e-3)


def test_pressure_solve_with_an_initial_guess():
    """A good initial guess (what a TRAINED Neumann HPNN provides) cuts the iteration count; the solution is the same."""
    from poisson_cnn_b200.solvers import pressure_poisson_solve
    rhs, dx = _rhs(2, 96, 112, seed=72)
    rhs, dx = rhs.cuda(), dx.cuda()
    ref, h0 = pressure_poisson_solve(rhs, dx, max_iter=900, rel_tol=1e-6, return_history=True)
    # a SMOOTH 0.1 % error, as a trained surrogate might leave (white noise would be the worst case: A amplifies it by ~8 n^2 / pi^2)
    guess = 0.999 * ref
    p, h1 = pressure_poisson_solve(rhs, dx, x0=guess, max_iter=900, rel_tol=1e-6, return_history=True)
    assert rel_l2(p, ref) < 2e-4
    its = lambda h: int((h.max(dim=1).values > 1.01e-6).sum())
    print("iterations to 1e-6: zero guess %d, guess with 0.1 %% smooth error %d" % (its(h0), its(h1)))
    assert its(h1) < its(h0)


def test_pressure_solve_seeded_by_the_neumann_hpnn():
    """The reference's use of the network (solvers.py:246-262): the Neumann HPNN's prediction, rescaled by (dx (n-1))^2 / sf,
    as x0.  The weights are seeded noise (no trained weights ship): the guess is ~1000x the solution and fp32 CG cannot
    remove it completely, so what is checked is the plumbing -- the rescale formula, that the iteration starts from the
    residual of THAT guess, reduces it, and stays finite."""
    from poisson_cnn_b200 import convert_tf_object_names, load_experiment, models, weights as W, ops
    from poisson_cnn_b200.solvers import pressure_poisson_solve, hpnn_initial_guess, neumann_laplacian_apply
    cfg = load_experiment("hpnn_neumann")["model"]
    w = W.synthetic_weights(W.hpnn_weight_specs(cfg, "hpnn/"), seed=0)
    model = models.Homogeneous_Poisson_NN_Legacy(**convert_tf_object_names(cfg)).load_weights(w, "hpnn/").set_precision("mixed")
    rhs, dx = _rhs(2, 112, 120, seed=71)
    rhs, dx = rhs.cuda(), dx.cuda()
    pred, scale = hpnn_initial_guess(model, rhs, dx)
    m = rhs.abs().amax(dim=(1, 2, 3))
    assert torch.equal(pred, model([rhs * (1.0 / m).view(-1, 1, 1, 1), dx]))                 # sf = 1 / max|rhs| (set_max_magnitude)
    assert torch.allclose(scale, (dx[:, 0] * 119.0) ** 2 * m, rtol=1e-6)                      # (dx (n-1))^2 / sf
    p, hist = pressure_poisson_solve(rhs, dx, model=model, max_iter=300, rel_tol=1e-6, return_history=True)
    x0 = pred * scale.view(-1, 1, 1, 1)
    b = -rhs - (-rhs).mean(dim=(1, 2, 3), keepdim=True)
    r0 = b - neumann_laplacian_apply(x0, dx)
    rel0 = (r0.flatten(1).norm(dim=1) / b.flatten(1).norm(dim=1)).double()
    print("HPNN-seeded solve (noise weights): initial relative residual %s, after 300 iterations %s" % (rel0.tolist(), hist[-1].tolist()))
    assert bool(torch.isfinite(p).all())
    assert bool((hist[0] <= rel0 * 1.5).all()) and bool((hist.min(dim=0).values < rel0).all())
